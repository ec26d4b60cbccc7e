#!/usr/bin/env python
"""bench.py — the headline metric of BASELINE.json on B200:

    pseudo-point logpdf+grad evals/s at N = 1M, M = 1024   (config 2: examples/dtc_example.jl-shaped
    single-output DTC pseudo-point GP, synthetic 1-D inputs, Matern-5/2)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one blocking evaluation of the DTC log-pdf and its gradient through the C ABI
(gpar_dtc_logpdf): Kuf/dKuf panel evaluation -> DMMA stream-K SYRK/GEMM -> M x M tail.
 * value : evals/s with X, Z, y resident in HBM (the optimiser's situation: dtc.jl:29-61 re-evaluates
           the closure on fixed data).
 * e2e   : the same call with HOST buffers re-uploaded every step (gpar_set_inputs/outputs from
           pinned memory inside the timed region) and the scalar + gradient read back.
 * N > 1 : one process per GPU (torchrun); the shards are independent objective evaluations
           (per-output conditional GPs x hyper-parameter restarts), no data-path collective; NCCL only
           all-gathers the (logpdf, gradient) scalars every step.  scaling = "weak".
 * --impl reference : the CPU restatement of the reference's algorithm (oracle port, all host
           threads) on a bounded sample, linearly extrapolated in N (every term is O(N)).
Extra keys report the other two BASELINE metrics measured in the same run (Kalman time-steps/s on
1024 x 10k independent Matern-5/2 sequences; the scaled-GPAR objective at N = 1M).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL, M_FULL = 1_000_000, 1024
THETA = np.log(np.array([1.0, 1.0, 0.1]))
METRIC = "pseudo-point logpdf+grad evals/s (N=1M,M=1024)"
FP64_PEAK_TFLOPS_FALLBACK = 36.96     # profiles/peaks_r01.json: DMMA m8n8k4 microbenchmark on this pool's B200
def _hbm_peak_gbps():
    """Measured copy bandwidth of this pool's B200 (driver-written MEASURED_PEAKS.json), else the value it held in round 1."""
    try:
        import json
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6554.2


HBM_PEAK_GBPS = _hbm_peak_gbps()
NCU_SYRK_DRAM_BYTES = 30.93e9          # dram__bytes_read.sum + dram__bytes_write.sum of panel_syrk_kernel, one ncu launch (profiles/ncu_syrk_traffic_r01c.csv)


def make_data(seed, n=N_FULL, m=M_FULL):
    """x ~ U(0,100) unsorted, z = linspace(min x, max x, M) (dtc_example.jl:82),
    y = sin(x) + 0.3 cos(3.1 x) + N(0, 0.1^2)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 100.0, n)
    z = np.linspace(x.min(), x.max(), m)
    y = np.sin(x) + 0.3 * np.cos(3.1 * x) + 0.1 * rng.normal(size=n)
    return x, z, y


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def cpu_port_eval(x, z, y, theta):
    """The oracle port of one logpdf+grad evaluation: BLAS-backed NumPy/SciPy float64 on all host
    threads (oracle.dtc.dtc_diag_value_and_grad_np, pinned on autograd of the oracle in tests/)."""
    from oracle.dtc import dtc_diag_value_and_grad_np
    return dtc_diag_value_and_grad_np(theta, x, z, y, 3)


def time_cpu_port(sample_n, steps, warmup, seed=1):
    x, z, y = make_data(seed)
    xs, ys = x[:sample_n], y[:sample_n]
    for _ in range(warmup):
        cpu_port_eval(xs, z, ys, THETA)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port_eval(xs, z, ys, THETA)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    evals_per_s = 1.0 / (dt * (N_FULL / sample_n))
    return evals_per_s, dt, os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_n = 31250     # 1/32 of the workload; ~1-3 s of all-core CPU work per step
    steps = max(1, min(args.steps, 5)); warmup = max(1, min(args.warmup, 2))
    v, dt, cores = time_cpu_port(sample_n, steps, warmup)
    sample = "N=%d of %d (1/32), M=%d, all data; %.2f s per sampled step, linearly extrapolated in N" % (sample_n, N_FULL, M_FULL, dt)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "evals/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "dtc_logpdf_grad N=1000000 M=1024 D=1 Matern52 (BASELINE configs[1])", "note": "CPU restatement of the reference algorithm (no Julia in the image); the reference itself has no gradient"},
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gpar_at_scale_b200 as gp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, W = args.steps, max(args.warmup, 3)

    x, z, y = make_data(1 + rank)        # each rank: its own independent problem of the named shape (weak scaling)
    # pinned host staging (what the e2e leg uploads from every step)
    xh = torch.empty(N_FULL, dtype=torch.float64).pin_memory(); xh.numpy()[:] = x
    yh = torch.empty(N_FULL, dtype=torch.float64).pin_memory(); yh.numpy()[:] = y
    xp, yp = xh.numpy(), yh.numpy()
    ctx = gp.Context(local)
    ctx.set_inputs(xp); ctx.set_pseudo(z); ctx.set_outputs(yp)
    gathered = [torch.zeros(4, dtype=torch.float64, device="cuda") for _ in range(world)] if world > 1 else None

    def step_resident():
        val, grad = ctx.dtc_logpdf(gp.MATERN52, THETA, grad=True)
        if world > 1:      # NCCL over NVLink only gathers the scalars
            dist.all_gather(gathered, torch.tensor([val, *grad], dtype=torch.float64, device="cuda"))
        return val, grad

    def step_e2e():
        ctx.set_inputs(xp); ctx.set_outputs(yp)
        return step_resident()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(W):
        step_resident()
    sampler = ClockSampler(local); sampler.start()
    launches = 0; dev_ms = []; syrk_ms = []; prod_ms = []; tail_ms = []
    sync(); t0 = time.perf_counter()
    for _ in range(K):
        val, grad = step_resident()
        ms, nl = ctx.last_timing(); ph = ctx.last_profile()
        launches += nl; dev_ms.append(ms); prod_ms.append(ph[0]); syrk_ms.append(ph[1]); tail_ms.append(ph[2])
    sync(); t1 = time.perf_counter()
    sampler.stop_flag.set(); sampler.join(timeout=2)
    step_s = (t1 - t0) / K
    for _ in range(2):
        step_e2e()
    sync(); t2 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    sync(); t3 = time.perf_counter()
    e2e_s = (t3 - t2) / K
    if world > 1:
        tt = torch.tensor([step_s, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_s, e2e_s = tt.tolist()

    extra = {}
    if rank == 0 and not args.no_extra:
        # the other two BASELINE metrics, same run (device-resident, CUDA-event timed by the library)
        rng = np.random.default_rng(2)
        B, NK = 1024, 10000
        tk = np.cumsum(rng.exponential(1 / 30, NK)); Yk = rng.normal(size=(B, NK))
        ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
        ctx.set_times(tk); ctx.set_outputs(Yk)
        ts = []
        for i in range(8):
            ctx.lgssm_logpdf(gp.MATERN52, ths); ts.append(ctx.last_timing()[0])
        kal_ms = float(np.median(ts[3:]))
        ts = []
        for i in range(4):
            ctx.lgssm_smooth(gp.MATERN52, ths[0]); ts.append(ctx.last_timing()[0])
        extra["kalman_filter_steps_per_s"] = B * NK / kal_ms * 1e3
        extra["kalman_filter_ms_1024x10k"] = kal_ms
        extra["kalman_smoother_steps_per_s"] = B * NK / float(np.median(ts[1:])) * 1e3

        def med_ms(fn, n=6, skip=2):
            out = []
            for _ in range(n):
                fn(); out.append(ctx.last_timing()[0])
            return float(np.median(out[skip:]))
        extra["kalman_logpdf_grad_ms_1024x10k"] = med_ms(lambda: ctx.lgssm_logpdf_grad(gp.MATERN52, ths), 4, 1)
        # the same 1024 sequences under ONE model (batching over data, the reference's M+1-column / MC-sample pattern)
        extra["kalman_filter_shared_model_steps_per_s"] = B * NK / med_ms(lambda: ctx.lgssm_logpdf(gp.MATERN52, ths[0])) * 1e3
        # the same batch on the regular grid range(0, step = 1/30) (toy_data.jl:6): steady-state path
        ctx.set_times_range(0.0, 1 / 30, NK)
        ms = med_ms(lambda: ctx.lgssm_logpdf(gp.MATERN52, ths))
        extra["kalman_filter_regular_grid_steps_per_s"] = B * NK / ms * 1e3
        # one 10M-step Matern-5/2 sequence (north-star shape): irregular grid, then the regular grid
        N10 = 10_000_000
        y10 = rng.normal(size=N10); th3 = np.log(np.array([1.0, 1.0, 0.1]))
        ctx.set_outputs(y10); ctx.set_times(np.cumsum(rng.exponential(1 / 30, N10)))
        ms = med_ms(lambda: ctx.lgssm_logpdf(gp.MATERN52, th3))
        extra["kalman_filter_1x10M_steps_per_s"] = N10 / ms * 1e3
        extra["kalman_filter_1x10M_ms"] = ms
        ctx.set_times_range(0.0, 1 / 30, N10)
        ms = med_ms(lambda: ctx.lgssm_logpdf(gp.MATERN52, th3))
        extra["kalman_filter_1x10M_regular_grid_steps_per_s"] = N10 / ms * 1e3
        extra["kalman_filter_1x10M_regular_grid_ms"] = ms
        del y10
        # sixteen 10M-step sequences, each with its own model (hyper-parameter candidates): the HBM-bound shape of
        # the single-pass steady-state filter; algorithmic traffic 8 B/step (y read once), peak = measured copy bandwidth
        B8 = 16
        ctx.set_outputs(rng.normal(size=(B8, N10)))
        ths8 = np.tile(th3, (B8, 1)) + 0.05 * rng.normal(size=(B8, 3))
        ms = med_ms(lambda: ctx.lgssm_logpdf(gp.MATERN52, ths8))
        extra["kalman_filter_16x10M_regular_grid_steps_per_s"] = B8 * N10 / ms * 1e3
        extra["kalman_filter_16x10M_regular_grid_ms"] = ms
        extra["kalman_filter_16x10M_regular_grid_hbm"] = {"achieved_GBps": B8 * N10 * 8 / (ms * 1e-3) / 1e9, "peak_GBps": HBM_PEAK_GBPS,
                                                         "frac": B8 * N10 * 8 / (ms * 1e-3) / 1e9 / HBM_PEAK_GBPS,
                                                         "note": "whole blocking call (set-up, head, main pass with fused finish), 8 B/step algorithmic"}
        tfull = np.arange(N_FULL) / 30.0
        ctx.set_inputs(xp); ctx.set_outputs(yp); ctx.set_times(tfull)
        th5 = np.log(np.array([1.0, 1.0, 1.0, 1.0, 0.1]))
        ts = []
        for i in range(5):
            ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, th5); ts.append(ctx.last_timing()[0])
        extra["scaled_gpar_objective_ms_N1M_M1024"] = float(np.median(ts[2:]))
        ts = []
        for i in range(4):
            ctx.scaled_dtc_grad(gp.MATERN52, gp.MATERN52, th5); ts.append(ctx.last_timing()[0])
        extra["scaled_gpar_objective_and_grad_ms_N1M_M1024"] = float(np.median(ts[1:]))

    if rank == 0:
        flops = N_FULL * M_FULL * (M_FULL + 1) + 2.0 * N_FULL * M_FULL * M_FULL     # G (symmetric) + H (forward-mode dG/dl)
        syrk = float(np.mean(syrk_ms))
        achieved = flops / (syrk * 1e-3) / 1e12
        peak = FP64_PEAK_TFLOPS_FALLBACK
        do_cpu = (not args.no_cpu) and world == 1
        cpu_v, cpu_dt, cores = time_cpu_port(31250, 2, 1) if do_cpu else (None, None, 0)
        line = {
            "metric": METRIC, "value": world / step_s, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "dtc_logpdf_grad N=1000000 M=1024 D=1 Matern52 (BASELINE configs[1])", "theta": THETA.tolist(),
                       "jitter": "sigma^2 (dtc.jl:35)", "l2": "inputs larger than L2: 2 x 8.4 GB operand panels streamed per step",
                       "parallelism": "independent evaluations per GPU (restarts/outputs); NCCL all-gather of scalars only"},
            "clocks": sampler.summary(),
            "e2e": {"value": world / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(16 * N_FULL), "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": NCU_SYRK_DRAM_BYTES,
                         "kernel": "panel_syrk_kernel (FP64 DMMA.8x8x4)", "traffic_note": "operand panels: 16.8 GB unique; L2 hit 74% with the two-level phase-aligned stream-K (plain stream-K: 181.9 GB)", "kernel_ms": syrk, "algorithmic_flops_per_launch": flops,
                         "peak_source": "FP64 DMMA peak measured on this pool (profiles/peaks_r01.json; cuBLAS DGEMM 8192^3 = 35.9); MEASURED_PEAKS.json has no FP64 entry",
                         "step_breakdown_ms": {"panel_producer": float(np.mean(prod_ms)), "dmma_syrk": syrk, "reduce_and_tail": float(np.mean(tail_ms)),
                                               "device_total": float(np.mean(dev_ms))}},
            "cpu_baseline": None if not do_cpu else {"value": cpu_v, "unit": "evals/s", "cores": cores, "kind": "port",
                                                       "sample": "N=31250 of 1000000 (1/32), M=1024; %.2f s per sampled eval, linearly extrapolated in N" % cpu_dt},
            "value_check": {"logpdf": val, "grad": list(map(float, grad))},
            "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the Kalman / scaled-GPAR extra metrics")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
