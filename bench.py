#!/usr/bin/env python
"""bench.py — the three metrics of BASELINE.json on B200, in one run:

    pseudo-point logpdf+grad evals/s (N=1M, M=1024)   <- the headline line (`metric`, `value`, `e2e`, `roofline`)
    Kalman time-steps/s                                 <- extra.kalman_* (with their own roofline blocks)
    GPAR fit s                                          <- extra.gpar_fit (BASELINE configs[4], STRONG scaling over --gpus)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-fit] [--no-extra] [--no-cpu]

A "step" of the headline is one blocking evaluation of the DTC log-pdf and its gradient through the C ABI
(gpar_dtc_logpdf): Kuf/dKuf panel evaluation -> DMMA stream-K SYRK/GEMM -> M x M tail (configs[1]:
examples/dtc_example.jl-shaped single-output pseudo-point GP, synthetic 1-D inputs, Matern-5/2).
 * value : evals/s with X, Z, y resident in HBM (the optimiser's situation: dtc.jl:29-61 re-evaluates the closure on
           fixed data).
 * e2e   : the same call with HOST buffers re-uploaded every step (gpar_set_inputs/outputs from pinned memory inside
           the timed region) and the scalar + gradient read back.
 * N > 1 : one process per GPU (torchrun); the shards are independent objective evaluations (per-output conditional
           GPs x hyper-parameter restarts), no data-path collective; NCCL only all-gathers the (logpdf, gradient)
           scalars every step.  scaling = "weak".
 * roofline.peak is MEASURED IN THIS RUN (gpar_measure_peaks: register-only DMMA loop, ~50 ms).
 * extra.gpar_fit : the 8-output GPAR chain fit of configs[4] (examples/GPAR_scaled_examples.jl:132-175 call pattern):
           N = 2 097 152, M = 2048, 8 restarts = 64 independent (output, restart) Nelder-Mead tasks with a FIXED
           iteration budget, partitioned over the ranks — total work fixed, so this is the strong-scaling curve of the
           run.  With >= 2 GPUs rank 0 then repeats the restart-0 tasks through the one-process group path
           (gpar_group_fit over all devices) and checks that both paths reach identical optima.
 * --impl reference : the CPU restatement of the reference's algorithm (oracle port, all host threads) on a bounded
           sample, linearly extrapolated in N (every term is O(N)).
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is "the CPU path with all the host threads it can use"
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FULL, M_FULL = 1_000_000, 1024
THETA = np.log(np.array([1.0, 1.0, 0.1]))
METRIC = "pseudo-point logpdf+grad evals/s (N=1M,M=1024)"
FP64_PEAK_TFLOPS_FALLBACK = 36.96     # profiles/peaks_r01.json; used only if the in-run measurement fails
FIT = {"outputs": 8, "N": 2_097_152, "M": 2048, "restarts": 8, "iterations": 10}


def _hbm_peak_gbps():
    """Measured copy bandwidth of this pool's B200 (driver-written MEASURED_PEAKS.json), else the value it held in round 1."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6554.2


HBM_PEAK_GBPS = _hbm_peak_gbps()


def ncu_traffic(kernel="panel_syrk_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the newest committed ncu capture
    (profiles/ncu_syrk_traffic_r*.csv) -> (bytes, file name); (None, None) if there is none."""
    best = (None, None)
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_syrk_traffic_r*.csv"))):
        rd = wr = None
        try:
            for line in open(path):
                if kernel not in line:
                    continue
                cols = [c.strip('"') for c in line.rstrip("\n").split('","')]
                if "dram__bytes_read.sum" in cols and rd is None:
                    rd = float(cols[-1].replace(",", "").strip('"'))
                if "dram__bytes_write.sum" in cols and wr is None:
                    wr = float(cols[-1].replace(",", "").strip('"'))
        except Exception:
            continue
        if rd is not None and wr is not None:
            best = (rd + wr, os.path.basename(path))
    return best


def make_data(seed, n=N_FULL, m=M_FULL):
    """x ~ U(0,100) unsorted, z = linspace(min x, max x, M) (dtc_example.jl:82),
    y = sin(x) + 0.3 cos(3.1 x) + N(0, 0.1^2)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.0, 100.0, n)
    z = np.linspace(x.min(), x.max(), m)
    y = np.sin(x) + 0.3 * np.cos(3.1 * x) + 0.1 * rng.normal(size=n)
    return x, z, y


def synth_chain(P, N, seed=4):
    """the big-set recipe of src/data/toy_data.jl:79-87 extended to P outputs, y_i = f(x, y_<i) + noise
    (std 0.64), on the regular grid x = k/30 (toy_data.jl:6)."""
    rng = np.random.default_rng(seed)
    x = np.arange(N) / 30.0
    Y = np.zeros((P, N))
    Y[0] = 3 - np.sin(np.pi / 10 * (x + 1) * 0.01) - (x * 0.01) ** 0.3 + 0.64 * rng.normal(size=N)
    for i in range(1, P):
        Y[i] = np.cos(Y[i - 1]) ** 2 + np.sin(np.pi / 20 * x * 0.01 * (i + 1)) + 0.1 * Y[max(i - 2, 0)] + 0.64 * rng.normal(size=N)
    return x, Y


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


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get("num_threads", 1)) for p in threadpool_info()] or [1])
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


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
            "config": {"workload": "dtc_logpdf_grad N=1000000 M=1024 D=1 Matern52 (BASELINE configs[1])", "note": "CPU restatement of the reference algorithm (no Julia in the image); the reference itself has no gradient; one CPU run whatever --gpus says"},
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "blas_threads": blas_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def kalman_roofline(steps, ms, bytes_per_step, peaks, note):
    """a Kalman step is 250 FP64 flop by the sequential count and `bytes_per_step` algorithmic HBM bytes; report
    both fractions and name the binding roof (the larger fraction)."""
    tf = steps * 250.0 / (ms * 1e-3) / 1e12
    gb = steps * bytes_per_step / (ms * 1e-3) / 1e9
    f_fp, f_hbm = tf / peaks["dfma_tflops"], gb / HBM_PEAK_GBPS
    return {"ms": ms, "steps_per_s": steps / ms * 1e3, "bound": "fp64" if f_fp >= f_hbm else "hbm",
            "fp64": {"achieved_TFLOPs": tf, "peak_TFLOPs": peaks["dfma_tflops"], "frac": f_fp, "flop_per_step": 250},
            "hbm": {"achieved_GBps": gb, "peak_GBps": HBM_PEAK_GBPS, "frac": f_hbm, "bytes_per_step": bytes_per_step},
            "frac": max(f_fp, f_hbm), "note": note}


def run_extra(ctx, gp, xp, yp, peaks):
    """The Kalman metric (BASELINE configs[2] and the north star's 10M-step sequence) and the scaled-GPAR objective at
    N = 1M, device-resident, CUDA-event timed by the library."""
    extra = {}
    rng = np.random.default_rng(2)
    B, NK = 1024, 10000
    tk = np.cumsum(rng.exponential(1 / 30, NK)); Yk = rng.normal(size=(B, NK))
    ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
    ctx.set_times(tk); ctx.set_outputs(Yk)

    def med_ms(fn, n=7, skip=3):
        out = []
        for _ in range(n):
            fn(); out.append(ctx.last_timing()[0])
        return float(np.median(out[skip:]))

    kal_ms = med_ms(lambda: ctx.lgssm_logpdf(gp.MATERN52, ths), 8, 3)
    extra["kalman_filter_steps_per_s"] = B * NK / kal_ms * 1e3
    extra["kalman_filter_ms_1024x10k"] = kal_ms
    extra["kalman_filter_1024x10k_roofline"] = kalman_roofline(
        B * NK, kal_ms, 8, peaks, "cfg 3: 1024 independent Matern-5/2 models x 10k steps on one irregular grid; y read once (8 B/step), t shared")
    sm_ms = med_ms(lambda: ctx.lgssm_smooth(gp.MATERN52, ths[0], keep_on_device=True), 5, 2)
    extra["kalman_smoother_steps_per_s"] = B * NK / sm_ms * 1e3
    extra["kalman_smoother_1024x10k_roofline"] = kalman_roofline(
        B * NK, sm_ms, 24, peaks, "cfg 3 filter + RTS smoother, 1024 sequences sharing one model: y in, mean out per sequence (+ shared var): 16-24 B/step; 250 flop/step is the per-model count — with a shared model the per-sequence work is ~30 flop/step, so HBM is the binding roof")
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
    extra["kalman_filter_1x10M_irregular_roofline"] = kalman_roofline(
        N10, ms, 16, peaks, "one 10M-step sequence, irregular grid: (t_k, y_k) read once = 16 B/step")
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
    extra["scaled_gpar_objective_ms_N1M_M1024"] = med_ms(lambda: ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, th5), 5, 2)
    ms = med_ms(lambda: ctx.scaled_dtc_grad(gp.MATERN52, gp.MATERN52, th5), 4, 1)
    extra["scaled_gpar_objective_and_grad_ms_N1M_M1024"] = ms
    fl = N_FULL * M_FULL * (M_FULL + 1) + 2.0 * N_FULL * M_FULL * M_FULL      # SYRK (symmetric) + S = beta P
    extra["scaled_gpar_objective_and_grad_roofline"] = {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peaks["dmma_tflops"], "unit": "TFLOP/s",
                                                        "frac": fl / (ms * 1e-3) / 1e12 / peaks["dmma_tflops"],
                                                        "note": "whole blocking call; DMMA work = N M (M+1) + 2 N M^2 (panel_syrk_kernel + panel_gemm_kernel, no library GEMM)"}
    # the reference's own problem size (GPAR_scaled_examples.jl: N = 8 496, M = 50): one candidate per call vs 64 candidates
    # (simplex vertices x restarts of dtc.jl:58-61) in one fused launch sequence (gpar_scaled_dtc_batch, scaled_small.cu)
    ns, msm = 8496, 50
    ts = np.arange(ns) / 30.0; xs = rng.normal(size=(ns, 1)); zs = np.linspace(xs.min(), xs.max(), msm)[:, None]
    ctx.set_inputs(xs); ctx.set_pseudo(zs); ctx.set_times(ts); ctx.set_outputs(np.sin(xs[:, 0]) + 0.3 * rng.normal(size=ns))
    th64 = np.tile(np.log([1.0, 1.0, 1.0, 1.0, 0.6]), (64, 1)) + 0.05 * rng.normal(size=(64, 5))

    def wall_ms(fn, n=20, skip=3):
        out = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); out.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(out[skip:]))
    one = wall_ms(lambda: ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, th64[0]))
    b64 = wall_ms(lambda: ctx.scaled_dtc_batch(gp.MATERN52, gp.MATERN52, th64))
    extra["scaled_gpar_reference_size_N8496_M50"] = {"ms_per_candidate_single_call": one, "ms_per_64_candidates_batched": b64,
                                                     "ms_per_candidate_batched": b64 / 64, "speedup": one * 64 / b64,
                                                     "note": "host wall-clock through the C ABI, including the result read-back"}
    return extra


def run_gpar_fit(ctx, gp, world, rank, local, iterations, side_group):
    """BASELINE configs[4]: 8 outputs x 8 restarts = 64 independent Nelder-Mead fits (fixed iteration budget) over the ranks."""
    import torch
    import torch.distributed as dist
    from gpar_at_scale_b200 import chain
    P, N, M, R = FIT["outputs"], FIT["N"], FIT["M"], FIT["restarts"]
    t, Y = synth_chain(P, N)
    X = np.ascontiguousarray(Y[:1].T)          # warm-up: one evaluation allocates the panels and plans the SYRK
    ctx.set_inputs(X); ctx.set_pseudo(chain.strided_pseudo_inputs(X, M)); ctx.set_times(t); ctx.set_outputs(Y[1]); ctx.set_noise_vector(None)
    ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, np.zeros(5))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    sync(); t0 = time.perf_counter()
    best, info = chain.fit_chain(t, Y, M, n_restarts=R, iterations=iterations, seed=4, ctx=ctx)
    sync(); dt = time.perf_counter() - t0
    stat = torch.tensor([float(info["objective_evals_this_rank"]), info["busy_seconds"], dt], dtype=torch.float64, device="cuda")
    if world > 1:
        stats = [torch.zeros_like(stat) for _ in range(world)]
        dist.all_gather(stats, stat)
        stats = [s.tolist() for s in stats]
    else:
        stats = [stat.tolist()]
    evals = [int(s[0]) for s in stats]; busy = [s[1] for s in stats]; dt = max(s[2] for s in stats)
    vals, thetas = info["minimum"], info["minimizer"]
    out = {"metric": "GPAR fit s", "seconds": dt, "scaling": "strong", "n_gpus": world,
           "workload": "GPAR chain fit: outputs=%d N=%d M=%d restarts=%d -> %d (output, restart) tasks, Nelder-Mead with a fixed budget of %d iterations per task (Optim.jl defaults), Matern-5/2 time and output kernels; output 1 is the time-only state-space GP (BASELINE configs[4], GPAR_scaled_examples.jl:132-175)" % (P, N, M, R, P * R, iterations),
           "objective_evaluations_total": int(sum(evals)), "objective_evaluations_per_rank": evals,
           "busy_seconds_per_rank": busy, "idle_fraction_slowest_vs_mean": 1.0 - float(np.mean(busy)) / max(busy),
           "evaluations_per_task_mean": float(sum(evals)) / (P * R),
           "best_nlml_per_output": {str(o): repr(best[o][0]) for o in sorted(best)},
           "optima_checksum": repr(float(np.sum(vals))),
           "identical_optima_note": "every task is a deterministic Nelder-Mead run on its own data: best_nlml_per_output / optima_checksum must be bit-identical for every --gpus"}
    # ONE scaled objective (the last output's, D = P - 1) with its ROWS sharded over the ranks: one process per GPU, the two
    # (gradient: three) exchanges through torch.distributed / NCCL (parallel.scaled_dtc_row_sharded over the gpar_scaled_slice_*
    # C ABI) — strong scaling of a single evaluation; the one-device time of the same call is in sharded_scaled_objective below
    if world > 1:
        from gpar_at_scale_b200 import parallel
        o = P - 1
        Xo = np.ascontiguousarray(Y[:o].T); Zo = chain.strided_pseudo_inputs(Xo, M)
        th5 = np.log([2.0, 0.5, 1.0, 1.0, 0.1])
        bnd = parallel.row_slice_bounds(N, world)
        ctx.set_times(t); ctx.set_outputs(Y[o]); ctx.set_pseudo(Zo); ctx.set_inputs(np.ascontiguousarray(Xo[bnd[rank]:bnd[rank + 1]])); ctx.set_noise_vector(None)
        dev = torch.device("cuda", local)

        def timed(fn, n, skip):
            ts = []
            for _ in range(n):
                sync(); t1 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t1)
            tt = torch.tensor([float(np.median(ts[skip:]))], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return r, float(tt.item()) * 1e3
        vsh, ms_v = timed(lambda: parallel.scaled_dtc_row_sharded(ctx, gp.MATERN52, gp.MATERN52, th5, bnd[rank], device=dev), 5, 2)
        (vgs, g5s), ms_g = timed(lambda: parallel.scaled_dtc_row_sharded(ctx, gp.MATERN52, gp.MATERN52, th5, bnd[rank], grad=True, device=dev), 4, 1)
        out["sharded_scaled_objective_torchrun"] = {
            "path": "parallel.scaled_dtc_row_sharded: %d ranks (one process per GPU), rows of ONE objective (N=%d, M=%d, D=%d) sliced; NCCL all-gather of slice summaries + all-reduce of (G, g) [+ all-gather of tangent summaries]" % (world, N, M, o),
            "ms_value": ms_v, "ms_value_and_grad": ms_g, "value": repr(vsh), "grad": [float(x) for x in g5s], "scaling": "strong",
            "timing": "max over ranks of the median wall-clock per call, barrier + device synchronisation on both sides"}
    # one process, all devices: the restart-0 task of every output through gpar_group_fit (dynamic hand-out over the
    # members, C++ Nelder-Mead twin) — must reproduce the torchrun path's optima for those tasks exactly
    if world > 1:
        if rank == 0:
            g = gp.Group(list(range(world)))
            tasks = []
            for o in range(P):
                th0 = np.random.default_rng([4, o, 0]).random(3 if o == 0 else 5)
                if o == 0:
                    tasks.append({"X": None, "Z": None, "y": Y[0], "theta0": th0})
                else:
                    Xo = np.ascontiguousarray(Y[:o].T)
                    tasks.append({"X": Xo, "Z": chain.strided_pseudo_inputs(Xo, M), "y": Y[o], "theta0": th0})
            tg = time.perf_counter()
            gmin, gth, calls, member = g.fit(t, tasks, gp.MATERN52, gp.MATERN52, iterations)
            tg = time.perf_counter() - tg
            ref = np.array([vals[o * R] for o in range(P)])
            out["group_fit"] = {"path": "gpar_group_fit, one process, %d devices, %d tasks (restart 0 of every output)" % (world, P), "seconds": tg,
                                "f_calls": calls.tolist(), "member_of": member.tolist(), "members_used": int(len(set(member.tolist()))),
                                "max_rel_diff_vs_torchrun_path": float(np.max(np.abs(gmin - ref) / np.abs(ref))),
                                "identical_optima": bool(np.all(gmin == ref))}
            # one scaled objective with its ROWS sharded over the devices (gpar_group_scaled_dtc_sharded): the last output's problem
            # (D = P - 1 observed outputs as inputs), one all-gather of slice summaries + one all-reduce of (G, g) per evaluation,
            # against the same evaluation on one device (this rank's context)
            try:
                o = P - 1
                Xo = np.ascontiguousarray(Y[:o].T); Zo = chain.strided_pseudo_inputs(Xo, M)
                th5 = np.log([2.0, 0.5, 1.0, 1.0, 0.1])
                ctx.set_inputs(Xo); ctx.set_pseudo(Zo); ctx.set_times(t); ctx.set_outputs(Y[o]); ctx.set_noise_vector(None)
                one = []
                for _ in range(4):
                    t1 = time.perf_counter(); v1 = ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, th5); one.append(time.perf_counter() - t1)
                lo = g.load_row_slices(Xo, Zo, t, Y[o])
                sh = []
                for _ in range(6):
                    t1 = time.perf_counter(); vs = g.scaled_dtc_sharded(gp.MATERN52, gp.MATERN52, th5, lo); sh.append(time.perf_counter() - t1)
                out["sharded_scaled_objective"] = {
                    "path": "gpar_group_scaled_dtc_sharded, one process, %d devices, rows of ONE objective (N=%d, M=%d, D=%d) sliced" % (world, N, M, o),
                    "ms_one_device": float(np.median(one[1:]) * 1e3), "ms_sharded": float(np.median(sh[2:]) * 1e3),
                    "speedup": float(np.median(one[1:]) / np.median(sh[2:])), "value_one_device": repr(v1), "value_sharded": repr(vs),
                    "rel_diff": float(abs(vs - v1) / abs(v1)), "scaling": "strong",
                    "collective_bytes_per_evaluation": int(8 * (M * M + M + world * (9 + 3 * M)))}
                tr = out.get("sharded_scaled_objective_torchrun")
                if tr:
                    tr["ms_value_one_device"] = out["sharded_scaled_objective"]["ms_one_device"]
                    tr["speedup_value"] = tr["ms_value_one_device"] / tr["ms_value"]
                    tr["value_rel_diff_vs_one_device"] = float(abs(float(tr["value"]) - v1) / abs(v1))
                try:        # value + gradient, row-sharded (a second all-gather carries the tangent states over the slice boundaries)
                    shg = []
                    for _ in range(4):
                        t1 = time.perf_counter(); vg, gg = g.scaled_dtc_sharded(gp.MATERN52, gp.MATERN52, th5, lo, grad=True); shg.append(time.perf_counter() - t1)
                    out["sharded_scaled_objective"]["ms_sharded_value_and_grad"] = float(np.median(shg[1:]) * 1e3)
                    out["sharded_scaled_objective"]["grad_sharded"] = [float(x) for x in gg]
                    out["sharded_scaled_objective"]["value_rel_diff_grad_call"] = float(abs(vg - v1) / abs(v1))
                except Exception as e:
                    out["sharded_scaled_objective"]["grad_error"] = repr(e)
                # weak scaling of the same call: 1 048 576 rows per device (8 GPUs: N = 8 388 608, M = 2048 — an operand panel of
                # 137 GB that no single device could hold next to this bench's own buffers)
                Nw = 1_048_576 * world
                rw = np.random.default_rng(9)
                tw = np.arange(Nw) / 30.0
                Xw = np.cumsum(rw.normal(size=(Nw, 2)), axis=0) / np.sqrt(Nw) * 3 + 0.3 * rw.normal(size=(Nw, 2))
                Zw = chain.strided_pseudo_inputs(Xw, M)
                yw = np.sin(Xw[:, 0]) + 0.5 * np.sin(0.05 * tw) + 0.1 * rw.normal(size=Nw)
                low = g.load_row_slices(Xw, Zw, tw, yw)
                shw = []
                for _ in range(5):
                    t1 = time.perf_counter(); vw = g.scaled_dtc_sharded(gp.MATERN52, gp.MATERN52, th5, low); shw.append(time.perf_counter() - t1)
                out["sharded_scaled_objective_weak"] = {
                    "path": "gpar_group_scaled_dtc_sharded, %d devices, 1048576 rows per device: N=%d, M=%d, D=2" % (world, Nw, M),
                    "ms_sharded": float(np.median(shw[2:]) * 1e3), "value": repr(vw), "scaling": "weak",
                    "panel_bytes_total": int(Nw) * M * 8}
            except Exception as e:      # the fit numbers above stand on their own
                out.setdefault("sharded_scaled_objective", {})["error"] = repr(e)
            g.close()
        dist.barrier(group=side_group)       # host-side wait (gloo): no NCCL kernel spins on the GPUs the group leg uses
    return out


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
    side_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        side_group = dist.new_group(backend="gloo")
    K, W = args.steps, max(args.warmup, 3)

    x, z, y = make_data(1 + rank)        # each rank: its own independent problem of the named shape (weak scaling)
    # pinned host staging (what the e2e leg uploads from every step)
    xh = torch.empty(N_FULL, dtype=torch.float64).pin_memory(); xh.numpy()[:] = x
    yh = torch.empty(N_FULL, dtype=torch.float64).pin_memory(); yh.numpy()[:] = y
    xp, yp = xh.numpy(), yh.numpy()
    ctx = gp.Context(local)
    try:
        peaks = ctx.measure_peaks(); peaks["source"] = "measured in this run (gpar_measure_peaks: register-only mma.sync.m8n8k4.f64 / DFMA loops, best of 3; 1 GiB copy)"
    except Exception as e:      # noqa: BLE001
        peaks = {"dmma_tflops": FP64_PEAK_TFLOPS_FALLBACK, "dfma_tflops": FP64_PEAK_TFLOPS_FALLBACK, "hbm_copy_gbs": HBM_PEAK_GBPS,
                 "source": "fallback profiles/peaks_r01.json (%s)" % e}
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
        extra = run_extra(ctx, gp, xp, yp, peaks)
    if not args.no_fit:
        fit = run_gpar_fit(ctx, gp, world, rank, local, args.fit_iterations, side_group)
        if rank == 0:
            extra["gpar_fit"] = fit

    if rank == 0:
        flops = N_FULL * M_FULL * (M_FULL + 1) + 2.0 * N_FULL * M_FULL * M_FULL     # G (symmetric) + H (forward-mode dG/dl)
        syrk = float(np.mean(syrk_ms))
        achieved = flops / (syrk * 1e-3) / 1e12
        peak = peaks["dmma_tflops"]
        traffic, traffic_src = ncu_traffic()
        do_cpu = (not args.no_cpu) and world == 1
        cpu_v, cpu_dt, cores = time_cpu_port(31250, 2, 1) if do_cpu else (None, None, 0)
        line = {
            "metric": METRIC, "value": world / step_s, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "dtc_logpdf_grad N=1000000 M=1024 D=1 Matern52 (BASELINE configs[1])", "theta": THETA.tolist(),
                       "jitter": "sigma^2 (dtc.jl:35)", "l2": "inputs larger than L2: 2 x 8.4 GB operand panels streamed per step",
                       "parallelism": "independent evaluations per GPU (restarts/outputs); NCCL all-gather of scalars only",
                       "strong_scaling_metric": "extra.gpar_fit.seconds (BASELINE configs[4], fixed total work over --gpus)"},
            "clocks": sampler.summary(),
            "e2e": {"value": world / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": int(16 * N_FULL), "d2h_bytes_per_step": 32},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "panel_syrk_kernel (FP64 DMMA.8x8x4)",
                         "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, parsed from profiles/%s; operand panels: 16.8 GB unique (phase-aligned stream-K keeps the re-reads in L2)" % traffic_src,
                         "kernel_ms": syrk, "algorithmic_flops_per_launch": flops,
                         "peak_source": peaks["source"], "peaks": peaks,
                         "step_breakdown_ms": {"panel_producer": float(np.mean(prod_ms)), "dmma_syrk": syrk, "reduce_and_tail": float(np.mean(tail_ms)),
                                               "device_total": float(np.mean(dev_ms))}},
            "cpu_baseline": None if not do_cpu else {"value": cpu_v, "unit": "evals/s", "cores": cores, "blas_threads": blas_threads(), "kind": "port",
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
    ap.add_argument("--no-fit", action="store_true", help="skip the GPAR chain fit (BASELINE configs[4], ~5 min on one GPU)")
    ap.add_argument("--fit-iterations", type=int, default=FIT["iterations"], help="Nelder-Mead iteration budget per (output, restart) task")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
