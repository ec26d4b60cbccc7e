"""GPAR fit at scale (BASELINE config 5 shape): P-output GPAR, N time steps, M pseudo-points per
output, R hyper-parameter restarts, a FIXED Nelder-Mead iteration budget instead of the reference's
wall-clock limits (GPAR_scaled_examples.jl:140,173).  Tasks (output, restart) are sharded over the
ranks (one process per GPU, torchrun); NCCL only gathers (minimum, minimizer) per task.

    python -m torch.distributed.run --nproc-per-node G tools/bench_gpar_fit.py --npoints 2097152 --pseudo 2048 --outputs 8 --restarts 2 --iterations 10
Prints one JSON line on rank 0: {"metric": "GPAR fit s", ...}.
"""
import argparse
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synth(P, N, seed=4):
    """Big-set recipe of src/data/toy_data.jl:79-87 extended to P outputs: y_i = f(x, y_<i) + noise."""
    rng = np.random.default_rng(seed)
    x = np.arange(N) / 30.0
    Y = np.zeros((P, N))
    Y[0] = 3 - np.sin(np.pi / 10 * (x + 1) * 0.01) - (x * 0.01) ** 0.3 + 0.64 * rng.normal(size=N)
    for i in range(1, P):
        Y[i] = np.cos(Y[i - 1]) ** 2 + np.sin(np.pi / 20 * x * 0.01 * (i + 1)) + 0.1 * Y[max(i - 2, 0)] + 0.64 * rng.normal(size=N)
    return x, Y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npoints", type=int, default=2097152); ap.add_argument("--pseudo", type=int, default=2048)
    ap.add_argument("--outputs", type=int, default=8); ap.add_argument("--restarts", type=int, default=2)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--optimizer", default="neldermead", choices=["neldermead", "lbfgs"])
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import gpar_at_scale_b200 as gp
    from gpar_at_scale_b200 import chain
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gp.Context(local)
    t, Y = synth(a.outputs, a.npoints)
    # warm-up: one objective evaluation allocates the panels
    X = np.ascontiguousarray(Y[:1].T)
    ctx.set_inputs(X); ctx.set_pseudo(chain.strided_pseudo_inputs(X, a.pseudo)); ctx.set_times(t); ctx.set_outputs(Y[1])
    ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, np.zeros(5))
    if a.optimizer == "lbfgs":
        ctx.scaled_dtc_grad(gp.MATERN52, gp.MATERN52, np.zeros(5))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    best, info = chain.fit_chain(t, Y, a.pseudo, n_restarts=a.restarts, iterations=a.iterations, seed=4, ctx=ctx, optimizer=a.optimizer)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    ev = torch.tensor([float(info["objective_evals_this_rank"]), dt], dtype=torch.float64, device="cuda")
    if world > 1:
        evs = [torch.zeros_like(ev) for _ in range(world)]
        dist.all_gather(evs, ev)
        evals = [int(e[0].item()) for e in evs]; dt = max(e[1].item() for e in evs)
    else:
        evals = [int(ev[0].item())]
    if rank == 0:
        print(json.dumps({"metric": "GPAR fit s", "value": dt, "unit": "s", "n_gpus": world, "higher_is_better": False, "scaling": "strong",
                          "config": {"workload": "gpar_fit outputs=%d N=%d M=%d restarts=%d %s_iterations=%d" % (a.outputs, a.npoints, a.pseudo, a.restarts, a.optimizer, a.iterations)},
                          "objective_evals_per_rank": evals, "tasks": info["tasks"],
                          "best_nlml": {str(o): best[o][0] for o in sorted(best)}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
