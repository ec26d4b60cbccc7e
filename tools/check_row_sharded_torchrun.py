"""torchrun --nproc-per-node N tools/check_row_sharded_torchrun.py: ONE scaled objective (value and gradient) with its rows
sharded over the ranks — one process per GPU, torch.distributed/NCCL owns the collectives (parallel.scaled_dtc_row_sharded) —
against the same evaluation on rank 0's device alone.  Prints one JSON line on rank 0."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import gpar_at_scale_b200 as gp
from gpar_at_scale_b200 import parallel

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_097_152
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
D = int(sys.argv[3]) if len(sys.argv) > 3 else 7
rng = np.random.default_rng(1)
t = np.arange(N) / 30.0
X = np.cumsum(rng.normal(size=(N, D)), axis=0) / np.sqrt(N) * 3 + 0.3 * rng.normal(size=(N, D))
Z = X[:: max(1, N // M)][:M].copy()
y = np.sin(X[:, 0]) + 0.5 * np.sin(0.05 * t) + 0.1 * rng.normal(size=N)
th = np.log([2.0, 0.5, 1.0, 1.0, 0.1])
ctx = gp.Context(local)
b = parallel.row_slice_bounds(N, world)
ctx.set_times(t); ctx.set_outputs(y); ctx.set_pseudo(Z); ctx.set_inputs(np.ascontiguousarray(X[b[rank]:b[rank + 1]])); ctx.set_noise_vector(None)
dev = torch.device("cuda", local)

def timed(fn, n=5, skip=2):
    ts = []
    for _ in range(n):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    tt = torch.tensor([float(np.median(ts[skip:]))], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return r, float(tt.item()) * 1e3

v, ms_v = timed(lambda: parallel.scaled_dtc_row_sharded(ctx, 3, 3, th, b[rank], device=dev))
(vg, g5), ms_g = timed(lambda: parallel.scaled_dtc_row_sharded(ctx, 3, 3, th, b[rank], grad=True, device=dev), 4, 1)
out = {"metric": "row-sharded scaled objective, one process per GPU (torch.distributed/NCCL)", "n_gpus": world, "N": N, "M": M, "D": D,
       "ms_value": ms_v, "ms_value_and_grad": ms_g, "value": repr(v), "grad": [float(x) for x in g5]}
if rank == 0:
    ctx.set_inputs(X)
    t0 = time.perf_counter(); v1 = ctx.scaled_dtc(3, 3, th); v1 = ctx.scaled_dtc(3, 3, th); t1 = time.perf_counter()
    vg1, g1 = ctx.scaled_dtc_grad(3, 3, th)
    t2 = time.perf_counter(); vg1, g1 = ctx.scaled_dtc_grad(3, 3, th); t3 = time.perf_counter()
    out.update({"ms_value_one_device": (t1 - t0) / 2 * 1e3, "ms_value_and_grad_one_device": (t3 - t2) * 1e3,
                "value_rel_diff": abs(v - v1) / abs(v1), "grad_rel_diff": float(np.max(np.abs(g5 - g1)) / np.max(np.abs(g1)))})
    out["speedup_value"] = out["ms_value_one_device"] / ms_v; out["speedup_value_and_grad"] = out["ms_value_and_grad_one_device"] / ms_g
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
