"""Strong scaling of ONE scaled-GPAR objective (compute_gpar_dtc_objective, dtc.jl:83-128) with the rows sharded over the
devices of a group (gpar_group_scaled_dtc_sharded): every member holds the full (t, y) and a row slice of the inputs; one
all-gather of slice summaries and one all-reduce of (G, g) per evaluation.

    python tools/bench_sharded_scaled.py --devices 2 --npoints 2097152 --pseudo 2048 --dim 7
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpar_at_scale_b200 as gp

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=1); ap.add_argument("--npoints", type=int, default=2_097_152)
ap.add_argument("--pseudo", type=int, default=2048); ap.add_argument("--dim", type=int, default=7); ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--no-single", action="store_true", help="skip the one-device evaluation (problems that do not fit one device)")
a = ap.parse_args()
rng = np.random.default_rng(1)
N, M, D = a.npoints, a.pseudo, a.dim
t = np.arange(N) / 30.0
X = np.cumsum(rng.normal(size=(N, D)), axis=0) / np.sqrt(N) * 3 + 0.3 * rng.normal(size=(N, D))
Z = X[:: max(1, N // M)][:M].copy()
y = np.sin(X[:, 0]) + 0.5 * np.sin(0.05 * t) + 0.1 * rng.normal(size=N)
th = np.log([2.0, 0.5, 1.0, 1.0, 0.1])
res = {"metric": "sharded scaled objective ms per evaluation", "unit": "ms", "n_gpus": a.devices, "scaling": "strong",
       "config": {"workload": "one gpar_scaled_dtc, N=%d rows sharded over the devices, M=%d, D=%d, Matern-5/2 time and output kernels" % (N, M, D)}}
if not a.no_single:
    ctx = gp.Context(0)
    ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(t); ctx.set_outputs(y)
    for _ in range(2):
        v1 = ctx.scaled_dtc(3, 3, th)
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter(); v1 = ctx.scaled_dtc(3, 3, th); ts.append(time.perf_counter() - t0)
    res["ms_one_device"] = float(np.median(ts) * 1e3); res["value_one_device"] = repr(v1)
    ctx.close()
g = gp.Group(list(range(a.devices)))
lo = g.load_row_slices(X, Z, t, y)
for _ in range(3):
    v = g.scaled_dtc_sharded(3, 3, th, lo)
ts = []
for _ in range(a.steps):
    t0 = time.perf_counter(); v = g.scaled_dtc_sharded(3, 3, th, lo); ts.append(time.perf_counter() - t0)
res["value"] = float(np.median(ts) * 1e3); res["value_sharded"] = repr(v)
if not a.no_single:
    res["speedup"] = res["ms_one_device"] / res["value"]; res["rel_diff"] = abs(v - v1) / abs(v1)
ts = []
for _ in range(a.steps):
    t0 = time.perf_counter(); vg, gr = g.scaled_dtc_sharded(3, 3, th, lo, grad=True); ts.append(time.perf_counter() - t0)
res["ms_sharded_value_and_grad"] = float(np.median(ts[1:]) * 1e3); res["grad_sharded"] = gr.tolist()
if not a.no_single:
    ctx = gp.Context(0)
    ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(t); ctx.set_outputs(y)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); v1g, g1 = ctx.scaled_dtc_grad(3, 3, th); ts.append(time.perf_counter() - t0)
    res["ms_one_device_value_and_grad"] = float(np.median(ts[1:]) * 1e3)
    res["grad_rel_diff"] = float(np.max(np.abs(gr - g1)) / np.max(np.abs(g1)))
    res["grad_speedup"] = res["ms_one_device_value_and_grad"] / res["ms_sharded_value_and_grad"]
    ctx.close()
print(json.dumps(res))
g.close()
