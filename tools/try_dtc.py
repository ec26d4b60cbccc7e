"""Scratch GPU check: DTC logpdf + grad vs the oracle at small sizes, then timing at N=1M."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
import oracle
from oracle.grad import dtc_diag_value_and_grad

rng = np.random.default_rng(0)
ctx = gp.Context(0)
for (N, M, D, kind, vfe, jit) in [(1000, 50, 1, 3, False, -1.0), (3001, 130, 2, 0, True, 1e-4), (777, 300, 3, 2, False, -1.0),
                                   (5000, 257, 1, 1, True, -1.0)]:
    X = rng.normal(size=(N, D)) * 2; Z = rng.normal(size=(M, D)) * 2; y = rng.normal(size=N)
    th = rng.uniform(-1, 0.5, 3)
    ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y)
    v, g = ctx.dtc_logpdf(kind, th, vfe=vfe, jitter=jit, grad=True)
    v2 = ctx.dtc_logpdf(kind, th, vfe=vfe, jitter=jit, grad=False)
    v0, g0 = dtc_diag_value_and_grad(th, X, Z, y, kind, vfe, jit)
    print(N, M, D, kind, vfe, jit, "val rel", abs(v - v0) / abs(v0), abs(v2 - v0) / abs(v0), "grad rel", np.max(np.abs(g - g0) / (np.abs(g0) + 1e-12)), ctx.last_timing())

N, M = 1_000_000, 1024
x = rng.uniform(0, 100, N); z = np.linspace(x.min(), x.max(), M)
y = np.sin(x) + 0.3 * np.cos(3.1 * x) + 0.1 * rng.normal(size=N)
th = np.log(np.array([1.0, 1.0, 0.1]))
ctx.set_inputs(x); ctx.set_pseudo(z); ctx.set_outputs(y)
for grad in (False, True):
    for it in range(4):
        t0 = time.perf_counter(); r = ctx.dtc_logpdf(3, th, grad=grad); t1 = time.perf_counter()
        print("N=1M M=1024 grad=%s" % grad, r, "wall ms", (t1 - t0) * 1e3, "dev ms / launches", ctx.last_timing())
