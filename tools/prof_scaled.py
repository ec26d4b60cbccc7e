"""Small driver for ncu: scaled-GPAR objective at N = 1M, M = 1024 (two evaluations)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(0)
N, M = 1_000_000, 1024
t = np.arange(N) / 30.0
x = rng.uniform(0, 100, N); z = np.linspace(0, 100, M)
y = np.sin(x) + 0.3 * np.cos(0.05 * t) + 0.1 * rng.normal(size=N)
th = np.log(np.array([1.0, 1.0, 1.0, 1.0, 0.1]))
ctx = gp.Context(0)
ctx.set_inputs(x); ctx.set_pseudo(z); ctx.set_times(t); ctx.set_outputs(y)
for i in range(3):
    r = ctx.scaled_dtc(3, 3, th)
    print("scaled", r, ctx.last_timing(), ctx.last_profile())
