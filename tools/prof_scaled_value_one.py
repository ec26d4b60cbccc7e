"""A few gpar_scaled_dtc evaluations at N = 1M, M = 1024 (for ncu launch lists)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpar_at_scale_b200 as gp
N, M = 1_000_000, 1024
rng = np.random.default_rng(0)
t = np.arange(N) / 30.0
x = rng.uniform(0, 100, N); z = np.linspace(0, 100, M)
y = np.sin(x) + 0.5 * np.sin(0.05 * t) + 0.1 * rng.normal(size=N)
th = np.log([2.0, 0.5, 1.0, 1.0, 0.1])
ctx = gp.Context(0)
ctx.set_inputs(x); ctx.set_pseudo(z); ctx.set_times(t); ctx.set_outputs(y)
for it in range(3):
    v = ctx.scaled_dtc(3, 3, th); print(v, ctx.last_timing(), ctx.last_profile())
