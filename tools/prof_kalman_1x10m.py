import sys, os
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
N2 = 10_000_000
t2 = np.cumsum(rng.exponential(1 / 30, N2)); y2 = rng.normal(size=N2)
ctx = gp.Context(0)
ctx.set_times(t2); ctx.set_outputs(y2)
for i in range(3):
    ctx.lgssm_logpdf(3, np.log([1.0, 1.0, 0.1])); print(ctx.last_timing())
