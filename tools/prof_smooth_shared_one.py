"""One shared-model smoother / filter call on 1024 x 10k (for ncu launch lists)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
ctx = gp.Context(0)
B, NK = 1024, 10000
ctx.set_times(np.cumsum(rng.exponential(1 / 30, NK))); ctx.set_outputs(rng.normal(size=(B, NK)))
th = np.log([1.0, 1.0, 0.1])
for i in range(3):
    ctx.lgssm_smooth(gp.MATERN52, th, keep_on_device=True); print("smooth", ctx.last_timing())
for i in range(3):
    ctx.lgssm_logpdf(gp.MATERN52, th); print("logpdf", ctx.last_timing())
