"""One configuration of the single-pass steady-state Kalman path (for ncu): python tools/prof_ss3_one.py <variant> <batch> [decorrelate]"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
variant, batch = sys.argv[1], int(sys.argv[2])
os.environ["GPAR_SS3_VARIANT"] = variant; os.environ["GPAR_FILTER_SHARED"] = "0"
rng = np.random.default_rng(2)
ctx = gp.Context(0)
N = 10_000_000
ctx.set_outputs(rng.normal(size=(batch, N))); ctx.set_times_range(0.0, 1 / 30, N)
ths = np.tile(np.log([1.0, 1.0, 0.1]), (batch, 1))
for i in range(3):
    v = ctx.lgssm_logpdf(3, ths)
    print(v[0], ctx.last_timing())
if len(sys.argv) > 3:
    ctx.lgssm_decorrelate(3, ths[0]); print(ctx.last_timing())
