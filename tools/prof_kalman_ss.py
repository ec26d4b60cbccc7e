import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
ctx = gp.Context(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "10m"
if mode == "10m":
    N = 10_000_000
    ctx.set_outputs(rng.normal(size=N)); ctx.set_times_range(0.0, 1 / 30, N)
    th = np.log([1.0, 1.0, 0.1])
else:
    B, N = 1024, 10000
    ctx.set_outputs(rng.normal(size=(B, N))); ctx.set_times_range(0.0, 1 / 30, N)
    th = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
for i in range(3):
    ctx.lgssm_logpdf(3, th)
    print("filter ms", ctx.last_timing())
