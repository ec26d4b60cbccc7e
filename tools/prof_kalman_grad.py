import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
ctx = gp.Context(0)
B, NK = 1024, 10000
ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
ctx.set_times(np.cumsum(rng.exponential(1 / 30, NK))); ctx.set_outputs(rng.normal(size=(B, NK)))
for i in range(2):
    ctx.lgssm_logpdf_grad(3, ths); print("grad", ctx.last_timing())
