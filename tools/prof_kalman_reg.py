import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
B, N = 1024, 10000
Y = rng.normal(size=(B, N))
ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
ctx = gp.Context(0)
ctx.set_times_range(0.0, 1 / 30, N); ctx.set_outputs(Y)
for i in range(3):
    ctx.lgssm_logpdf(3, ths)
    print("filter ms", ctx.last_timing())
