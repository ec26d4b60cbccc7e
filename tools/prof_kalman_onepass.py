"""One call each of the one-pass log-pdf on 1 x 10M (irregular grid) and cfg 3 (1024 x 10k, own model each), for ncu."""
import sys, os
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
ctx = gp.Context(0)
N2 = 10_000_000
ctx.set_times(np.cumsum(rng.exponential(1 / 30, N2))); ctx.set_outputs(rng.normal(size=N2))
for i in range(2):
    ctx.lgssm_logpdf(3, np.log([1.0, 1.0, 0.1])); print("1x10M", ctx.last_timing())
B, NK = 1024, 10000
ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
ctx.set_times(np.cumsum(rng.exponential(1 / 30, NK))); ctx.set_outputs(rng.normal(size=(B, NK)))
for i in range(2):
    ctx.lgssm_logpdf(3, ths); print("cfg3", ctx.last_timing())
