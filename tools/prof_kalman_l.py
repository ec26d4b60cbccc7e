import sys, os
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
B, N = 1024, 10000
t = np.cumsum(rng.exponential(1 / 30, N)); Y = rng.normal(size=(B, N))
ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
ctx = gp.Context(0)
ctx.set_times(t); ctx.set_outputs(Y)
ts = []
for i in range(6):
    ctx.lgssm_logpdf(3, ths); ts.append(ctx.last_timing()[0])
ss = []
for i in range(3):
    ctx.lgssm_smooth(3, ths[0]); ss.append(ctx.last_timing()[0])
N2 = 10_000_000
t2 = np.cumsum(rng.exponential(1 / 30, N2)); y2 = rng.normal(size=N2)
ctx.set_times(t2); ctx.set_outputs(y2)
t1 = []
for i in range(4):
    ctx.lgssm_logpdf(3, ths[0]); t1.append(ctx.last_timing()[0])
print("L=%s  1024x10k filter %.3f ms  smooth %.3f ms   1x10M filter %.3f ms" % (os.environ.get("GPAR_KF_L", "auto"), min(ts[2:]), min(ss[1:]), min(t1[1:])))
