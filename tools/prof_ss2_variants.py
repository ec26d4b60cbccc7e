import sys, os
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
ctx = gp.Context(0)
N = 10_000_000
ctx.set_outputs(rng.normal(size=N)); ctx.set_times_range(0.0, 1 / 30, N)
th = np.log([1.0, 1.0, 0.1])
ms = []
for i in range(6):
    v = ctx.lgssm_logpdf(3, th); ms.append(ctx.last_timing()[0])
print(os.environ.get("GPAR_B200_LIB", "default")[-12:], "lml %.9f  best %.1f us  median %.1f us" % (v[0], min(ms) * 1e3, float(np.median(ms[2:])) * 1e3))
