"""Phase timings of the scaled-GPAR objective (and its gradient) at the BASELINE config-5 shape: N = 2 097 152, M = 2048, D inputs."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpar_at_scale_b200 as gp
from gpar_at_scale_b200 import chain
N, M = 2_097_152, 2048
D = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rng = np.random.default_rng(0)
t = np.arange(N) / 30.0
X = rng.normal(size=(N, D)); y = np.sin(X[:, 0]) + 0.5 * np.sin(0.05 * t) + 0.3 * rng.normal(size=N)
ctx = gp.Context(0)
ctx.set_inputs(X); ctx.set_pseudo(chain.strided_pseudo_inputs(X, M)); ctx.set_times(t); ctx.set_outputs(y)
th = np.log([2.0, 0.5, 1.0, 1.0, 0.3])
for it in range(3):
    v = ctx.scaled_dtc(3, 3, th); ms, L = ctx.last_timing(); ph = ctx.last_profile()
    print("value      : %.1f ms device (producers %.1f, syrk %.1f, rest %.1f), %d launches" % (ms, ph[0], ph[1], ph[2], L))
for it in range(3):
    v, g = ctx.scaled_dtc_grad(3, 3, th); ms, L = ctx.last_timing(); ph = ctx.last_profile()
    print("value+grad : %.1f ms device (producers %.1f, syrk %.1f, rest %.1f), %d launches" % (ms, ph[0], ph[1], ph[2], L))
