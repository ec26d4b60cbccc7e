"""Accuracy of gpar_scaled_dtc_grad against torch autograd of the oracle as the problem becomes ill conditioned
(output-kernel variance growing: cov(u) + G spans more and more orders of magnitude)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpar_at_scale_b200 as gp
from gpar_at_scale_b200 import chain
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
import toy_data as data
from oracle.grad import scaled_dtc_value_and_grad
rng = np.random.default_rng(5)
x, y_obs, x_true, y_true = data.generate_big_dataset(rng, data_samples=3000, true_samples=4000)
Y = np.stack(y_obs); o = 2
X = np.ascontiguousarray(Y[:o].T); Z = chain.strided_pseudo_inputs(X, 40)
ctx = gp.Context(0)
ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(x); ctx.set_outputs(Y[o])
base = np.array([4.6535755, 2.31191692, 4.10235784, 9.21555909, -0.43283099])
for th3 in (0.0, 2.0, 4.0, 6.0, 8.0, 9.2155):
    th = base.copy(); th[3] = th3
    v, g = ctx.scaled_dtc_grad(3, 3, th)
    v0, g0 = scaled_dtc_value_and_grad(th, X, Z, x, Y[o], 3, 3)
    print("theta3 = %.2f  value rel.err %.1e   grad device %s   autograd %s   max rel.err %.1e"
          % (th3, abs(v - v0) / abs(v0), np.array2string(g, precision=4), np.array2string(g0, precision=4), np.max(np.abs(g - g0)) / np.max(np.abs(g0))))
