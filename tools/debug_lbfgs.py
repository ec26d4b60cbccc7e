import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpar_at_scale_b200 as gp
from gpar_at_scale_b200 import chain, data, lbfgs, neldermead, api
ctx = gp.Context(0)
rng = np.random.default_rng(5)
x, y_obs, x_true, y_true = data.generate_big_dataset(rng, data_samples=3000, true_samples=4000)
Y = np.stack(y_obs)
o = 2
X = np.ascontiguousarray(Y[:o].T)
ctx.set_inputs(X); ctx.set_pseudo(chain.strided_pseudo_inputs(X, 40)); ctx.set_times(x); ctx.set_outputs(Y[o])
th0 = np.random.default_rng([1, o, 0]).random(5)
def fg(th):
    try:
        v, g = ctx.scaled_dtc_grad(3, 3, th)
    except api._ffi.PosDefException:
        return np.inf, np.zeros(5)
    return -v, -g
res = lbfgs.optimize(fg, th0, iterations=100, show_trace=True)
print("lbfgs", res.minimum, res.minimizer, res.iterations, res.f_calls, res.converged)
f = lambda th: fg(th)[0]
res2 = neldermead.optimize(f, th0, iterations=400)
print("nm", res2.minimum, res2.minimizer, res2.f_calls)
print("grad at nm optimum", fg(res2.minimizer))
res3 = lbfgs.optimize(fg, res2.minimizer, iterations=100, show_trace=True)
print("lbfgs from nm", res3.minimum, res3.minimizer)
