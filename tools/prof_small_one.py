"""scaled_dtc at the reference's own size (N = 8 496, M = 50) and the EEG shape (N = M = 156), for ncu launch lists."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
from gpar_at_scale_b200 import data
rng = np.random.default_rng(0)
x, y_obs, x_true, y_true = data.generate_big_dataset(rng, true_samples=1000)
ctx = gp.Context(0)
th5 = np.log(np.array([1.0, 1.0, 1.0, 1.0, 0.6]))
X = y_obs[0][:, None]; Z = np.linspace(X.min(), X.max(), 50)[:, None]
ctx.set_times(x); ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y_obs[1])
for i in range(3):
    ctx.scaled_dtc(3, 3, th5); print(ctx.last_timing())
