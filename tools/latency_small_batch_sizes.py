import sys, os, time
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
sys.path.insert(0, 'examples')
import toy_data as data
rng = np.random.default_rng(0)
x, y_obs, x_true, y_true = data.generate_big_dataset(rng, true_samples=1000)
ctx = gp.Context(0)
th5 = np.log(np.array([1.0, 1.0, 1.0, 1.0, 0.6]))
def bench(f, n=200):
    for _ in range(5): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
X = y_obs[0][:, None]; Z = np.linspace(X.min(), X.max(), 50)[:, None]
ctx.set_times(x); ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y_obs[1])
for B in (1, 2, 6, 12, 24, 64, 128):
    ths = np.tile(th5, (B, 1)) + 0.05 * rng.normal(size=(B, 5))
    print("N=8496 M=50: %3d candidates fused: %.3f ms per call" % (B, bench(lambda: ctx.scaled_dtc_batch(3, 3, ths), n=50)))
print("single path: %.3f ms" % bench(lambda: ctx.scaled_dtc(3, 3, th5)))
X = np.stack(y_obs[:2], axis=1); d1 = np.linspace(X[:, 0].min(), X[:, 0].max(), 9); d2 = np.linspace(X[:, 1].min(), X[:, 1].max(), 9)
Z = np.array([[a, b] for b in d2 for a in d1])
ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y_obs[2])
for B in (1, 6, 64):
    ths = np.tile(th5, (B, 1)) + 0.05 * rng.normal(size=(B, 5))
    print("N=8496 M=81 D=2: %3d candidates fused: %.3f ms per call" % (B, bench(lambda: ctx.scaled_dtc_batch(3, 3, ths), n=50)))
print("single path: %.3f ms" % bench(lambda: ctx.scaled_dtc(3, 3, th5)))
