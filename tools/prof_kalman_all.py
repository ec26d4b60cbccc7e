"""Kalman timings (device ms via gpar_last_timing): 1024 x 10k independent Matern-5/2 models and one
10M-step sequence, irregular grid vs regular grid (gpar_set_times_range), filter / gradient / smoother."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
ctx = gp.Context(0)


def best(fn, n=4):
    ms = []
    for _ in range(n):
        fn(); ms.append(ctx.last_timing()[0])
    return min(ms)


B, N = 1024, 10000
Y = rng.normal(size=(B, N))
ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
ctx.set_outputs(Y)
for name, setter in (("irregular", lambda: ctx.set_times(np.cumsum(rng.exponential(1 / 30, N)))), ("regular", lambda: ctx.set_times_range(0.0, 1 / 30, N))):
    setter()
    f = best(lambda: ctx.lgssm_logpdf(3, ths)); g = best(lambda: ctx.lgssm_logpdf_grad(3, ths)); s = best(lambda: ctx.lgssm_smooth(3, ths[0]), 2)
    fs = best(lambda: ctx.lgssm_logpdf(3, ths[0]))
    print("1024x10k %-9s filter %.3f ms (%.1f Gsteps/s)  logpdf+grad %.3f ms  shared theta: filter %.3f ms (%.1f Gsteps/s), smoother %.3f ms" % (name, f, B * N / f / 1e6, g, fs, B * N / fs / 1e6, s))
N2 = 10_000_000
y2 = rng.normal(size=N2); th = np.log([1.0, 1.0, 0.1])
ctx.set_outputs(y2)
for name, setter in (("irregular", lambda: ctx.set_times(np.cumsum(rng.exponential(1 / 30, N2)))), ("regular", lambda: ctx.set_times_range(0.0, 1 / 30, N2))):
    setter()
    f = best(lambda: ctx.lgssm_logpdf(3, th)); g = best(lambda: ctx.lgssm_logpdf_grad(3, th)); d = best(lambda: ctx.lgssm_decorrelate(3, th), 2)
    s = best(lambda: ctx.lgssm_smooth(3, th), 2)
    print("1x10M    %-9s filter %.3f ms (%.1f Gsteps/s)  logpdf+grad %.3f ms  decorrelate %.3f ms  smoother %.3f ms" % (name, f, N2 / f / 1e6, g, d, s))
