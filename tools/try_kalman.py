"""Scratch GPU check: LGSSM logpdf / decorrelate / smooth vs the C oracle, then timing."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
from oracle import cport

rng = np.random.default_rng(0)
ctx = gp.Context(0)
def rel(a, b): return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300)))
for kind in (3, 2, 1):
    for N, batch, irregular, usevec in [(1, 1, False, False), (5, 2, True, False), (31, 1, False, True), (32, 3, True, False), (33, 1, True, True),
                                        (1000, 4, True, True), (40000, 2, True, False), (70001, 1, False, True)]:
        t = np.cumsum(rng.exponential(1 / 30, N)) if irregular else np.arange(N) / 30.0
        Y = rng.normal(size=(batch, N))
        rv = np.where(rng.uniform(size=N) < 0.1, 1e10, 0.3 ** 2) if usevec else None
        th = rng.uniform(-1.5, 0.5, 3)
        l, var, sig = np.exp(th) + 1e-3
        ctx.set_times(t); ctx.set_outputs(Y); ctx.set_noise_vector(rv)
        lml, alpha = ctx.lgssm_decorrelate(kind, th)
        noise = rv if usevec else sig ** 2
        lml0, alpha0 = cport.kalman_filter_batch(kind, t, Y, l, var ** 2, sig ** 2, rvec=rv, want_alpha=True)
        lml_s, mean, v = ctx.lgssm_smooth(kind, th)
        l0s, mean0, v0 = cport.kalman_smooth_batch(kind, t, Y, l, var ** 2, noise)
        # independent models
        ths = rng.uniform(-1.5, 0.5, (batch, 3))
        lmli = ctx.lgssm_logpdf(kind, ths)
        pp = np.exp(ths) + 1e-3
        lmli0 = cport.kalman_filter_batch(kind, t, Y, pp[:, 0], pp[:, 1] ** 2, pp[:, 2] ** 2, rvec=rv)
        print(kind, N, batch, irregular, usevec, "lml", rel(lml, lml0), rel(lml_s, lml0), rel(lmli, lmli0), "alpha", float(np.max(np.abs(alpha - alpha0))),
              "mean", float(np.max(np.abs(mean - mean0))), "var", rel(v, v0))
ctx.set_noise_vector(None)
# timing: 1024 x 10k independent models
B, N = 1024, 10000
t = np.cumsum(rng.exponential(1 / 30, N)); Y = rng.normal(size=(B, N))
ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
ctx.set_times(t); ctx.set_outputs(Y)
for it in range(3):
    t0 = time.perf_counter(); r = ctx.lgssm_logpdf(3, ths); t1 = time.perf_counter()
    ms, nl = ctx.last_timing()
    print("1024x10k logpdf: wall ms", (t1 - t0) * 1e3, "dev ms", ms, "launches", nl, "Gsteps/s", B * N / ms / 1e6)
for it in range(2):
    t0 = time.perf_counter(); r = ctx.lgssm_smooth(3, ths[0]); t1 = time.perf_counter()
    ms, nl = ctx.last_timing()
    print("1024x10k smooth: wall ms", (t1 - t0) * 1e3, "dev ms", ms, "launches", nl, "Gsteps/s", B * N / ms / 1e6)
N = 10_000_000
t = np.arange(N) / 30.0; y = rng.normal(size=N)
ctx.set_times(t); ctx.set_outputs(y)
th = np.log([1.0, 1.0, 0.1])
for it in range(3):
    r = ctx.lgssm_logpdf(3, th); ms, nl = ctx.last_timing()
    print("1x10M logpdf:", r, "dev ms", ms, "launches", nl, "Gsteps/s", N / ms / 1e6)
print("oracle", cport.kalman_logpdf(3, t, y, 1.001, 1.001 ** 2, 0.101 ** 2))
