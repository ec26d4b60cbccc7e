"""Prediction protocol of get_sde_predictions at scale: host merge/sort/un-sort (numpy) vs the device protocol
(gpar_set_merged / gpar_take_test).  N training + N* test points, Matern-5/2, wall-clock per stage."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
rng = np.random.default_rng(0)
t = np.sort(rng.uniform(0, n / 30, n)); ts = np.sort(rng.uniform(0, n / 30, ns)); y = np.sin(0.01 * t) + 0.1 * rng.normal(size=n)
th = np.log([1.0, 1.0, 0.1]); sig2 = (np.exp(th[2]) + 1e-3) ** 2
ctx = gp.Context(0)
for rep in range(3):
    t0 = time.perf_counter()
    tc = np.concatenate([t, ts]); perm = np.argsort(tc, kind="stable"); rev = np.argsort(perm, kind="stable")
    st = tc[perm]; sy = np.concatenate([y, np.zeros(ns)])[perm]; sr = np.concatenate([np.full(n, sig2), np.full(ns, 1e10)])[perm]
    t1 = time.perf_counter()
    ctx.set_times(st); ctx.set_outputs(sy); ctx.set_noise_vector(sr)
    t2 = time.perf_counter()
    _, mean, var = ctx.lgssm_smooth(3, th)
    t3 = time.perf_counter()
    m = mean[0][rev][n:]; v = var[0][rev][n:]
    t4 = time.perf_counter()
    print("host   : merge+sort %.1f ms, upload %.1f ms, smooth+D2H %.1f ms (device %.2f ms), un-sort %.1f ms, total %.1f ms"
          % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, ctx.last_timing()[0], (t4 - t3) * 1e3, (t4 - t0) * 1e3))
    t0 = time.perf_counter()
    ctx.set_merged(t, y, ts, sig2)
    t1 = time.perf_counter(); dms = ctx.last_timing()[0]
    ctx.lgssm_smooth(3, th, keep_on_device=True)
    t2 = time.perf_counter()
    a, b = ctx.take_test()
    t3 = time.perf_counter()
    print("device : set_merged %.1f ms (device %.2f ms), smooth %.1f ms, take_test %.1f ms, total %.1f ms   identical=%s"
          % ((t1 - t0) * 1e3, dms, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, np.array_equal(a, m) and np.array_equal(b, v)))
    ctx.set_noise_vector(None)
