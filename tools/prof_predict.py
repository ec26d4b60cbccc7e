"""Prediction of one scaled-GPAR output at the reference's big-set size (GPAR_scaled_examples.jl: 8 496 training + 100 000
test points, 81 pseudo-points, 100 Monte-Carlo samples) and at a 10x larger one: device protocol, per-stage wall clock."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpar_at_scale_b200 as gp
from gpar_at_scale_b200 import chain
ctx = gp.Context(0)
rng = np.random.default_rng(0)
for n, ns, m, S in ((8496, 100000, 81, 100), (200000, 1000000, 512, 100)):
    t = np.sort(rng.uniform(0, n / 30, n)); ts = np.sort(rng.uniform(0, n / 30, ns))
    X = rng.normal(size=(n, 2)); Xs = rng.normal(size=(ns, 2)); y = np.sin(X[:, 0]) + 0.3 * rng.normal(size=n)
    Z = chain.strided_pseudo_inputs(X, m)
    params = np.array([1.3, 0.9, 1.1, 0.8, 0.3])
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(t); ctx.set_outputs(y); ctx.set_noise_vector(None)
        ctx.sample_q_u(3, 3, params, 1, S); d1 = ctx.last_timing()[0]
        t1 = time.perf_counter()
        ctx.set_merged(t, y, ts, params[4] ** 2, X=X, Xs=Xs); d2 = ctx.last_timing()[0]
        t2 = time.perf_counter()
        ctx.scaled_predict(3, 3, params, None, keep_on_device=True); d3 = ctx.last_timing()[0]
        t3 = time.perf_counter()
        mean, sd = ctx.take_test()
        t4 = time.perf_counter()
    print("N=%d N*=%d M=%d S=%d: q_u+draws %.1f ms (device %.2f), merge %.1f ms (device %.2f), predict %.1f ms (device %.2f), take %.1f ms, total %.1f ms"
          % (n, ns, m, S, (t1 - t0) * 1e3, d1, (t2 - t1) * 1e3, d2, (t3 - t2) * 1e3, d3, (t4 - t3) * 1e3, (t4 - t0) * 1e3))
