"""Scratch GPU check: scaled GPAR objective + q_u vs the oracle; timing at scale."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
import oracle
from oracle import cport
from oracle.dtc import scaled_gpar_objective

rng = np.random.default_rng(0)
ctx = gp.Context(0)
for (N, M, D, kt, ko) in [(30, 10, 1, 3, 3), (500, 20, 2, 3, 3), (4100, 130, 1, 3, 0), (3000, 81, 2, 2, 3), (2500, 33, 3, 1, 2)]:
    t = np.sort(rng.uniform(0, N / 30, N)); X = rng.normal(size=(N, D)); Z = rng.normal(size=(M, D)); y = rng.normal(size=N)
    th = rng.uniform(-1.0, 0.3, 5)
    ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(t); ctx.set_outputs(y)
    v, A = ctx.scaled_dtc(kt, ko, th, return_A=True)
    tl, tv, ol, ov, ns = oracle.unpack_gpar(th)
    Cfu = oracle.pairwise(ko, X, Z, ol, ov ** 2); cu = oracle.pairwise(ko, Z, Z, ol, ov ** 2) + ns ** 2 * np.eye(M)
    v0, A0 = oracle.compute_gpar_dtc_objective(Cfu, cu, t, y, kt, tl, tv ** 2, ns ** 2, dense_logdet=False, decorrelate=cport.kalman_decorrelate)
    params = np.array([tl, tv, ol, ov, ns])
    try:
        m_e, Dinv, U_u = ctx.compute_q_u(kt, ko, params)
        m0, D0, U0 = oracle.compute_q_u(Cfu, oracle.pairwise(ko, Z, Z, ol, ov ** 2), t, y, kt, tl, tv ** 2, ns ** 2, decorrelate=cport.kalman_decorrelate)
        q = (float(np.max(np.abs(m_e - m0)) / np.max(np.abs(m0))), float(np.max(np.abs(Dinv - D0)) / np.max(np.abs(D0))), float(np.max(np.abs(U_u - U0))))
    except gp.PosDefException as e:
        q = "posdef: %s" % e
    print(N, M, D, kt, ko, "dtc rel", abs(v - v0) / abs(v0), "A max abs", float(np.max(np.abs(A - A0))), "q_u", q, ctx.last_timing())

N, M = 1_000_000, 1024
t = np.arange(N) / 30.0
x = rng.uniform(0, 100, N); z = np.linspace(0, 100, M)
y = np.sin(x) + 0.3 * np.cos(0.05 * t) + 0.1 * rng.normal(size=N)
th = np.log(np.array([1.0, 1.0, 1.0, 1.0, 0.1]))
ctx.set_inputs(x); ctx.set_pseudo(z); ctx.set_times(t); ctx.set_outputs(y)
for it in range(3):
    t0 = time.perf_counter(); r = ctx.scaled_dtc(3, 3, th); t1 = time.perf_counter()
    print("scaled N=1M M=1024", r, "wall ms", (t1 - t0) * 1e3, ctx.last_timing())
# oracle at N/32 for cross-check at scale
Ns = N // 32
ctx.set_inputs(x[:Ns]); ctx.set_times(t[:Ns]); ctx.set_outputs(y[:Ns])
r = ctx.scaled_dtc(3, 3, th)
r0 = scaled_gpar_objective(th, x[:Ns, None], z[:, None], t[:Ns], y[:Ns], decorrelate=cport.kalman_decorrelate)
print("scaled N=%d M=1024 gpu %r oracle %r rel %g" % (Ns, r, r0, abs(r - r0) / abs(r0)))
