"""Every entry point once on small ragged shapes — run under compute-sanitizer (memcheck) on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp

rng = np.random.default_rng(0)
ctx = gp.Context(0)
for (n, m, d) in [(1, 1, 1), (37, 5, 2), (515, 131, 3), (4099, 257, 1)]:
    X = rng.normal(size=(n, d)); Z = rng.normal(size=(m, d)); y = rng.normal(size=n); t = np.sort(rng.uniform(0, 5, n))
    ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y); ctx.set_times(t)
    th3 = np.array([0.1, -0.2, -1.0]); th5 = np.array([0.1, -0.2, 0.3, 0.1, -1.0])
    for kind in (0, 3):
        ctx.dtc_logpdf(kind, th3, vfe=True, grad=True); ctx.dtc_logpdf(kind, th3)
    ctx.scaled_dtc(3, 3, th5, return_A=True)
    ctx.compute_q_u(2, 3, np.exp(th5) + 1e-3)
    Y = rng.normal(size=(3, n)); ctx.set_outputs(Y)
    rv = np.where(rng.uniform(size=n) < 0.2, 1e10, 0.05)
    for rvec in (None, rv):
        ctx.set_noise_vector(rvec)
        for kind in (1, 2, 3):
            ctx.lgssm_logpdf(kind, rng.uniform(-1, 0, (3, 3))); ctx.lgssm_decorrelate(kind, th3); ctx.lgssm_smooth(kind, th3)
    ctx.set_outputs(y)
    W = rng.normal(size=(m, 5))
    ctx.scaled_predict(3, 3, np.exp(th5) + 1e-3, W)
    ctx.set_noise_vector(None)
    ctx.set_outputs(Y)
    ctx.exact_logpdf(0, 3, th3)
    if d >= 2:
        ctx.exact_logpdf(0, 3, th5); ctx.exact_posterior(0, 3, th5, rng.normal(size=(7, d)))
    ctx.exact_posterior(3, 3, th3, rng.normal(size=(7, d)))
    # round-2 entry points: gradients and batched candidates
    ctx.set_outputs(y)
    ctx.lgssm_logpdf_grad(3, th3); ctx.lgssm_logpdf(3, rng.uniform(-1, 0, (5, 3)))       # candidates on one sequence
    ctx.scaled_dtc_grad(3, 3, th5); ctx.scaled_dtc_batch(3, 3, np.tile(th5, (3, 1)) + 0.01 * rng.normal(size=(3, 5)))
    if n <= 200:
        ctx.exact_logpdf_batch(3, 3, np.tile(th3, (4, 1)))
    print("ok", n, m, d)
print("sanitize driver finished")
