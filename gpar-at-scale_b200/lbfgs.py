"""L-BFGS — the gradient-based alternative to the reference's Nelder-Mead loop.

The reference optimises every objective with `Optim.optimize(nlml, params, NelderMead(), ...)`
(src/gp/dtc.jl:58-61; temporal_gp_inference.jl:82) because it has no gradients; the library's
`gpar_scaled_dtc_grad` / `gpar_lgssm_logpdf_grad` / `gpar_dtc_logpdf` entry points supply them, and in
a Julia deployment `Optim.LBFGS()` with an `fg!` that `ccall`s them takes this module's place.  Only
`fg` touches the GPU.  Two-loop recursion (memory 10), backtracking line search with cubic
interpolation on the Armijo condition, curvature-guarded updates; non-finite values (a failed
Cholesky reported as +inf by the caller) shorten the step.
"""
import time
import numpy as np
from .neldermead import OptimResult


def optimize(fg, x0, iterations=100, g_tol=1e-6, f_reltol=1e-10, memory=10, time_limit=float("nan"), show_trace=False):
    """Minimise f.  fg(x) -> (f, grad).  Returns an OptimResult (f_calls counts fg evaluations)."""
    x = np.asarray(x0, dtype=np.float64).copy()
    f, g = fg(x)
    calls = 1
    if not np.isfinite(f):
        return OptimResult(x, float(f), 0, calls, False, False)
    S, Y = [], []
    t0 = time.time()
    it = 0
    converged = False
    stopped_by_time = False
    step0 = 1.0 / max(np.linalg.norm(g), 1.0)
    while it < iterations:
        if np.max(np.abs(g)) <= g_tol:
            converged = True
            break
        if time_limit == time_limit and time.time() - t0 > time_limit:
            stopped_by_time = True
            break
        it += 1
        # two-loop recursion
        q = g.copy()
        alphas = []
        for s, y in zip(reversed(S), reversed(Y)):
            a = (s @ q) / (y @ s)
            alphas.append(a)
            q -= a * y
        if S:
            q *= (S[-1] @ Y[-1]) / (Y[-1] @ Y[-1])
        for (s, y), a in zip(zip(S, Y), reversed(alphas)):
            b = (y @ q) / (y @ s)
            q += (a - b) * s
        d = -q
        dg = d @ g
        if dg >= 0:                      # not a descent direction: restart from steepest descent
            S, Y = [], []
            d = -g
            dg = d @ g
        step = 1.0 if S else step0
        fn, gn = np.inf, None
        for ls in range(25):
            xn = x + step * d
            fn, gn = fg(xn)
            calls += 1
            if np.isfinite(fn) and fn <= f + 1e-4 * step * dg:
                break
            if np.isfinite(fn):          # minimiser of the quadratic through f, dg, fn, safeguarded
                new = -dg * step * step / (2.0 * (fn - f - dg * step))
                step = min(max(new, 0.1 * step), 0.5 * step)
            else:
                step *= 0.25
        else:
            break                        # line search failed: x is (numerically) stationary
        s = xn - x
        y = gn - g
        if s @ y > 1e-12 * np.linalg.norm(s) * np.linalg.norm(y):
            S.append(s); Y.append(y)
            if len(S) > memory:
                S.pop(0); Y.pop(0)
        df = f - fn
        x, f, g = xn, fn, gn
        if show_trace:
            print("%6d   %.10e   |g| %.3e" % (it, f, np.max(np.abs(g))))
        if df <= f_reltol * abs(f):
            converged = True
            break
    return OptimResult(x, float(f), it, calls, converged, stopped_by_time)
