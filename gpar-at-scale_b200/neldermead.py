"""Nelder-Mead simplex optimiser — host control loop above the hot path.

The reference calls `Optim.optimize(nlml, params, NelderMead(), Optim.Options(time_limit=...))`
(src/gp/dtc.jl:58-61; optimized.jl:45,164; temporal_gp_inference.jl:82).  In a Julia deployment
Optim.jl itself keeps doing this (the ccall boundary sits inside `nlml`); this module restates
Optim.jl's algorithm [un-vendored dependency, defaults from its documentation: AffineSimplexer
(a = 0.025, b = 0.5), AdaptiveParameters (alpha = 1, beta = 1 + 2/n, gamma = 0.75 - 1/2n,
delta = 1 - 1/n), g_tol = 1e-8 on sqrt(var(f) n/(n+1))..., 1000 iterations] for the Python host
mirror.  Only `f` touches the GPU.
"""
import time
import numpy as np


class OptimResult:
    def __init__(self, minimizer, minimum, iterations, f_calls, converged, stopped_by_time):
        self.minimizer = minimizer
        self.minimum = minimum
        self.iterations = iterations
        self.f_calls = f_calls
        self.converged = converged
        self.stopped_by_time = stopped_by_time


def _simplex_steps(x0, iterations, g_tol, time_limit, show_trace, speculative=False):
    """The algorithm as a coroutine: yields the list of points it needs evaluated next, receives their values, and
    finally returns the OptimResult.  `optimize` drives it with a scalar objective; `optimize_batch` drives several
    instances in lockstep so that one BATCHED objective call (gpar_exact_logpdf_batch, gpar_lgssm_logpdf with one
    parameter set per candidate) serves all of them — simplex vertices x restarts evaluated concurrently.
    speculative: every iteration asks for its reflection, expansion and both contraction points AT ONCE (they depend only
    on the centroid and the worst vertex) and then takes exactly the decisions of the sequential algorithm with the values
    it would have requested — same simplices, same `f_calls` (the count of the sequential algorithm); the extra
    evaluations ride in the same batched call, which at small problem sizes costs no more than a single evaluation."""
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.size
    m = n + 1
    alpha, beta, gamma, delta = 1.0, 1.0 + 2.0 / n, 0.75 - 1.0 / (2.0 * n), 1.0 - 1.0 / n
    simplex = np.tile(x0, (m, 1))
    for i in range(n):
        simplex[i + 1, i] = 1.5 * x0[i] + 0.025          # AffineSimplexer: (1 + b) x_i + a
    fv = np.array((yield [v.copy() for v in simplex]), dtype=np.float64)
    calls = m
    t0 = time.time()
    it = 0
    converged = False
    stopped_by_time = False
    while it < iterations:
        order = np.argsort(fv, kind="stable")
        simplex, fv = simplex[order], fv[order]
        if np.sqrt(np.var(fv, ddof=1) * (n / m)) <= g_tol:
            converged = True
            break
        if time_limit == time_limit and time.time() - t0 > time_limit:
            stopped_by_time = True
            break
        it += 1
        centroid = simplex[:-1].mean(axis=0)
        xr = centroid + alpha * (centroid - simplex[-1])
        spec = None
        if speculative:
            spec = (yield [xr, centroid + beta * (xr - centroid), centroid + gamma * (xr - centroid), centroid - gamma * (xr - centroid)])
            fr = spec[0]; calls += 1
        else:
            fr = (yield [xr])[0]; calls += 1
        if fr < fv[0]:
            xe = centroid + beta * (xr - centroid)
            fe = spec[1] if speculative else (yield [xe])[0]
            calls += 1
            if fe < fr:
                simplex[-1], fv[-1] = xe, fe
            else:
                simplex[-1], fv[-1] = xr, fr
        elif fr < fv[-2]:
            simplex[-1], fv[-1] = xr, fr
        else:
            if fr < fv[-1]:       # outside contraction
                xc = centroid + gamma * (xr - centroid)
                fc = spec[2] if speculative else (yield [xc])[0]
                calls += 1
                ok = fc <= fr
            else:                 # inside contraction
                xc = centroid - gamma * (xr - centroid)
                fc = spec[3] if speculative else (yield [xc])[0]
                calls += 1
                ok = fc < fv[-1]
            if ok:
                simplex[-1], fv[-1] = xc, fc
            else:                 # shrink towards the best vertex: the n new vertices are one batch
                for i in range(1, m):
                    simplex[i] = simplex[0] + delta * (simplex[i] - simplex[0])
                fv[1:] = (yield [simplex[i].copy() for i in range(1, m)]); calls += m - 1
        if show_trace:
            print("%6d   %.6e" % (it, float(np.min(fv))))
    best = int(np.argmin(fv))
    xbest, fbest = simplex[best].copy(), float(fv[best])
    centroid = simplex.mean(axis=0)
    fcen = (yield [centroid])[0]; calls += 1
    if fcen < fbest:
        xbest, fbest = centroid, float(fcen)
    return OptimResult(xbest, fbest, it, calls, converged, stopped_by_time)


def optimize(f, x0, iterations=1000, g_tol=1e-8, time_limit=float("nan"), show_trace=False):
    gen = _simplex_steps(x0, iterations, g_tol, time_limit, show_trace)
    pts = next(gen)
    while True:
        try:
            pts = gen.send([f(p) for p in pts])
        except StopIteration as e:
            return e.value


def optimize_speculative(fbatch, x0, iterations=1000, g_tol=1e-8, time_limit=float("nan"), show_trace=False):
    """ONE Nelder-Mead run whose evaluations go through a batched objective `fbatch(points (K, n)) -> values (K,)`: the
    initial simplex and the shrink steps as one batch each, and the four candidate points of an iteration speculatively
    in one call (see _simplex_steps).  Same minimiser, minimum, iterations and f_calls as `optimize`."""
    gen = _simplex_steps(x0, iterations, g_tol, time_limit, show_trace, speculative=True)
    pts = next(gen)
    while True:
        try:
            pts = gen.send(list(np.asarray(fbatch(np.array(pts, dtype=np.float64)), dtype=np.float64)))
        except StopIteration as e:
            return e.value


def optimize_batch(fbatch, X0, iterations=1000, g_tol=1e-8):
    """len(X0) independent Nelder-Mead runs (restarts) in lockstep: every round, the points all runs need next go to ONE
    call `fbatch(points (K, n)) -> values (K,)`.  Each run performs exactly the operations `optimize` would (same minimum,
    minimiser, iteration and call counts); only the evaluation is shared.  -> list of OptimResult."""
    gens = [_simplex_steps(x, iterations, g_tol, float("nan"), False) for x in X0]
    pending = {k: next(g) for k, g in enumerate(gens)}
    results = [None] * len(gens)
    while pending:
        keys = sorted(pending)
        pts = np.array([p for k in keys for p in pending[k]], dtype=np.float64)
        vals = np.asarray(fbatch(pts), dtype=np.float64)
        pos = 0
        nxt = {}
        for k in keys:
            cnt = len(pending[k])
            try:
                nxt[k] = gens[k].send(list(vals[pos:pos + cnt]))
            except StopIteration as e:
                results[k] = e.value
            pos += cnt
        pending = nxt
    return results
