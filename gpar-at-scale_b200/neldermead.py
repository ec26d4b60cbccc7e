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


def optimize(f, x0, iterations=1000, g_tol=1e-8, time_limit=float("nan"), show_trace=False):
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.size
    m = n + 1
    alpha, beta, gamma, delta = 1.0, 1.0 + 2.0 / n, 0.75 - 1.0 / (2.0 * n), 1.0 - 1.0 / n
    simplex = np.tile(x0, (m, 1))
    for i in range(n):
        simplex[i + 1, i] = 1.5 * x0[i] + 0.025          # AffineSimplexer: (1 + b) x_i + a
    fv = np.array([f(v) for v in simplex])
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
        fr = f(xr); calls += 1
        if fr < fv[0]:
            xe = centroid + beta * (xr - centroid)
            fe = f(xe); calls += 1
            if fe < fr:
                simplex[-1], fv[-1] = xe, fe
            else:
                simplex[-1], fv[-1] = xr, fr
        elif fr < fv[-2]:
            simplex[-1], fv[-1] = xr, fr
        else:
            if fr < fv[-1]:       # outside contraction
                xc = centroid + gamma * (xr - centroid)
                fc = f(xc); calls += 1
                ok = fc <= fr
            else:                 # inside contraction
                xc = centroid - gamma * (xr - centroid)
                fc = f(xc); calls += 1
                ok = fc < fv[-1]
            if ok:
                simplex[-1], fv[-1] = xc, fc
            else:                 # shrink towards the best vertex
                for i in range(1, m):
                    simplex[i] = simplex[0] + delta * (simplex[i] - simplex[0])
                    fv[i] = f(simplex[i]); calls += 1
        if show_trace:
            print("%6d   %.6e" % (it, float(np.min(fv))))
    best = int(np.argmin(fv))
    xbest, fbest = simplex[best].copy(), float(fv[best])
    centroid = simplex.mean(axis=0)
    fcen = f(centroid); calls += 1
    if fcen < fbest:
        xbest, fbest = centroid, float(fcen)
    return OptimResult(xbest, fbest, it, calls, converged, stopped_by_time)
