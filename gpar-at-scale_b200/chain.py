"""The GPAR chain at scale: fit every output's conditional GP (independent tasks, sharded over the
GPUs) and predict down the chain (sequential over outputs).

Mirrors the call pattern of examples/GPAR_scaled_examples.jl:86-175 (`big_synthetic_dataset`):
output 1 is a time-only state-space GP (`get_sde_predictions`, :102-111); output i > 1 is a scaled
GPAR on the OBSERVED earlier outputs (`get_gpar_scaled_predictions([y1], ...)`, `([y1, y2], ...)`,
:132-175), predicted with the earlier outputs' predicted means as inputs (:172 `[test_y1, y2_out]`).
Hyper-parameter restarts draw theta0 ~ U(0,1)^p like the reference's missing-parameter rule
(src/util.jl:128-134).
"""
import time
import numpy as np
from . import api, lbfgs, neldermead, parallel
from .context import Context


def strided_pseudo_inputs(X, M):
    """Pseudo-inputs = strided subsample of the training inputs (the EEG example reuses training
    inputs as pseudo-inputs, examples/eeg.jl:217-220)."""
    n = X.shape[0]
    idx = np.linspace(0, n - 1, min(M, n)).round().astype(np.int64)
    return np.ascontiguousarray(X[np.unique(idx)])


def make_tasks(n_outputs, n_restarts):
    tasks = [(o, r) for o in range(n_outputs) for r in range(n_restarts)]
    costs = [0.02 if o == 0 else 1.0 + 0.03 * o for o, r in tasks]   # output 1 is LGSSM-only
    return tasks, costs


def fit_chain(t, Y, M, n_restarts=1, iterations=200, seed=0, ctx=None, time_kernel=None, out_kernel=None, verbose=False,
              optimizer="neldermead", batched=None):
    """Fits all outputs.  Y: (P, N) observed outputs on the sorted time grid t.  Tasks (output,
    restart) are partitioned over the ranks; every rank returns the gathered best parameters:
    {output: (nlml, theta, restart)}, plus timing info.  optimizer: "neldermead" (the reference's,
    dtc.jl:58-61) or "lbfgs" (uses the library's analytic gradients; `iterations` then bounds the
    L-BFGS iterations).  batched (default: automatic — Nelder-Mead with several restarts at sizes the fused
    small-problem path covers, M <= 160 and N <= 2^18): the restarts of an output run in lock-step and every round of
    candidates is ONE gpar_scaled_dtc_batch / gpar_lgssm_logpdf call; tasks are then whole outputs."""
    time_kernel = time_kernel or api.Matern52(); out_kernel = out_kernel or api.Matern52()
    rank, world = parallel.dist_info()
    ctx = ctx or api.default_context()
    P, N = Y.shape
    tasks, costs = make_tasks(P, n_restarts)
    evals = [0]

    def run_task(task):
        o, r = task
        rng = np.random.default_rng([seed, o, r])
        if o == 0:
            ctx.set_times(t); ctx.set_outputs(Y[0]); ctx.set_noise_vector(None)
            f = lambda th: -ctx.lgssm_logpdf(time_kernel.code, th)[0]                         # temporal_gp_inference.jl:69-79
            th0 = rng.random(3)
            if optimizer == "lbfgs":
                def fg(th):
                    v, g = ctx.lgssm_logpdf_grad(time_kernel.code, th)
                    return -v[0], -g[0]
                res = lbfgs.optimize(fg, th0, iterations=iterations)
            else:
                res = neldermead.optimize(f, th0, iterations=iterations)
            out = np.full(5, np.nan); out[:3] = res.minimizer
            evals[0] += res.f_calls
            return res.minimum, out
        X = np.ascontiguousarray(Y[:o].T)                                                     # observed earlier outputs as inputs
        ctx.set_inputs(X); ctx.set_pseudo(strided_pseudo_inputs(X, M)); ctx.set_times(t); ctx.set_outputs(Y[o])

        def f(th):                                                                            # dtc.jl:29-48
            try:
                return -ctx.scaled_dtc(time_kernel.code, out_kernel.code, th)
            except api._ffi.PosDefException:
                return np.inf
        if optimizer == "lbfgs":
            def fg(th):
                try:
                    v, g = ctx.scaled_dtc_grad(time_kernel.code, out_kernel.code, th)
                except api._ffi.PosDefException:
                    return np.inf, np.zeros(5)
                return -v, -g
            res = lbfgs.optimize(fg, rng.random(5), iterations=iterations)
        else:
            res = neldermead.optimize(f, rng.random(5), iterations=iterations)
        evals[0] += res.f_calls
        return res.minimum, res.minimizer

    if batched is None:
        batched = optimizer == "neldermead" and n_restarts > 1 and M <= 160 and N <= (1 << 18)

    def run_output(o):
        """All restarts of output o in lock-step (same starting points as the separate runs) -> list of OptimResult."""
        if o == 0:
            ctx.set_times(t); ctx.set_outputs(Y[0]); ctx.set_noise_vector(None)
            X0 = np.stack([np.random.default_rng([seed, 0, r]).random(3) for r in range(n_restarts)])
            return neldermead.optimize_batch(lambda P_: -ctx.lgssm_logpdf(time_kernel.code, P_), X0, iterations=iterations)
        X = np.ascontiguousarray(Y[:o].T)
        ctx.set_inputs(X); ctx.set_pseudo(strided_pseudo_inputs(X, M)); ctx.set_times(t); ctx.set_outputs(Y[o])
        X0 = np.stack([np.random.default_rng([seed, o, r]).random(5) for r in range(n_restarts)])

        def fb(P_):
            vals_, codes_ = ctx.scaled_dtc_batch(time_kernel.code, out_kernel.code, P_)
            return np.where(codes_ == 0, -vals_, np.inf)
        return neldermead.optimize_batch(fb, X0, iterations=iterations)

    t0 = time.perf_counter()
    device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            device = torch.device("cuda", ctx.device)
    timing = {}
    if batched:
        out_costs = [0.02 if o == 0 else 1.0 + 0.03 * o for o in range(P)]
        local = {}
        tb = time.perf_counter()
        for o in parallel.shard_tasks(out_costs, world, rank):
            for r, res in enumerate(run_output(o)):
                th = np.full(5, np.nan); th[:len(res.minimizer)] = res.minimizer
                local[o * n_restarts + r] = (res.minimum, th)
                evals[0] += res.f_calls
        timing["busy_seconds"] = time.perf_counter() - tb
        vals, thetas = parallel.gather_results(local, len(tasks), 5, device)
    else:
        vals, thetas = parallel.fit_tasks(tasks, costs, run_task, 5, device, timing)
    dt = time.perf_counter() - t0
    best = parallel.best_per_output(tasks, vals, thetas)
    if verbose and rank == 0:
        for o in sorted(best):
            print("output %d: nlml %.6g theta %s (restart %d)" % (o, best[o][0], np.round(best[o][1], 4), best[o][2]))
    return best, {"seconds": dt, "objective_evals_this_rank": evals[0], "tasks": len(tasks), "world": world,
                  "busy_seconds": timing.get("busy_seconds", dt), "minimum": vals, "minimizer": thetas}


def predict_chain(t, Y, t_star, M, best, nsamples=100, seed=0, ctx=None, time_kernel=None, out_kernel=None):
    """Sequential prediction down the chain at t_star -> (means (P, N*), spreads (P, N*)).
    Output 1: smoothed state-space GP; output i: get_gpar_scaled_predictions with the predicted means
    of outputs < i as inference inputs.  spreads: .P[1] (a variance) for output 1, MC std otherwise —
    exactly what the reference's drivers plot (GPAR_scaled_examples.jl:128-129,162)."""
    time_kernel = time_kernel or api.Matern52(); out_kernel = out_kernel or api.Matern52()
    ctx = ctx or api.default_context()
    P, N = Y.shape
    means = np.zeros((P, len(t_star))); spreads = np.zeros((P, len(t_star)))
    th0 = best[0][1][:3]
    l, var, sig = api.unpack_gp(th0)
    tc = np.concatenate([t, t_star]); perm = np.argsort(tc, kind="stable"); rev = np.argsort(perm, kind="stable")
    ctx.set_times(tc[perm]); ctx.set_outputs(np.concatenate([Y[0], np.zeros(len(t_star))])[perm])
    ctx.set_noise_vector(np.concatenate([np.full(N, sig ** 2), np.full(len(t_star), 1e10)])[perm])
    _, m, v = ctx.lgssm_smooth(time_kernel.code, th0)
    ctx.set_noise_vector(None)
    means[0] = m[0][rev][N:]; spreads[0] = v[0][rev][N:]
    for o in range(1, P):
        X = np.ascontiguousarray(Y[:o].T)
        Xs = np.ascontiguousarray(means[:o].T)
        mo, so = api.get_gpar_scaled_predictions(X, strided_pseudo_inputs(X, M), t, Y[o], t_star, Xs, out_kernel_structure=out_kernel,
                                                 time_kernel_structure=time_kernel, ctx=ctx, seed=(int(seed) << 8) + o,
                                                 nsamples=nsamples, opt_params=api.unpack_gpar(best[o][1]))
        means[o] = mo; spreads[o] = so
    return means, spreads
