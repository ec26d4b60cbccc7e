"""Prediction protocols: merge / sort / 1e10-noise / smooth / un-sort
(oracle; test infrastructure only)."""
import numpy as np
from scipy.linalg import solve_triangular
from .kernels import pairwise
from .lgssm import kalman_smooth, INF_NOISE


def merge_sort(t_train, t_test):
    """src/gp/temporal_gp_inference.jl:55-66 / gpar_scaled_inference.jl:75-87: concat, sortperm
    (Julia's default sort is stable), inverse permutation."""
    tc = np.concatenate([np.asarray(t_train, float), np.asarray(t_test, float)])
    perm = np.argsort(tc, kind="stable")
    rev = np.argsort(perm, kind="stable")
    return tc, perm, rev


def sde_predictions(kind, t_train, y_train, t_test, l, var, noise_sigma, smooth=kalman_smooth):
    """The post-optimisation part of ``get_sde_predictions`` —
    src/gp/temporal_gp_inference.jl:93-113: noise vector sigma^2 / 1e10, sorted, smooth, un-sort,
    keep the test entries.  -> (mean = .m[1], var = .P[1]) at t_test, in the caller's order."""
    ntr = len(t_train)
    tc, perm, rev = merge_sort(t_train, t_test)
    yc = np.concatenate([np.asarray(y_train, float), np.zeros(len(t_test))])
    rc = np.concatenate([np.full(ntr, noise_sigma ** 2), np.full(len(t_test), INF_NOISE)])
    _, mean, v = smooth(kind, tc[perm], yc[perm], l, var ** 2, rc[perm])
    return mean[rev][ntr:], v[rev][ntr:]


def gpar_scaled_predict_given_eps(k_out, k_time, X, Z, t, y, t_star, X_star, params, m_e, U_u, eps,
                                  smooth=kalman_smooth):
    """Deterministic core of ``get_gpar_scaled_predictions`` —
    src/gp/gpar_scaled_inference.jl:74-135 — for given draws ``eps[j] ~ q_u`` (the reference draws
    them from Julia's unseeded global RNG, :94, so only this part is comparable):
    fx_j = Cf*u (U_u \\ eps_j) (:91-97); f*_j = fx_j + smooth(y* - fx_j).m[1] (:113-121);
    sample mean and corrected std over j (:125); un-sort, keep test entries (:132-133)."""
    time_l, time_var, out_l, out_var, noise_sigma = params
    ntr = len(t)
    tc, perm, rev = merge_sort(t, t_star)
    Xc = np.concatenate([X, X_star], axis=0)[perm]
    yc = np.concatenate([np.asarray(y, float), np.zeros(len(t_star))])[perm]
    rc = np.concatenate([np.full(ntr, noise_sigma ** 2), np.full(len(t_star), INF_NOISE)])[perm]
    Cfu_star = pairwise(k_out, Xc, Z, l=out_l, s=out_var ** 2)
    acc = []
    for e in eps:
        fx = Cfu_star @ solve_triangular(U_u, e, lower=False)
        _, sm, _ = smooth(k_time, tc[perm], yc - fx, time_l, time_var ** 2, rc)
        acc.append(fx + sm)
    acc = np.stack(acc)
    mean = acc.mean(axis=0)
    std = acc.std(axis=0, ddof=1) if len(eps) > 1 else np.zeros_like(mean)
    return mean[rev][ntr:], std[rev][ntr:]
