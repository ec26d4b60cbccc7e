"""Matern state-space (LGSSM) construction, Kalman filter and RTS smoother
(oracle; test infrastructure only).  Pure NumPy, sequential loops: use for small N; the plain-C
twin in ``oracle/c/oracle_kalman.c`` (see ``oracle.cport``) handles large N and the CPU baseline.

Restates TemporalGPs.jl ~0.2 [un-vendored dependency] as used at
``src/gp/temporal_gp_inference.jl:15-39`` (``create_lgssm``: ``to_sde(GP(kernel(k; l, s)), SArrayStorage)``
then ``sde(t, sigma^2 | noise_vector)``), ``src/gp/dtc.jl:101-102,106,115`` (``decorrelate``),
``src/gp/temporal_gp_inference.jl:78`` (``logpdf``) and ``:109`` / ``gpar_scaled_inference.jl:117``
(``smooth``).

Model (state dim d = 1, 2, 3 for Matern-1/2, 3/2, 5/2; lambda = sqrt(2 nu)):
  F = companion matrix of (s + lambda)^d, H = e_1^T, x_0 ~ N(0, s P_inf),
  A_k = exp(F dt_k / l), Q_k = s P_inf - A_k (s P_inf) A_k^T,
  the time vector is prefixed with t_1 - 1, so step 1 also performs a predict (which leaves the
  stationary prior unchanged).
"""
import numpy as np
from .params import MATERN12, MATERN32, MATERN52

LOG2PI = float(np.log(2.0 * np.pi))
SMOOTH_JITTER = 1e-12   # TemporalGPs `smooth`: cholesky(P_pred + 1e-12 I)   [from memory]
INF_NOISE = 1e10        # src/gp/temporal_gp_inference.jl:95, gpar_scaled_inference.jl:102


def sde_matrices(kind):
    """(F, P_inf, H) of the unit-variance, unit-length-scale Matern SDE."""
    if kind == MATERN12:
        return np.array([[-1.0]]), np.array([[1.0]]), np.array([1.0])
    if kind == MATERN32:
        lam = np.sqrt(3.0)
        F = np.array([[0.0, 1.0], [-lam ** 2, -2.0 * lam]])
        return F, np.diag([1.0, 3.0]), np.array([1.0, 0.0])
    if kind == MATERN52:
        lam = np.sqrt(5.0)
        F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3, -3.0 * lam ** 2, -3.0 * lam]])
        kap = 5.0 / 3.0
        P = np.array([[1.0, 0.0, -kap], [0.0, kap, 0.0], [-kap, 0.0, 25.0]])
        return F, P, np.array([1.0, 0.0, 0.0])
    raise ValueError("kernel %r has no finite-dimensional SDE form" % (kind,))


def transition(kind, a):
    """A = exp(F a) in closed form: (F + lambda I) is nilpotent of index d, hence
    A = e^{-lambda a} sum_{j<d} (N a)^j / j!  with N = F + lambda I."""
    F, _, _ = sde_matrices(kind)
    d = F.shape[0]
    lam = {1: 1.0, 2: np.sqrt(3.0), 3: np.sqrt(5.0)}[d]
    Nm = F + lam * np.eye(d)
    A = np.eye(d)
    term = np.eye(d)
    for j in range(1, d):
        term = term @ Nm * (a / j)
        A = A + term
    return np.exp(-lam * a) * A


def build_lgssm(kind, t, l, s, noise):
    """``create_lgssm`` — src/gp/temporal_gp_inference.jl:15-39.  ``noise`` is the scalar
    sigma^2 or a per-step vector R_k (the 1e10 trick).  -> dict(As, Qs, P0, R, H)."""
    t = np.asarray(t, dtype=np.float64)
    n = t.shape[0]
    _, Pinf, H = sde_matrices(kind)
    P0 = s * Pinf
    tt = np.concatenate([[t[0] - 1.0], t])
    dts = np.diff(tt)
    As = np.stack([transition(kind, dt / l) for dt in dts]) if n else np.zeros((0,) + Pinf.shape)
    Qs = np.stack([P0 - A @ P0 @ A.T for A in As]) if n else np.zeros((0,) + Pinf.shape)
    R = np.full(n, float(noise)) if np.ndim(noise) == 0 else np.asarray(noise, dtype=np.float64)
    return dict(As=As, Qs=Qs, P0=P0, R=R, H=H, d=Pinf.shape[0])


def _filter(model, y, store=False):
    As, Qs, R = model["As"], model["Qs"], model["R"]
    d = model["d"]
    n = As.shape[0]
    y = np.asarray(y, dtype=np.float64)
    m = np.zeros(d)
    P = model["P0"].copy()
    alpha = np.zeros(n)
    logS = np.zeros(n)
    if store:
        mf = np.zeros((n, d)); Pf = np.zeros((n, d, d)); mp_ = np.zeros((n, d)); Pp_ = np.zeros((n, d, d))
    for k in range(n):
        A = As[k]
        mp = A @ m
        Pp = A @ P @ A.T + Qs[k]
        S = Pp[0, 0] + R[k]
        sq = np.sqrt(S)
        B = Pp[0, :] / sq                       # U' \ (H Pp)
        a = (y[k] - mp[0]) / sq                 # U' \ (y - H mp)
        m = mp + B * a
        P = Pp - np.outer(B, B)
        alpha[k] = a
        logS[k] = np.log(S)
        if store:
            mf[k], Pf[k], mp_[k], Pp_[k] = m, P, mp, Pp
    lml = -0.5 * (n * LOG2PI + logS.sum() + np.dot(alpha, alpha))
    if store:
        return lml, alpha, mf, Pf, mp_, Pp_
    return lml, alpha


def kalman_decorrelate(kind, t, y, l, s, noise):
    """``decorrelate(lgssm, y)`` -> (lml, alpha) — src/gp/dtc.jl:106,115;
    gpar_scaled_inference.jl:175,183.  alpha = chol(K_time + R)^{-1} y (innovations whitening)."""
    return _filter(build_lgssm(kind, t, l, s, noise), y)


def kalman_logpdf(kind, t, y, l, s, noise):
    """``logpdf(lgssm, y)`` — src/gp/temporal_gp_inference.jl:78."""
    return _filter(build_lgssm(kind, t, l, s, noise), y)[0]


def kalman_smooth(kind, t, y, l, s, noise, full=False):
    """``smooth(lgssm, y)`` -> (lml, mean_k = m^s_k[1], var_k = P^s_k[1,1]) —
    src/gp/temporal_gp_inference.jl:109, gpar_scaled_inference.jl:117 (callers read ``.m[1]``,
    ``.P[1]``).  RTS with G_k^T = (P^-_{k+1} + 1e-12 I)^{-1} A_{k+1} P_k."""
    model = build_lgssm(kind, t, l, s, noise)
    lml, _, mf, Pf, mp, Pp = _filter(model, y, store=True)
    n, d = mf.shape
    ms = mf.copy(); Ps = Pf.copy()
    I = np.eye(d)
    for k in range(n - 2, -1, -1):
        A1 = model["As"][k + 1]
        Gt = np.linalg.solve(Pp[k + 1] + SMOOTH_JITTER * I, A1 @ Pf[k])
        ms[k] = mf[k] + Gt.T @ (ms[k + 1] - mp[k + 1])
        Ps[k] = Pf[k] + Gt.T @ (Ps[k + 1] - Pp[k + 1]) @ Gt
    if full:
        return lml, ms, Ps
    return lml, ms[:, 0].copy(), Ps[:, 0, 0].copy()


def dense_time_cov(kind, t, l, s, noise):
    """``cov(time_prior(t, sigma^2))`` — the dense N x N ``noise_matrix`` of src/gp/dtc.jl:98-99."""
    from .kernels import pairwise
    t = np.asarray(t, dtype=np.float64)
    K = pairwise(kind, t[:, None], t[:, None], l=l, s=s)
    R = np.full(t.shape[0], float(noise)) if np.ndim(noise) == 0 else np.asarray(noise)
    return K + np.diag(R)
