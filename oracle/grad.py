"""Gradient oracle (test infrastructure only): the oracle objectives restated in torch float64 so
that autograd supplies d/d theta.  The reference has NO gradients (Zygote is imported at
src/gp/optimized.jl:3 and temporal_gp_inference.jl:4 but never called; every optimiser is
NelderMead), so gradient parity is pinned on autograd of these functions, cross-checked by central
differences of the NumPy oracle in tests/."""
import math
import torch

LOG2PI = math.log(2.0 * math.pi)


def _unpack(theta):
    return torch.exp(theta) + 1e-3


def _kern(kind, d2):
    d2 = torch.clamp(d2, min=0.0)
    if kind == 0:
        return torch.exp(-0.5 * d2)
    # sqrt has an infinite derivative at 0: guard it (the kernels themselves are C^1 there)
    r = torch.sqrt(d2 + 1e-300)
    if kind == 1:
        return torch.exp(-r)
    if kind == 2:
        a = math.sqrt(3.0) * r
        return (1.0 + a) * torch.exp(-a)
    a = math.sqrt(5.0) * r
    return (1.0 + a + a * a / 3.0) * torch.exp(-a)


def pairwise_t(kind, X, Z, l, s):
    d2 = ((X[:, None, :] - Z[None, :, :]) ** 2).sum(-1) / (l * l)
    return s * _kern(kind, d2)


def dtc_diag_t(theta, X, Z, y, kind, vfe=False, jitter=-1.0):
    """torch twin of oracle.dtc.dtc_diag / elbo_diag with theta = (log l, log var, log sigma)
    (unpack_gp, src/util.jl:36-43); jitter < 0 -> sigma^2 (dtc.jl:35)."""
    l, var, sig = _unpack(theta)
    s = var * var
    nv = sig * sig
    j = nv if jitter < 0 else torch.as_tensor(jitter, dtype=torch.float64)
    Cfu = pairwise_t(kind, X, Z, l, s)
    Kuu = pairwise_t(kind, Z, Z, l, s) + j * torch.eye(Z.shape[0], dtype=torch.float64)
    Lu = torch.linalg.cholesky(Kuu)
    A = torch.linalg.solve_triangular(Lu, Cfu.T, upper=False) / sig
    m = A.shape[0]
    Ll = torch.linalg.cholesky(A @ A.T + torch.eye(m, dtype=torch.float64))
    delta = y / sig
    c = torch.linalg.solve_triangular(Ll, (A @ delta)[:, None], upper=False)[:, 0]
    n = y.shape[0]
    tmp = n * torch.log(nv) + 2.0 * torch.log(torch.diagonal(Ll)).sum() + delta @ delta - c @ c
    val = -(n * LOG2PI + tmp) / 2.0
    if vfe:
        val = val - 0.5 * (n * s / nv - (A * A).sum())
    return val


def dtc_diag_value_and_grad(theta, X, Z, y, kind, vfe=False, jitter=-1.0, wrt_Z=False):
    """-> (value, d/dtheta) or, with wrt_Z, (value, d/dtheta, d/dZ (M x D)) — pseudo-input gradients."""
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    X = torch.as_tensor(X, dtype=torch.float64)
    Z = torch.tensor(Z, dtype=torch.float64)
    if X.ndim == 1:
        X = X[:, None]
    if Z.ndim == 1:
        Z = Z[:, None]
    Z = Z.clone().requires_grad_(wrt_Z)
    v = dtc_diag_t(th, X, Z, torch.as_tensor(y, dtype=torch.float64), kind, vfe, jitter)
    v.backward()
    if wrt_Z:
        return float(v.detach()), th.grad.numpy().copy(), Z.grad.numpy().copy()
    return float(v.detach()), th.grad.numpy().copy()


# ---- scaled-GPAR objective (src/gp/dtc.jl:83-128) in torch, for autograd gradients -------------------
def _sde_torch(kind):
    import numpy as np
    from .lgssm import sde_matrices
    F, Pinf, H = sde_matrices(kind)
    lam = {1: 1.0, 2: math.sqrt(3.0), 3: math.sqrt(5.0)}[F.shape[0]]
    return (torch.as_tensor(F), torch.as_tensor(Pinf), lam)


def _transition_t(F, lam, a):
    d = F.shape[0]
    Nm = F + lam * torch.eye(d, dtype=torch.float64)
    A = torch.eye(d, dtype=torch.float64)
    term = torch.eye(d, dtype=torch.float64)
    for j in range(1, d):
        term = term @ Nm * (a / j)
        A = A + term
    return torch.exp(-lam * a) * A


def kalman_decorrelate_t(kind, t, V, l, s, noise):
    """Differentiable sequential Kalman `decorrelate` of the columns of V (n x c) sharing one model
    (torch twin of oracle.lgssm._filter; small n only).  -> (sum log S, alpha (n x c))."""
    F, Pinf, lam = _sde_torch(kind)
    d = F.shape[0]
    n = V.shape[0]
    P0 = s * Pinf
    P = P0
    M = torch.zeros(d, V.shape[1], dtype=torch.float64)
    alphas = []
    logS = torch.zeros((), dtype=torch.float64)
    tprev = t[0] - 1.0
    for k in range(n):
        A = _transition_t(F, lam, (t[k] - tprev) / l)
        tprev = t[k]
        Q = P0 - A @ P0 @ A.T
        Mp = A @ M
        Pp = A @ P @ A.T + Q
        S = Pp[0, 0] + (noise[k] if getattr(noise, 'ndim', 0) == 1 else noise)
        sq = torch.sqrt(S)
        B = Pp[0, :] / sq
        a = (V[k] - Mp[0]) / sq
        M = Mp + B[:, None] * a[None, :]
        P = Pp - torch.outer(B, B)
        alphas.append(a)
        logS = logS + torch.log(S)
    return logS, torch.stack(alphas)


def lgssm_logpdf_value_and_grad(theta, t, y, kind=3, rvec=None):
    """logpdf(lgssm, y) (temporal_gp_inference.jl:78) and its autograd gradient w.r.t. the raw theta;
    sequential torch filter — small n only.  rvec: optional per-step noise (then sigma is unused)."""
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    p = _unpack(th)
    t = torch.as_tensor(t, dtype=torch.float64)
    y = torch.as_tensor(y, dtype=torch.float64)
    if rvec is None:
        logS, a = kalman_decorrelate_t(kind, t, y[:, None], p[0], p[1] * p[1], p[2] * p[2])
    else:
        logS, a = kalman_decorrelate_t(kind, t, y[:, None], p[0], p[1] * p[1], torch.as_tensor(rvec, dtype=torch.float64))
    v = -(y.shape[0] * LOG2PI + logS + (a * a).sum()) / 2.0
    v.backward()
    g = th.grad.numpy().copy() if th.grad is not None else None
    return float(v.detach()), g


def scaled_dtc_t(theta, X, Z, t, y, k_time, k_out):
    """torch twin of oracle.dtc.scaled_gpar_objective (dtc.jl:29-47 + :83-128, O(N) log-determinant)."""
    p = _unpack(theta)
    time_l, time_var, out_l, out_var, sig = p[0], p[1], p[2], p[3], p[4]
    nv = sig * sig
    Cfu = pairwise_t(k_out, X, Z, out_l, out_var * out_var)
    cov_u = pairwise_t(k_out, Z, Z, out_l, out_var * out_var) + nv * torch.eye(Z.shape[0], dtype=torch.float64)
    logS, W = kalman_decorrelate_t(k_time, t, torch.cat([y[:, None], Cfu], dim=1), time_l, time_var * time_var, nv)
    alpha, beta = W[:, 0], W[:, 1:]
    Lu = torch.linalg.cholesky(cov_u)
    A = torch.linalg.solve_triangular(Lu, beta.T, upper=False)
    m = A.shape[0]
    Ll = torch.linalg.cholesky(A @ A.T + torch.eye(m, dtype=torch.float64))
    c = torch.linalg.solve_triangular(Ll, (A @ alpha)[:, None], upper=False)[:, 0]
    n = y.shape[0]
    return -(n * LOG2PI + logS + 2.0 * torch.log(torch.diagonal(Ll)).sum() + alpha @ alpha - c @ c) / 2.0


def scaled_dtc_value_and_grad(theta, X, Z, t, y, k_time=3, k_out=3):
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    X = torch.as_tensor(X, dtype=torch.float64); Z = torch.as_tensor(Z, dtype=torch.float64)
    if X.ndim == 1:
        X = X[:, None]
    if Z.ndim == 1:
        Z = Z[:, None]
    v = scaled_dtc_t(th, X, Z, torch.as_tensor(t, dtype=torch.float64), torch.as_tensor(y, dtype=torch.float64), k_time, k_out)
    v.backward()
    return float(v.detach()), th.grad.numpy().copy()
