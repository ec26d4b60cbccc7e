"""Pseudo-point (DTC / VFE) objectives and q(u) (oracle; test infrastructure only)."""
import numpy as np
from scipy.linalg import solve_triangular
from .kernels import pairwise
from .lgssm import kalman_decorrelate, dense_time_cov
from .params import unpack_gpar, MATERN52

LOG2PI = float(np.log(2.0 * np.pi))


def _chol(a):
    return np.linalg.cholesky(0.5 * (a + a.T))


def dtc_dense(Cfu, cov_u, Sigma_y, y):
    """Stheno's dense DTC as restated in examples/dtc_example.jl:10-23
    (``_compute_intermediates``).  -> (dtc, A) with A = L_u^{-1} (L_y^{-1} Cfu)^T."""
    Ly = _chol(Sigma_y)
    Lu = _chol(cov_u)
    A = solve_triangular(Lu, solve_triangular(Ly, Cfu, lower=True).T, lower=True)
    m = A.shape[0]
    Ll = _chol(A @ A.T + np.eye(m))
    delta = solve_triangular(Ly, y, lower=True)
    tmp = (2.0 * np.log(np.diag(Ly)).sum() + 2.0 * np.log(np.diag(Ll)).sum() + delta @ delta
           - np.sum(solve_triangular(Ll, A @ delta, lower=True) ** 2))
    return float(-(y.shape[0] * LOG2PI + tmp) / 2.0), A


def compute_gpar_dtc_objective(Cfu, cov_u, t, y, k_time, time_l, time_s, noise_var,
                               dense_logdet=True, decorrelate=kalman_decorrelate):
    """``compute_gpar_dtc_objective`` — src/gp/dtc.jl:83-128, line by line.
    Cfu = cov(f,u) (:104), cov_u = cov(u) = Kuu + noise_sigma^2 I (:119, dtc.jl:35),
    alpha/beta by M+1 Kalman ``decorrelate`` passes (:106-117),
    A = chol(cov_u).U' \\ beta' (:119), Lambda = chol(A A' + I) (:120), and
    dtc = -(N log 2pi + logdet(noise_matrix) + logdet Lambda + sum alpha^2 - ||Lambda.U' \\ A alpha||^2)/2
    (:122-125).  ``dense_logdet`` keeps the literal dense N x N ``logdet(noise_matrix)`` (:99,123);
    False uses the identical O(N) value logdet Sigma_y = -2 lml - N log 2pi - sum alpha^2 from the
    filter's own lml (the first return of ``decorrelate`` that :106 discards).  -> (dtc, A)."""
    y = np.asarray(y, dtype=np.float64)
    n, m = Cfu.shape
    lml, alpha = decorrelate(k_time, t, y, time_l, time_s, noise_var)
    beta = np.zeros((n, m))
    for col in range(m):
        _, beta[:, col] = decorrelate(k_time, t, np.ascontiguousarray(Cfu[:, col]), time_l, time_s, noise_var)
    if dense_logdet:
        Ln = _chol(dense_time_cov(k_time, t, time_l, time_s, noise_var))
        logdet_noise = 2.0 * np.log(np.diag(Ln)).sum()
    else:
        logdet_noise = -2.0 * lml - n * LOG2PI - alpha @ alpha
    Lu = _chol(cov_u)
    A = solve_triangular(Lu, beta.T, lower=True)
    Ll = _chol(A @ A.T + np.eye(m))
    tmp = (logdet_noise + 2.0 * np.log(np.diag(Ll)).sum() + alpha @ alpha
           - np.sum(solve_triangular(Ll, A @ alpha, lower=True) ** 2))
    return float(-(n * LOG2PI + tmp) / 2.0), A


def gpar_dtc_collapsed(G, g, aa, logdet_noise, n, cov_u):
    """The same objective from the collapsed statistics only (no N x M array):
    G = beta^T beta, g = beta^T alpha, aa = alpha^T alpha; Lambda = I + L_u^{-1} G L_u^{-T};
    c = L_Lambda^{-1} L_u^{-1} g.  Used to check the GPU path's intermediate results."""
    Lu = _chol(cov_u)
    B = solve_triangular(Lu, solve_triangular(Lu, G, lower=True).T, lower=True)
    Ll = _chol(B + np.eye(G.shape[0]))
    c = solve_triangular(Ll, solve_triangular(Lu, g, lower=True), lower=True)
    return float(-(n * LOG2PI + logdet_noise + 2.0 * np.log(np.diag(Ll)).sum() + aa - c @ c) / 2.0)


def scaled_gpar_objective(theta, X, Z, t, y, k_out=MATERN52, k_time=MATERN52, dense_logdet=False,
                          decorrelate=kalman_decorrelate):
    """The ``nlml`` closure body of src/gp/dtc.jl:29-47 (returns +dtc, the closure returns -dtc)."""
    time_l, time_var, out_l, out_var, noise_sigma = unpack_gpar(theta)
    Cfu = pairwise(k_out, X, Z, l=out_l, s=out_var ** 2)
    cov_u = pairwise(k_out, Z, Z, l=out_l, s=out_var ** 2) + noise_sigma ** 2 * np.eye(len(Z))
    return compute_gpar_dtc_objective(Cfu, cov_u, t, y, k_time, time_l, time_var ** 2,
                                      noise_sigma ** 2, dense_logdet=dense_logdet,
                                      decorrelate=decorrelate)[0]


def dtc_diag(Cfu, cov_u, noise_var, y):
    """Plain DTC with Sigma_y = sigma^2 I (Stheno ``dtc(f(x, sigma^2), y, u)``;
    examples/dtc_example.jl:10-23 specialised to a diagonal noise matrix).
    -> (dtc, trace(A A^T))."""
    y = np.asarray(y, dtype=np.float64)
    n, m = Cfu.shape
    sig = np.sqrt(noise_var)
    Lu = _chol(cov_u)
    A = solve_triangular(Lu, Cfu.T, lower=True) / sig
    Ll = _chol(A @ A.T + np.eye(m))
    delta = y / sig
    tmp = (n * np.log(noise_var) + 2.0 * np.log(np.diag(Ll)).sum() + delta @ delta
           - np.sum(solve_triangular(Ll, A @ delta, lower=True) ** 2))
    return float(-(n * LOG2PI + tmp) / 2.0), float(np.sum(A * A))


def elbo_diag(Cfu, cov_u, kff_diag_sum, noise_var, y):
    """Titsias VFE bound, Stheno ``elbo``: dtc - 1/2 (tr(Sigma_y^{-1} K_ff) - ||A||_F^2)
    [un-vendored dependency, from memory; never called by the reference]."""
    d, tr_aat = dtc_diag(Cfu, cov_u, noise_var, y)
    return float(d - 0.5 * (kff_diag_sum / noise_var - tr_aat))


def compute_q_u(Cfu, Cuu, t, y, k_time, time_l, time_s, noise_var, decorrelate=kalman_decorrelate):
    """``compute_q_u`` — src/gp/gpar_scaled_inference.jl:141-196.  Bare Cuu (no jitter, :157-159),
    raw outputs (:183).  -> (m_e, inv(D), U_u) with q_u = MvNormal(m_e, inv(D))."""
    n, m = Cfu.shape
    Lu = _chol(Cuu)
    beta = np.zeros((n, m))
    for col in range(m):
        _, beta[:, col] = decorrelate(k_time, t, np.ascontiguousarray(Cfu[:, col]), time_l, time_s, noise_var)
    B = solve_triangular(Lu, beta.T, lower=True)
    _, b_y = decorrelate(k_time, t, np.asarray(y, dtype=np.float64), time_l, time_s, noise_var)
    D = B @ B.T + np.eye(m)
    Ld = _chol(D)
    rhs = B @ b_y
    m_e = solve_triangular(Ld.T, solve_triangular(Ld, rhs, lower=True), lower=False)
    Dinv = np.linalg.inv(0.5 * (D + D.T))
    return m_e, 0.5 * (Dinv + Dinv.T), Lu.T


def _l_dkdl(kind, d2):
    """(kappa, l dkappa/dl) of the base kernels as functions of d2 = r^2 (already length-scaled)."""
    if kind == 0:
        k = np.exp(-0.5 * d2); return k, d2 * k
    if kind == 1:
        r = np.sqrt(d2); k = np.exp(-r); return k, r * k
    if kind == 2:
        a = np.sqrt(3.0 * d2); e = np.exp(-a); return (1.0 + a) * e, a * a * e
    a = np.sqrt(5.0 * d2); e = np.exp(-a)
    return (1.0 + a + a * a / 3.0) * e, a * a * (1.0 + a) * e / 3.0


def dtc_diag_value_and_grad_np(theta, X, Z, y, kind, vfe=False, jitter=-1.0):
    """BLAS-backed CPU port of one logpdf+grad evaluation (the CPU baseline of bench.py): the same
    collapsed-statistics algebra as DESIGN.md (G = Kuf Kfu by dsyrk/dgemm, forward-mode dG/dlog l,
    analytic M x M adjoints), float64 NumPy/SciPy on all host threads.  Pinned against torch
    autograd of dtc_diag (oracle/grad.py) in tests/test_oracle.py.  theta = (log l, log var,
    log sigma) as unpack_gp (src/util.jl:36-43); jitter < 0 -> sigma^2 (dtc.jl:35)."""
    from .params import unpack_gp
    X = np.asarray(X, dtype=np.float64); Z = np.asarray(Z, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    if Z.ndim == 1:
        Z = Z[:, None]
    l, v, sg = unpack_gp(theta)
    s = v * v; nz = sg * sg; j = nz if jitter < 0 else jitter
    N, M = len(y), len(Z)
    d2 = np.zeros((N, M))
    for c in range(X.shape[1]):
        df = X[:, c][:, None] - Z[:, c][None, :]
        d2 += df * df
    d2 /= l * l
    k, dk = _l_dkdl(kind, d2)
    K = s * k; dK = s * dk
    G = K.T @ K; H = K.T @ dK; g = K.T @ y; h = dK.T @ y; yy = y @ y
    d2u = np.zeros((M, M))
    for c in range(Z.shape[1]):
        df = Z[:, c][:, None] - Z[:, c][None, :]
        d2u += df * df
    ku, dku = _l_dkdl(kind, d2u / (l * l))
    Kj = s * ku + j * np.eye(M); dKu = s * dku
    ip = 1.0 / nz
    Lu = _chol(Kj)
    B = solve_triangular(Lu, solve_triangular(Lu, G, lower=True).T, lower=True) * ip
    trB = np.trace(B)
    Ll = _chol(B + np.eye(M))
    c = solve_triangular(Ll, solve_triangular(Lu, g, lower=True), lower=True) * ip
    val = -0.5 * (N * LOG2PI + N * np.log(nz) + 2.0 * np.log(np.diag(Ll)).sum() + yy * ip - c @ c)
    if vfe:
        val += -0.5 * (N * s * ip - trB)
    V = solve_triangular(Lu, np.eye(M), lower=True); Kinv = V.T @ V
    R = solve_triangular(Ll, V, lower=True); P = R.T @ R
    w = nz * solve_triangular(Lu.T, solve_triangular(Ll.T, c, lower=False), lower=False)
    ip2 = ip * ip
    trTH = (P * H).sum() + w @ H @ w * ip2; trTG = (P * G).sum() + w @ G @ w * ip2
    trTKdK = (P * dKu).sum() + w @ dKu @ w * ip2 - (Kinv * dKu).sum()
    trTKK = (P * Kj).sum() + w @ Kj @ w * ip2 - (Kinv * Kj).sum()
    trTK = np.trace(P) + w @ w * ip2 - np.trace(Kinv)
    gw = g @ w; hw = h @ w
    dlogl = -ip * trTH + ip2 * hw - 0.5 * trTKdK
    ds = (-ip * trTG + ip2 * gw - 0.5 * (trTKK - j * trTK)) / s
    dn = -0.5 * (N * ip - yy * ip2 + 2.0 * gw * ip2 * ip - trTG * ip2)
    if jitter < 0:
        dn += -0.5 * trTK
    if vfe:
        C = Kinv @ G @ Kinv; trC = np.trace(C)
        dlogl += -0.5 * ip * (-2.0 * (Kinv * H).sum() + (C * dKu).sum())
        ds += -0.5 * (N * ip - (trB + j * ip * trC) / s)
        dn += -0.5 * (-N * s * ip2 + trB * ip)
        if jitter < 0:
            dn += -0.5 * ip * trC
    grad = np.array([dlogl * (l - 1e-3) / l, ds * 2.0 * v * (v - 1e-3), dn * 2.0 * sg * (sg - 1e-3)])
    return float(val), grad
