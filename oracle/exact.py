"""Exact (dense) GP log-marginal likelihood and posterior (oracle; test infrastructure only)."""
import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular

LOG2PI = float(np.log(2.0 * np.pi))


def exact_logpdf(K, noise_var, y):
    """``logpdf(f(x, sigma^2), y)`` with zero mean — src/gp/optimized.jl:34,152:
    -1/2 [N log 2pi + logdet(K + sigma^2 I) + y^T (K + sigma^2 I)^{-1} y]."""
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0]
    L = np.linalg.cholesky(K + noise_var * np.eye(n))
    w = solve_triangular(L, y, lower=True)
    return float(-0.5 * (n * LOG2PI + 2.0 * np.log(np.diag(L)).sum() + w @ w))


def exact_posterior(Kff, Ksf, kss_diag, noise_var, y, obs_noise=1e-18):
    """``gp | (gp(x, sigma^2) <- y)`` then ``marginals(post(x*))`` — src/gp/optimized.jl:94,236;
    examples/eeg.jl:185-208.  mean = K*f (K+sigma^2 I)^{-1} y;
    var = k** - diag(K*f (K+sigma^2 I)^{-1} Kf*) + 1e-18 (Stheno's default FiniteGP noise of
    ``post(x*)``) [from memory].  -> (mean, var)."""
    n = Kff.shape[0]
    L = np.linalg.cholesky(Kff + noise_var * np.eye(n))
    w = solve_triangular(L, np.asarray(y, dtype=np.float64), lower=True)
    V = solve_triangular(L, Ksf.T, lower=True)          # N x N*
    mean = V.T @ w
    var = kss_diag - np.einsum("ij,ij->j", V, V) + obs_noise
    return mean, var
