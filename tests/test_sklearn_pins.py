"""Independent third-party pins of the oracle (CPU): scikit-learn's Gaussian-process kernels / regressor and SciPy's
multivariate normal — code written by neither the reference's authors nor this repository — against rows K (kernels),
E (exact log-pdf), T (exact posterior) and D/Y (state-space log-pdf) of SURVEY 8a.  The reference's own numbers can
only come from julia/make_reference_golden.jl (tests/test_reference_golden.py); until that file exists these pins are
what stands between the oracle and "from memory" for the Stheno kernel conventions
(kernel(k; l, s) = s k(r / l); Matern-1/2, 3/2, 5/2, EQ)."""
import numpy as np
import pytest
from scipy.linalg import cho_factor, cho_solve
from scipy.stats import multivariate_normal
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
import oracle
from oracle import EQ, MATERN12, MATERN32, MATERN52


def sk_kernel(kind, l, s):
    base = RBF(length_scale=l) if kind == EQ else Matern(length_scale=l, nu={MATERN12: 0.5, MATERN32: 1.5, MATERN52: 2.5}[kind])
    return ConstantKernel(constant_value=s) * base


@pytest.mark.parametrize("kind", [EQ, MATERN12, MATERN32, MATERN52])
@pytest.mark.parametrize("D", [1, 2, 3, 7])
def test_kernel_matrices_match_sklearn(kind, D):
    rng = np.random.default_rng(100 + D)
    X = rng.normal(size=(40, D)); Z = rng.normal(size=(9, D))
    l, s = 1.37, 0.8 ** 2
    ours = oracle.pairwise(kind, X, Z, l=l, s=s)
    theirs = sk_kernel(kind, l, s)(X, Z)
    assert np.max(np.abs(ours - theirs)) <= 1e-13 * s


@pytest.mark.parametrize("kind", [EQ, MATERN32, MATERN52])
def test_exact_logpdf_and_posterior_match_sklearn_gpr(kind):
    """optimized.jl:28-36 (log-pdf) and :94 + marginals (posterior mean / std, noise-free latent as Stheno's post(x*))."""
    rng = np.random.default_rng(7)
    n, D = 60, 2
    X = rng.normal(size=(n, D)); y = np.sin(X[:, 0]) + 0.3 * X[:, 1] + 0.1 * rng.normal(size=n)
    Xs = rng.normal(size=(25, D))
    l, var, sig = oracle.unpack_gp(np.array([0.2, -0.1, -1.5]))
    gpr = GaussianProcessRegressor(kernel=sk_kernel(kind, l, var ** 2), alpha=sig ** 2, optimizer=None, normalize_y=False).fit(X, y)
    K = oracle.pairwise(kind, X, X, l=l, s=var ** 2)
    ours = oracle.exact_logpdf(K, sig ** 2, y)
    theirs = gpr.log_marginal_likelihood_value_
    assert abs(ours - theirs) <= 1e-10 * abs(theirs)
    mean, var_post = oracle.exact_posterior(K, oracle.pairwise(kind, Xs, X, l=l, s=var ** 2), np.full(len(Xs), var ** 2), sig ** 2, y)
    m_sk, sd_sk = gpr.predict(Xs, return_std=True)
    assert np.max(np.abs(mean - m_sk)) <= 1e-9 * max(1.0, np.max(np.abs(m_sk)))
    assert np.max(np.abs(np.sqrt(var_post) - sd_sk)) <= 1e-7 * np.max(sd_sk)      # sklearn clips / recomputes the variance less stably


def test_gpar_kernel_composition_matches_sklearn_pieces():
    """optimized.jl:132-144: time_var^2 k_t(|dx_1| / time_l) + out_var^2 k_o(||dx_2:D|| / out_l), built from sklearn's kernels
    on the masked columns."""
    rng = np.random.default_rng(3)
    X = rng.normal(size=(30, 4)); Z = rng.normal(size=(11, 4))
    tl, tv, ol, ov, _ = oracle.unpack_gpar(np.array([0.3, -0.2, 0.1, 0.2, -1.0]))
    for kt, ko in ((MATERN52, MATERN52), (EQ, EQ), (MATERN52, EQ), (MATERN12, MATERN32)):
        ours = oracle.gpar_kernel_matrix(kt, ko, X, Z, tl, tv, ol, ov)
        theirs = sk_kernel(kt, tl, tv ** 2)(X[:, :1], Z[:, :1]) + sk_kernel(ko, ol, ov ** 2)(X[:, 1:], Z[:, 1:])
        assert np.max(np.abs(ours - theirs)) <= 1e-13 * (tv ** 2 + ov ** 2)


@pytest.mark.parametrize("kind", [MATERN12, MATERN32, MATERN52])
def test_state_space_logpdf_matches_scipy_mvn_with_sklearn_kernel(kind):
    """temporal_gp_inference.jl:15-39,78: the LGSSM log-pdf equals log N(y; 0, K_time + R) with K_time from sklearn's
    Matern kernel and the density from SciPy — both the scalar-noise model and the 1e10 noise-vector trick (:93-97)."""
    rng = np.random.default_rng(11)
    n = 120
    t = np.cumsum(rng.exponential(1 / 30, n)); t[40] = t[39]
    y = np.sin(3 * t) + 0.2 * rng.normal(size=n)
    l, var, sig = oracle.unpack_gp(np.array([-0.4, 0.1, -1.2]))
    K = sk_kernel(kind, l, var ** 2)(t[:, None], t[:, None])
    ours = oracle.kalman_logpdf(kind, t, y, l, var ** 2, sig ** 2)
    theirs = multivariate_normal(mean=np.zeros(n), cov=K + sig ** 2 * np.eye(n), allow_singular=False).logpdf(y)
    assert abs(ours - theirs) <= 1e-9 * abs(theirs)
    rv = np.full(n, sig ** 2); rv[rng.choice(n, 12, replace=False)] = 1e10
    ours = oracle.kalman_logpdf(kind, t, y, l, var ** 2, rv)
    # (SciPy's MVN rejects / truncates a covariance whose eigenvalues span 1e-1 .. 1e10; LAPACK's Cholesky does not)
    c, low = cho_factor(K + np.diag(rv), lower=True)
    theirs = -0.5 * (n * np.log(2 * np.pi) + 2 * np.log(np.diag(c)).sum() + y @ cho_solve((c, low), y))
    assert abs(ours - theirs) <= 1e-8 * abs(theirs)
