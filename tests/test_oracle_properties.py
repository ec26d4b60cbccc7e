"""CPU: size-independent properties of the oracle (hypothesis-driven) — the same invariants the GPU tests use at full size,
here on the checker itself: permutation invariance of the pseudo-point objectives, the joint scale identity of the covariance
blocks, linearity of the smoother in the data, VFE <= DTC, and the filter's chain rule over a split sequence."""
import numpy as np
from hypothesis import given, settings, strategies as st
import oracle
from oracle import cport
from oracle.dtc import dtc_diag, elbo_diag, scaled_gpar_objective

KINDS = st.sampled_from([oracle.EQ, oracle.MATERN12, oracle.MATERN32, oracle.MATERN52])
SS_KINDS = st.sampled_from([oracle.MATERN12, oracle.MATERN32, oracle.MATERN52])


def problem(seed, n, m, d):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)); Z = rng.normal(size=(m, d))
    t = np.cumsum(rng.exponential(1 / 30, n))
    y = np.sin(3 * t) + 0.3 * X[:, 0] + 0.1 * rng.normal(size=n)
    return rng, X, Z, t, y


@settings(max_examples=12, deadline=None)
@given(seed=st.integers(0, 10_000), kind=KINDS, n=st.integers(5, 60), m=st.integers(1, 12), d=st.integers(1, 4))
def test_plain_dtc_is_invariant_under_row_permutations_and_vfe_is_below_it(seed, kind, n, m, d):
    rng, X, Z, _, y = problem(seed, n, m, d)
    l, s, nv = 1.3, 0.8, 0.05
    Cfu = oracle.pairwise(kind, X, Z, l=l, s=s); cov_u = oracle.pairwise(kind, Z, Z, l=l, s=s) + nv * np.eye(m)
    v, _ = dtc_diag(Cfu, cov_u, nv, y)
    p = rng.permutation(n)
    vp, _ = dtc_diag(Cfu[p], cov_u, nv, y[p])
    assert abs(v - vp) <= 1e-10 * max(1.0, abs(v))
    e = elbo_diag(Cfu, cov_u, n * s, nv, y)
    assert e <= v + 1e-9 * max(1.0, abs(v))                # the trace penalty of the ELBO is non-negative


@settings(max_examples=8, deadline=None)
@given(seed=st.integers(0, 10_000), kt=SS_KINDS, ko=KINDS, n=st.integers(8, 50), m=st.integers(1, 8), c=st.floats(0.3, 3.0))
def test_scaled_objective_scale_identity(seed, kt, ko, n, m, c):
    """Every covariance block of the scaled-GPAR model is jointly linear in (time_s, out_s, sigma^2):
    scaling all three by c is the model of sqrt(c) y, i.e. F(c Sigma; y) = F(Sigma; y / sqrt(c)) - (n / 2) log c."""
    _, X, Z, t, y = problem(seed, n, m, 2)
    pos = np.array([0.7, 0.9, 1.1, 0.8, 0.3])                # (time_l, time_var, out_l, out_var, sigma) AFTER unpacking
    th = lambda q: np.log(q - 1e-3)
    f1 = scaled_gpar_objective(th(pos), X, Z, t, y, k_out=ko, k_time=kt)
    q = pos.copy(); q[[1, 3, 4]] *= c ** 0.5                 # variances enter squared (util.jl:36-55; dtc.jl:31-37)
    f2 = scaled_gpar_objective(th(q), X, Z, t, y * c ** 0.5, k_out=ko, k_time=kt)
    assert abs(f2 - (f1 - 0.5 * n * np.log(c))) <= 1e-8 * max(1.0, abs(f1))


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 10_000), kind=SS_KINDS, n=st.integers(2, 80), a=st.floats(-2.0, 2.0), b=st.floats(-2.0, 2.0))
def test_smoother_mean_is_linear_in_the_data_and_variance_ignores_it(seed, kind, n, a, b):
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.exponential(1 / 30, n)); y1 = rng.normal(size=n); y2 = rng.normal(size=n)
    rv = np.where(rng.uniform(size=n) < 0.2, 1e10, 0.04)     # the "infinite noise at test points" trick
    _, m1, v1 = cport.kalman_smooth_batch(kind, t, y1[None], 0.7, 1.2, rv)
    _, m2, v2 = cport.kalman_smooth_batch(kind, t, y2[None], 0.7, 1.2, rv)
    _, m3, v3 = cport.kalman_smooth_batch(kind, t, (a * y1 + b * y2)[None], 0.7, 1.2, rv)
    assert np.max(np.abs(m3 - (a * m1 + b * m2))) <= 1e-9 * max(1.0, np.max(np.abs(m1)) + np.max(np.abs(m2)))
    assert np.array_equal(v1, v2) and np.array_equal(v1, v3)


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 10_000), kind=SS_KINDS, n=st.integers(3, 60))
def test_filter_logpdf_equals_the_dense_gaussian(seed, kind, n):
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.exponential(1 / 30, n)); y = rng.normal(size=n)
    l, s, nv = 0.6, 1.4, 0.09
    K = oracle.lgssm.dense_time_cov(kind, t, l, s, nv)
    sign, logdet = np.linalg.slogdet(K)
    dense = -0.5 * (n * np.log(2 * np.pi) + logdet + y @ np.linalg.solve(K, y))
    assert abs(oracle.kalman_logpdf(kind, t, y, l, s, nv) - dense) <= 1e-9 * max(1.0, abs(dense))
