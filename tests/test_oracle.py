"""CPU tests of the oracle itself (no GPU).  The reference ships no tests; these pin the oracle on
the only self-check it contains (examples/dtc_example.jl:8-64), the mask docstring example
(src/util.jl:57-96), independent dense ground truths and 50-digit mpmath spot values."""
import numpy as np
import pytest
import oracle
from oracle import cport
from oracle.dtc import scaled_gpar_objective
from oracle.kernels import stretched_pairwise


def small_dataset(rng, n=30):
    """Shape of generate_small_dataset (src/data/toy_data.jl:59-74): x = range(0, n/30, n), three
    chained outputs, noise std 0.05^2 (the reference's quirk, toy_data.jl:29)."""
    x = np.linspace(0.0, n / 30.0, n)
    nz = lambda: rng.normal(0.0, 0.05 ** 2, n)
    y1 = -np.sin(10 * np.pi * (x + 1)) / (2 * x + 1) - x ** 4 + nz()
    y2 = np.cos(y1) ** 2 + np.sin(3 * x) + nz()
    y3 = y2 * y1 ** 2 + 3 * x + nz()
    return x, y1, y2, y3


def test_unpack_transforms():
    # src/util.jl:36-55: exp(p) + 1e-3
    assert oracle.unpack_gp([0.0, np.log(2.0), -50.0]) == pytest.approx((1.001, 2.001, 1e-3 + np.exp(-50.0)), rel=1e-15)
    assert oracle.unpack_gpar([0, 0, 0, 0, 0]) == pytest.approx((1.001,) * 5, rel=1e-15)


def test_mask_docstring_example():
    # src/util.jl:57-96, literal numbers
    xs = oracle.to_colvecs([[-0.08, 0.39, 1.31], [-1.60, 0.58, -2.34], [-1.71, -0.16, 1.26]])
    ys = oracle.to_colvecs([[1.82, 1.06, 0.54], [0.14, 1.02, -0.38], [-1.45, -0.007, 0.006]])
    assert xs.shape == (3, 3) and xs[0].tolist() == [-0.08, -1.60, -1.71]   # N records of D features
    time_K = stretched_pairwise(oracle.EQ, xs, ys, oracle.get_time_mask(3))
    assert np.allclose(time_K, oracle.pairwise(oracle.EQ, xs[:, :1], ys[:, :1]), rtol=1e-15)
    out_K = stretched_pairwise(oracle.EQ, xs, ys, oracle.get_output_mask(3))
    assert np.allclose(out_K, oracle.pairwise(oracle.EQ, xs[:, 1:], ys[:, 1:]), rtol=1e-14)
    with pytest.raises(ValueError):
        oracle.get_output_mask(1)


def test_kernel_mpmath_spot_values():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    for r in (0.0, 1e-3, 0.37, 2.5, 11.0):
        R = mp.mpf(r)
        ref = {oracle.EQ: mp.exp(-R * R / 2), oracle.MATERN12: mp.exp(-R),
               oracle.MATERN32: (1 + mp.sqrt(3) * R) * mp.exp(-mp.sqrt(3) * R),
               oracle.MATERN52: (1 + mp.sqrt(5) * R + 5 * R * R / 3) * mp.exp(-mp.sqrt(5) * R)}
        for kind, v in ref.items():
            assert float(oracle.base_kernel(kind, r)) == pytest.approx(float(v), rel=2e-15)


@pytest.mark.parametrize("kind", [oracle.MATERN12, oracle.MATERN32, oracle.MATERN52])
def test_transition_closed_form_vs_expm(kind):
    from scipy.linalg import expm, solve_continuous_lyapunov
    F, Pinf, H = oracle.sde_matrices(kind)
    for a in (1e-6, 0.03, 0.7, 4.0):
        assert np.allclose(oracle.transition(kind, a), expm(F * a), rtol=1e-13, atol=1e-15)
    # stationary covariance solves the Lyapunov equation F P + P F^T + L q L^T = 0 (only last entry driven)
    res = F @ Pinf + Pinf @ F.T
    res[-1, -1] = 0.0
    assert np.max(np.abs(res)) < 1e-13
    # the state-space covariance function reproduces the kernel: H A(a) Pinf H^T = kappa(a)
    for a in (0.0, 0.2, 1.7):
        assert H @ oracle.transition(kind, a) @ Pinf @ H == pytest.approx(float(oracle.base_kernel(kind, a)), rel=1e-13)


@pytest.mark.parametrize("kind", [oracle.MATERN12, oracle.MATERN32, oracle.MATERN52])
def test_filter_equals_dense_gp(kind):
    rng = np.random.default_rng(kind)
    n = 200
    t = np.sort(rng.uniform(0, 8, n)); y = rng.normal(size=n)
    l, s, nv = 0.7, 1.3, 0.04
    K = oracle.dense_time_cov(kind, t, l, s, nv)
    lml, alpha = oracle.kalman_decorrelate(kind, t, y, l, s, nv)
    assert lml == pytest.approx(oracle.exact_logpdf(K, 0.0, y), rel=1e-12)
    # decorrelate == Cholesky whitening (why it can replace chol_Sigma_y.U' \ . in Stheno's DTC)
    assert np.allclose(alpha, np.linalg.solve(np.linalg.cholesky(K), y), atol=1e-10)
    # C twin
    lml_c, alpha_c = cport.kalman_decorrelate(kind, t, y, l, s, nv)
    assert lml_c == pytest.approx(lml, rel=1e-13) and np.allclose(alpha_c, alpha, atol=1e-12)


@pytest.mark.parametrize("kind", [oracle.MATERN12, oracle.MATERN32, oracle.MATERN52])
def test_smoother_equals_dense_posterior_with_inf_noise_trick(kind):
    """get_sde_predictions protocol (temporal_gp_inference.jl:55-113): 1e10 noise at test points.
    Matches the exact GP posterior to ~1e-9 only (the trick is not exact) — parity target is this
    formulation."""
    rng = np.random.default_rng(10 + kind)
    ntr, nte = 60, 40
    ttr = np.sort(rng.uniform(0, 5, ntr)); tte = rng.uniform(0, 5.5, nte); ytr = np.sin(2 * ttr) + 0.1 * rng.normal(size=ntr)
    l, var, sig = 0.8, 1.1, 0.1
    mean, v = oracle.sde_predictions(kind, ttr, ytr, tte, l, var, sig)
    Kff = oracle.pairwise(kind, ttr[:, None], ttr[:, None], l, var ** 2)
    Ksf = oracle.pairwise(kind, tte[:, None], ttr[:, None], l, var ** 2)
    m0, v0 = oracle.exact_posterior(Kff, Ksf, np.full(nte, var ** 2), sig ** 2, ytr, obs_noise=0.0)
    assert np.allclose(mean, m0, atol=2e-8)
    assert np.allclose(v, v0, rtol=1e-6, atol=1e-9)
    mean_c, v_c = oracle.sde_predictions(kind, ttr, ytr, tte, l, var, sig, smooth=cport.kalman_smooth)
    assert np.allclose(mean_c, mean, atol=1e-12) and np.allclose(v_c, v, rtol=1e-10)


def test_dtc_lgssm_equals_dense_dtc_reference_selfcheck():
    """examples/dtc_example.jl:8-64 — compare_dtc_with_Stheno_dtc: N=30 small set, pseudo-points =
    every third y1, Matern52 both, sigma_obs=0.05, temporal sigma=0.04 (:41-50)."""
    rng = np.random.default_rng(0)
    x, y1, y2, _ = small_dataset(rng)
    pseudo = y1[::3]
    assert len(pseudo) == 10
    Cfu = oracle.pairwise(oracle.MATERN52, y1[:, None], pseudo[:, None])
    cov_u = oracle.pairwise(oracle.MATERN52, pseudo[:, None], pseudo[:, None]) + 0.05 ** 2 * np.eye(10)
    gpar_dtc, gpar_A = oracle.compute_gpar_dtc_objective(Cfu, cov_u, x, y2, oracle.MATERN52, 1.0, 1.0, 0.04 ** 2)
    func_dtc, func_A = oracle.dtc_dense(Cfu, cov_u, oracle.dense_time_cov(oracle.MATERN52, x, 1.0, 1.0, 0.04 ** 2), y2)
    assert abs(gpar_dtc - func_dtc) <= 1e-11 * abs(func_dtc)
    assert np.max(np.abs(gpar_A - func_A)) < 1e-9
    # O(N) log-determinant identity replaces the dense logdet(noise_matrix) (dtc.jl:99,123)
    dtc_on, _ = oracle.compute_gpar_dtc_objective(Cfu, cov_u, x, y2, oracle.MATERN52, 1.0, 1.0, 0.04 ** 2, dense_logdet=False)
    assert abs(dtc_on - gpar_dtc) <= 1e-12 * abs(gpar_dtc)


def test_scaled_objective_equals_direct_gaussian():
    rng = np.random.default_rng(3)
    n, m = 200, 17
    t = np.sort(rng.uniform(0, 3, n)); X = rng.normal(size=(n, 2)); Z = rng.normal(size=(m, 2)); y = rng.normal(size=n)
    th = np.array([0.1, -0.2, 0.3, 0.1, -1.0])
    tl, tv, ol, ov, ns = oracle.unpack_gpar(th)
    val = scaled_gpar_objective(th, X, Z, t, y, dense_logdet=True)
    Cfu = oracle.pairwise(oracle.MATERN52, X, Z, ol, ov ** 2)
    cu = oracle.pairwise(oracle.MATERN52, Z, Z, ol, ov ** 2) + ns ** 2 * np.eye(m)
    Sig = Cfu @ np.linalg.solve(cu, Cfu.T) + oracle.dense_time_cov(oracle.MATERN52, t, tl, tv ** 2, ns ** 2)
    assert val == pytest.approx(oracle.exact_logpdf(Sig, 0.0, y), rel=1e-11)
    # C column filters == numpy column filters
    val_c = scaled_gpar_objective(th, X, Z, t, y, decorrelate=cport.kalman_decorrelate)
    assert val_c == pytest.approx(val, rel=1e-12)
    # collapsed statistics form
    lml, alpha = oracle.kalman_decorrelate(oracle.MATERN52, t, y, tl, tv ** 2, ns ** 2)
    beta = cport.decorrelate_columns(oracle.MATERN52, t, Cfu, tl, tv ** 2, ns ** 2)
    logdet = -2 * lml - n * np.log(2 * np.pi) - alpha @ alpha
    assert oracle.gpar_dtc_collapsed(beta.T @ beta, beta.T @ alpha, alpha @ alpha, logdet, n, cu) == pytest.approx(val, rel=1e-12)


def test_plain_dtc_and_vfe_against_dense():
    rng = np.random.default_rng(4)
    n, m = 150, 12
    X = rng.normal(size=(n, 1)); Z = np.linspace(-2, 2, m)[:, None]; y = rng.normal(size=n)
    Cfu = oracle.pairwise(oracle.EQ, X, Z, 0.9, 1.2); cu = oracle.pairwise(oracle.EQ, Z, Z, 0.9, 1.2) + 1e-6 * np.eye(m)
    d, traat = oracle.dtc_diag(Cfu, cu, 0.05, y)
    d0, A = oracle.dtc_dense(Cfu, cu, 0.05 * np.eye(n), y)
    assert d == pytest.approx(d0, rel=1e-12) and traat == pytest.approx(np.sum(A * A), rel=1e-12)
    Qff = Cfu @ np.linalg.solve(cu, Cfu.T)
    assert d == pytest.approx(oracle.exact_logpdf(Qff, 0.05, y), rel=1e-9)
    elbo = oracle.elbo_diag(Cfu, cu, n * 1.2, 0.05, y)
    assert elbo == pytest.approx(d - 0.5 * (n * 1.2 - np.trace(Qff)) / 0.05, rel=1e-9)
    assert elbo <= oracle.exact_logpdf(oracle.pairwise(oracle.EQ, X, X, 0.9, 1.2), 0.05, y) + 1e-9   # a lower bound


def test_compute_q_u_against_dense():
    """q(u) of gpar_scaled_inference.jl:141-196 against the dense formula
    D = I + L_u^{-1} Cuf Sigma_y^{-1} Cfu L_u^{-T},  m_e = D^{-1} L_u^{-1} Cuf Sigma_y^{-1} y."""
    rng = np.random.default_rng(5)
    n, m = 80, 7
    t = np.sort(rng.uniform(0, 3, n)); X = rng.normal(size=(n, 1)); Z = np.linspace(-2, 2, m)[:, None]; y = rng.normal(size=n)
    Cfu = oracle.pairwise(oracle.MATERN52, X, Z, 1.1, 0.8); Cuu = oracle.pairwise(oracle.MATERN52, Z, Z, 1.1, 0.8)
    m_e, Dinv, U_u = oracle.compute_q_u(Cfu, Cuu, t, y, oracle.MATERN52, 0.6, 1.4, 0.03)
    Sy = oracle.dense_time_cov(oracle.MATERN52, t, 0.6, 1.4, 0.03)
    Lu = np.linalg.cholesky(Cuu)
    B = np.linalg.solve(Lu, Cfu.T)
    D = np.eye(m) + B @ np.linalg.solve(Sy, B.T)
    assert np.allclose(U_u, Lu.T, rtol=1e-12)
    assert np.allclose(Dinv, np.linalg.inv(D), rtol=1e-8, atol=1e-12)
    assert np.allclose(m_e, np.linalg.solve(D, B @ np.linalg.solve(Sy, y)), rtol=1e-8, atol=1e-12)


def test_exact_gpar_kernel_composition():
    # optimized.jl:132-144: time_var^2 k_t(|dx1|/time_l) + out_var^2 k_o(||dx_2:D||/out_l)
    rng = np.random.default_rng(6)
    X = rng.normal(size=(9, 3))
    K = oracle.gpar_kernel_matrix(oracle.EQ, oracle.MATERN52, X, X, 0.5, 1.5, 2.0, 0.7)
    i, j = 2, 7
    kt = 1.5 ** 2 * np.exp(-0.5 * ((X[i, 0] - X[j, 0]) / 0.5) ** 2)
    r = np.linalg.norm(X[i, 1:] - X[j, 1:]) / 2.0
    ko = 0.7 ** 2 * (1 + np.sqrt(5) * r + 5 * r * r / 3) * np.exp(-np.sqrt(5) * r)
    assert K[i, j] == pytest.approx(kt + ko, rel=1e-14)
    assert np.allclose(np.diag(K), 1.5 ** 2 + 0.7 ** 2)


def test_gradient_oracle_against_central_differences():
    from oracle.grad import dtc_diag_value_and_grad
    rng = np.random.default_rng(7)
    n, m = 120, 9
    X = rng.normal(size=(n, 2)); Z = rng.normal(size=(m, 2)); y = rng.normal(size=n)
    th = np.array([0.2, -0.1, -1.2])

    def f(thv, kind, vfe):
        l, var, sig = oracle.unpack_gp(thv)
        Cfu = oracle.pairwise(kind, X, Z, l, var ** 2); cu = oracle.pairwise(kind, Z, Z, l, var ** 2) + sig ** 2 * np.eye(m)
        return oracle.elbo_diag(Cfu, cu, n * var ** 2, sig ** 2, y) if vfe else oracle.dtc_diag(Cfu, cu, sig ** 2, y)[0]

    for kind in (oracle.EQ, oracle.MATERN52):
        for vfe in (False, True):
            v, g = dtc_diag_value_and_grad(th, X, Z, y, kind, vfe)
            assert v == pytest.approx(f(th, kind, vfe), rel=1e-12)
            for i in range(3):
                e = np.zeros(3); e[i] = 1e-5
                fd = (f(th + e, kind, vfe) - f(th - e, kind, vfe)) / 2e-5
                assert g[i] == pytest.approx(fd, rel=1e-6, abs=1e-7)


def test_blas_cpu_port_of_logpdf_grad_matches_autograd():
    """bench.py's CPU baseline (oracle.dtc.dtc_diag_value_and_grad_np) == torch autograd oracle."""
    from oracle.grad import dtc_diag_value_and_grad
    from oracle.dtc import dtc_diag_value_and_grad_np
    rng = np.random.default_rng(8)
    for kind in range(4):
        for D in (1, 3):
            for vfe in (False, True):
                for jit in (-1.0, 1e-4):
                    X = rng.normal(size=(150, D)) * 2; Z = rng.normal(size=(13, D)) * 2; y = rng.normal(size=150)
                    th = rng.uniform(-1, 0.5, 3)
                    v0, g0 = dtc_diag_value_and_grad(th, X, Z, y, kind, vfe, jit)
                    v1, g1 = dtc_diag_value_and_grad_np(th, X, Z, y, kind, vfe, jit)
                    assert v1 == pytest.approx(v0, rel=1e-9)
                    assert np.allclose(g1, g0, rtol=1e-6, atol=1e-8 * np.max(np.abs(g0)))


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_lgssm_gradient_oracle_against_central_differences(kind):
    """The torch twin of the sequential filter reproduces the NumPy oracle's logpdf, its autograd
    gradient agrees with central differences, and the scale identity the device uses for d/ds holds:
    s dF/ds + sigma^2 dF/dsigma^2 = -1/2 (N - sum alpha^2)."""
    from oracle.grad import lgssm_logpdf_value_and_grad
    rng = np.random.default_rng(40 + kind)
    n = 60
    t = np.cumsum(rng.exponential(0.1, n)); y = rng.normal(size=n)
    th = rng.uniform(-1.0, 0.3, 3)

    def f(th_):
        l, var, sig = oracle.unpack_gp(th_)
        return oracle.kalman_logpdf(kind, t, y, l, var ** 2, sig ** 2)

    v, g = lgssm_logpdf_value_and_grad(th, t, y, kind)
    assert abs(v - f(th)) <= 1e-11 * abs(v)
    h = 1e-6
    for i in range(3):
        e = np.zeros(3); e[i] = h
        fd = (f(th + e) - f(th - e)) / (2 * h)
        assert abs(g[i] - fd) <= 1e-6 * max(1.0, abs(fd))
    l, var, sig = oracle.unpack_gp(th)
    _, alpha = oracle.kalman_decorrelate(kind, t, y, l, var ** 2, sig ** 2)
    ds = g[1] / (2 * var * np.exp(th[1])); dn = g[2] / (2 * sig * np.exp(th[2]))
    assert abs(var ** 2 * ds + sig ** 2 * dn + 0.5 * (n - alpha @ alpha)) <= 1e-9 * n


def test_scaled_gradient_oracle_against_central_differences():
    """torch twin of the scaled objective == NumPy oracle value; autograd == central differences."""
    from oracle.grad import scaled_dtc_value_and_grad
    rng = np.random.default_rng(9)
    n, m, d = 50, 6, 2
    t = np.sort(rng.uniform(0, 3, n)); X = rng.normal(size=(n, d)); Z = rng.normal(size=(m, d)); y = rng.normal(size=n)
    th = rng.uniform(-1.0, 0.3, 5)
    v, g = scaled_dtc_value_and_grad(th, X, Z, t, y, k_time=3, k_out=3)
    f = lambda th_: scaled_gpar_objective(th_, X, Z, t, y, k_out=3, k_time=3)
    assert abs(v - f(th)) <= 1e-10 * abs(v)
    h = 1e-6
    for i in range(5):
        e = np.zeros(5); e[i] = h
        fd = (f(th + e) - f(th - e)) / (2 * h)
        assert abs(g[i] - fd) <= 1e-6 * max(1.0, abs(fd))
