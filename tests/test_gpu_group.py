"""GPU: the single-process multi-device part of the C ABI (gpar_group_*).  With one member the group
must reproduce the single-context entry points bit for bit; with two devices (skipped on a one-GPU box) the
members work on different data concurrently and the NCCL-gathered tables must equal the per-context results."""
import ctypes
import os
import numpy as np
import pytest
import oracle
from oracle.dtc import scaled_gpar_objective

pytestmark = pytest.mark.gpu


def device_count():
    lib = ctypes.CDLL("libcudart.so.12")
    n = ctypes.c_int(0)
    lib.cudaGetDeviceCount(ctypes.byref(n))
    return n.value


def problem(seed, n=3000, m=48, d=2):
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.exponential(1 / 30, n))
    X = rng.normal(size=(n, d)); Z = X[:: n // m][:m].copy()
    y = np.sin(t) + 0.5 * X[:, 0] + 0.1 * rng.normal(size=n)
    return t, X, Z, y


def load(c, t, X, Z, y):
    c.set_times(t); c.set_inputs(X); c.set_pseudo(Z); c.set_outputs(y); c.set_noise_vector(None)


def test_group_of_one_reproduces_the_context_entry_points(ctx):
    import gpar_at_scale_b200 as gp
    t, X, Z, y = problem(7)
    g = gp.Group([0])
    try:
        load(g.members[0], t, X, Z, y); load(ctx, t, X, Z, y)
        th3 = np.array([[0.1, -0.2, -1.5]]); th5 = np.array([[0.2, 0.1, -0.3, 0.2, -1.0]])
        v, gr, codes = g.dtc_logpdf(gp.MATERN52, th3, grad=True)
        v0, g0 = ctx.dtc_logpdf(gp.MATERN52, th3[0], grad=True)
        assert codes[0] == 0 and v[0] == v0 and np.array_equal(gr[0], g0)
        v, gr, codes = g.scaled_dtc(gp.MATERN52, gp.MATERN52, th5, grad=True)
        v0, g0 = ctx.scaled_dtc_grad(gp.MATERN52, gp.MATERN52, th5[0])
        assert codes[0] == 0 and v[0] == v0 and np.array_equal(gr[0], g0)
        # and against the oracle, through the group
        ref = scaled_gpar_objective(th5[0], X, Z, t, y)
        assert abs(g.scaled_dtc(gp.MATERN52, gp.MATERN52, th5)[0][0] - ref) <= 1e-8 * abs(ref)
        # a non-PD member evaluation is reported in codes, not raised
        bad = np.array([[0.0, 40.0, 0.0, 40.0, -30.0]])
        _, codes = g.scaled_dtc(gp.MATERN52, gp.MATERN52, bad)
        assert codes[0] in (0, 3)
    finally:
        g.close()


def test_group_chain_broadcast_and_input_column(ctx):
    """Posterior means passed down the chain: broadcast (host values, then a resident smoother result) and
    gpar_set_inputs_column; the objective must equal the one on host-assembled inputs."""
    import gpar_at_scale_b200 as gp
    t, X, Z, y = problem(8)
    g = gp.Group([0])
    try:
        m = g.members[0]
        load(m, t, X, Z, y)
        col = np.cos(t)
        back = g.broadcast(0, values=col)
        assert np.array_equal(back, col)
        m.set_inputs_column(1)                         # from the chain buffer
        X2 = X.copy(); X2[:, 1] = col
        load(ctx, t, X2, Z, y)
        th5 = np.array([0.2, 0.1, -0.3, 0.2, -1.0])
        assert m.scaled_dtc(gp.MATERN52, gp.MATERN52, th5) == ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, th5)
        # resident result of a smoother -> chain buffer -> column 0
        _, mean, _ = m.lgssm_smooth(gp.MATERN52, np.array([0.0, 0.0, -1.0]), keep_on_device=False)
        back = g.broadcast(0, n=len(t))
        assert np.array_equal(back, mean[0])
        m.set_inputs_column(0)
        X2[:, 0] = mean[0]
        ctx.set_inputs(X2)
        assert m.scaled_dtc(gp.MATERN52, gp.MATERN52, th5) == ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, th5)
        with pytest.raises(gp.GparError):
            m.set_inputs_column(5)
    finally:
        g.close()


def test_group_chain_merged_means_reach_the_next_output(ctx):
    """The prediction chain `[test_y1, y2_out]` (GPAR_scaled_examples.jl:172) device to device.  After a smoother /
    prediction on a MERGED train+test problem the resident result is in sorted train+test order; what
    gpar_group_broadcast sends down the chain is the N* test means in TEST order (gpar_take_test's first array), and
    gpar_set_merged_test_column writes them into the later output's merged inputs at the test locations."""
    import gpar_at_scale_b200 as gp
    t, X, Z, y = problem(9)
    rng = np.random.default_rng(90)
    ns = 1700
    ts = np.sort(rng.uniform(t[0] - 1.0, t[-1] + 2.0, ns)); ts[::9] = t[rng.integers(0, len(t), len(ts[::9]))]; ts = np.sort(ts)
    Xs = rng.normal(size=(ns, X.shape[1]))
    th = np.log([1.5, 1.0, 0.2]); sig2 = (np.exp(th[2]) + 1e-3) ** 2
    params = np.array([1.3, 0.9, 1.1, 0.8, 0.3]); W = rng.normal(size=(Z.shape[0], 8))
    g = gp.Group([0])
    try:
        m = g.members[0]
        m.set_merged(t, y, ts, sig2)                                  # earlier output: time-only smoother on the merged grid
        m.lgssm_smooth(gp.MATERN52, th, keep_on_device=True)
        means, _ = m.take_test()
        back = g.broadcast(0, n=ns)
        assert np.array_equal(back, means)                            # test order, not the first N* sorted entries
        with pytest.raises(gp.GparError):
            g.broadcast(0, n=len(t) + ns)                             # a merged result is N* values, nothing else
        Xs0 = Xs.copy(); Xs0[:, 1] = 0.0                              # later output: column 1 of the test inputs comes from the chain
        m.set_pseudo(Z); m.set_merged(t, y, ts, params[4] ** 2, X=X, Xs=Xs0)
        m.set_merged_test_column(1)
        m.scaled_predict(gp.MATERN52, gp.MATERN52, params, W, keep_on_device=True)
        a1, b1 = m.take_test()
        Xs1 = Xs.copy(); Xs1[:, 1] = means                            # host path of chain.predict_chain
        ctx.set_pseudo(Z); ctx.set_merged(t, y, ts, params[4] ** 2, X=X, Xs=Xs1)
        ctx.scaled_predict(gp.MATERN52, gp.MATERN52, params, W, keep_on_device=True)
        a2, b2 = ctx.take_test()
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
        ctx.set_merged_test_column(1, means + 1.0)                    # host values through the same entry point
        with pytest.raises(gp.GparError):
            ctx.set_merged_test_column(7)
        # a later compute call reuses the result buffers: take_test / broadcast must fail loudly, not read stale memory
        m.set_merged(t, y, ts, sig2); m.lgssm_smooth(gp.MATERN52, th, keep_on_device=True); m.take_test()
        m.lgssm_logpdf(gp.MATERN52, th)
        with pytest.raises(gp.GparError):
            m.take_test()
        with pytest.raises(gp.GparError):
            g.broadcast(0, n=ns)
    finally:
        ctx.set_noise_vector(None)
        g.close()


def test_group_sharded_dtc_single_member_is_the_plain_entry_point(ctx):
    """gpar_group_dtc_logpdf_sharded with one member: same kernels, an all-reduce over one rank -> identical numbers,
    value and gradient, DTC and VFE."""
    import gpar_at_scale_b200 as gp
    rng = np.random.default_rng(31)
    n, m = 6000, 70
    x = rng.uniform(0, 10, n); z = np.linspace(0, 10, m); y = np.sin(x) + 0.1 * rng.normal(size=n)
    th = np.log([1.3, 0.9, 0.2])
    g = gp.Group([0])
    try:
        mb = g.members[0]
        mb.set_inputs(x); mb.set_pseudo(z); mb.set_outputs(y)
        ctx.set_inputs(x); ctx.set_pseudo(z); ctx.set_outputs(y)
        for vfe in (False, True):
            v, gr = g.dtc_logpdf_sharded(gp.MATERN52, th, vfe=vfe, grad=True)
            v0, g0 = ctx.dtc_logpdf(gp.MATERN52, th, vfe=vfe, grad=True)
            assert v == v0 and np.array_equal(gr, g0)
            assert g.dtc_logpdf_sharded(gp.MATERN52, th, vfe=vfe) == pytest.approx(v0, rel=1e-12)
    finally:
        g.close()


@pytest.mark.skipif(device_count() < 2, reason="needs two devices (gpurun --gpus 2)")
def test_group_sharded_dtc_two_devices(ctx):
    """Rows split unevenly over two devices: the all-reduced statistics give the objective of the whole data set."""
    import gpar_at_scale_b200 as gp
    rng = np.random.default_rng(32)
    n, m, d = 9001, 130, 2
    X = rng.normal(size=(n, d)) * 2; Z = rng.normal(size=(m, d)) * 2; y = np.sin(X[:, 0]) + 0.1 * rng.normal(size=n)
    th = np.log([1.5, 1.1, 0.15])
    ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y)
    g = gp.Group([0, 1])
    try:
        cut = 3777
        for mb, sl in zip(g.members, (slice(0, cut), slice(cut, n))):
            mb.set_inputs(X[sl]); mb.set_pseudo(Z); mb.set_outputs(y[sl])
        for vfe in (False, True):
            v, gr = g.dtc_logpdf_sharded(gp.MATERN52, th, vfe=vfe, grad=True)
            v0, g0 = ctx.dtc_logpdf(gp.MATERN52, th, vfe=vfe, grad=True)
            assert abs(v - v0) <= 1e-11 * abs(v0) and np.max(np.abs(gr - g0)) <= 1e-9 * np.max(np.abs(g0))
        g.members[1].set_pseudo(Z[:100])
        with pytest.raises(gp.GparError, match="pseudo-inputs"):
            g.dtc_logpdf_sharded(gp.MATERN52, th)
    finally:
        g.close()


def _loopback_group(n):
    import gpar_at_scale_b200 as gp
    os.environ["GPAR_GROUP_LOOPBACK"] = "1"
    try:
        return gp.Group([0] * n)
    finally:
        del os.environ["GPAR_GROUP_LOOPBACK"]


@pytest.mark.parametrize("nmem", [1, 2, 3, 5])
def test_group_scaled_dtc_row_sharded_loopback(ctx, nmem):
    """gpar_group_scaled_dtc_sharded on a loopback group (several members on device 0, collectives as device copies): the
    slice logic — filter on the full (t, y), per-slice pass 1 / transitions, slice summaries, entering states, pass 2, SYRK,
    summed (G, g), tail — reproduces the one-device objective to 1e-11 for uneven slices (including one shorter than a
    whitening chunk), all three time kernels and a multi-tile M; a poorly conditioned cov(u) takes the whitened-panel
    path on every member; the sharded plain DTC goes through the same loopback collectives."""
    import gpar_at_scale_b200 as gp
    rng = np.random.default_rng(40 + nmem)
    g = _loopback_group(nmem)
    try:
        for n, m, d, kt, ko in ((9001, 130, 2, 3, 3), (5003, 40, 1, 2, 0), (3001, 17, 3, 1, 2)):
            t = np.sort(rng.uniform(0, n / 30, n)); X = rng.normal(size=(n, d)); Z = rng.normal(size=(m, d)); y = rng.normal(size=n)
            th = rng.uniform(-1.0, 0.3, 5)
            ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(t); ctx.set_outputs(y); ctx.set_noise_vector(None)
            v0 = ctx.scaled_dtc(kt, ko, th)
            cuts = sorted(int(c) * 4 for c in rng.choice(np.arange(1, n // 4), size=nmem - 1, replace=False))
            if nmem == 3:
                cuts[1] = cuts[0] + 8                        # a slice far shorter than a whitening chunk
            bounds = [0] + cuts + [n]
            lo = g.load_row_slices(X, Z, t, y, bounds)
            v = g.scaled_dtc_sharded(kt, ko, th, lo)
            assert abs(v - v0) <= 1e-11 * abs(v0), (nmem, n, m, bounds, v, v0)
            # value + gradient: forward-mode tangents per slice, tangent carries over the slice boundaries (second all-gather)
            vg0, g0 = ctx.scaled_dtc_grad(kt, ko, th)
            vg, gr = g.scaled_dtc_sharded(kt, ko, th, lo, grad=True)
            assert abs(vg - vg0) <= 1e-11 * abs(vg0) and np.max(np.abs(gr - g0)) <= 1e-9 * np.max(np.abs(g0)), (nmem, n, m, bounds, gr, g0)
        # default (equal) slices and the error behaviour
        n, m = 20000, 64
        t = np.arange(n) / 30.0; X = rng.normal(size=(n, 2)); Z = rng.normal(size=(m, 2)); y = rng.normal(size=n)
        th = np.array([0.2, 0.1, -0.3, 0.2, -1.0])
        ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(t); ctx.set_outputs(y)
        lo = g.load_row_slices(X, Z, t, y)
        assert abs(g.scaled_dtc_sharded(3, 3, th, lo) - ctx.scaled_dtc(3, 3, th)) <= 1e-11 * abs(ctx.scaled_dtc(3, 3, th))
        # poorly conditioned cov(u) (large output variance): the whitened-panel path on every slice
        thi = np.array([0.2, 0.1, 1.5, 7.0, -1.0])
        vi0 = ctx.scaled_dtc(3, 3, thi)
        assert abs(g.scaled_dtc_sharded(3, 3, thi, lo) - vi0) <= 1e-9 * abs(vi0)
        # ... and its gradient in whitened coordinates on every slice (a whitened copy of the slice's panel), as on one device
        vgi0, gi0 = ctx.scaled_dtc_grad(3, 3, thi)
        vgi, gi = g.scaled_dtc_sharded(3, 3, thi, lo, grad=True)
        assert abs(vgi - vgi0) <= 1e-9 * abs(vgi0) and np.max(np.abs(gi - gi0)) <= 1e-6 * np.max(np.abs(gi0)), (gi, gi0)
        os.environ["GPAR_GRAD_WHITENED"] = "1"                  # forced on the well-conditioned problem: equals the collapsed form
        try:
            vgw, gw = g.scaled_dtc_sharded(3, 3, th, lo, grad=True)
        finally:
            del os.environ["GPAR_GRAD_WHITENED"]
        vg0, g0 = ctx.scaled_dtc_grad(3, 3, th)
        assert abs(vgw - vg0) <= 1e-10 * abs(vg0) and np.max(np.abs(gw - g0)) <= 1e-8 * np.max(np.abs(g0)), (gw, g0)
        # q(u) on the same slices (bare Cuu: the whitened-panel path on every member) against the one-device compute_q_u
        params = np.array([0.9, 1.2, 1.4, 0.8, 0.12])
        m0, D0, U0 = ctx.compute_q_u(3, 3, params)
        m1, D1, U1 = g.compute_q_u_sharded(3, 3, params, lo)
        # (two float64 evaluations with different summation orders: the whitened panel A = U_u' \ beta' carries ~ eps cond(U_u),
        # which inv(D) = (I + A A')^-1 amplifies by cond(D))
        eps = np.finfo(float).eps
        tolq = max(1e-8, 100 * eps * np.linalg.cond(U0) ** 2); told = max(tolq, 100 * eps * np.linalg.cond(U0) * np.linalg.cond(D0))
        rel = lambda a, b_: np.max(np.abs(a - b_)) / np.max(np.abs(b_))
        assert rel(U1, U0) <= 1e-12 and rel(m1, m0) <= tolq and rel(D1, D0) <= told, (rel(m1, m0), rel(D1, D0), tolq, told)
        # the seeded device sampler on the sharded statistics: the draws of gpar_sample_q_u with the same seed
        W0, E0 = ctx.sample_q_u(3, 3, params, 11, 7, return_host=True)
        W1, E1 = g.sample_q_u_sharded(3, 3, params, lo, 11, 7)
        assert rel(E1, E0) <= 10 * told and rel(W1, W0) <= 10 * told * np.linalg.cond(U0), (rel(E1, E0), rel(W1, W0))
        if nmem > 1:
            bad = lo.copy(); bad[1] += 4
            with pytest.raises(gp.GparError, match="starts at row"):
                g.scaled_dtc_sharded(3, 3, th, bad)
            g.members[1].set_pseudo(Z[:50])
            with pytest.raises(gp.GparError, match="pseudo-inputs"):
                g.scaled_dtc_sharded(3, 3, th, lo)
            # plain DTC through the loopback all-reduce
            x1 = rng.uniform(0, 10, n); z1 = np.linspace(0, 10, m); y1 = np.sin(x1) + 0.1 * rng.normal(size=n)
            th3 = np.log([1.3, 0.9, 0.2])
            ctx.set_inputs(x1); ctx.set_pseudo(z1); ctx.set_outputs(y1)
            edges = np.linspace(0, n, nmem + 1).astype(int)
            for i, mb in enumerate(g.members):
                mb.set_inputs(x1[edges[i]:edges[i + 1]]); mb.set_pseudo(z1); mb.set_outputs(y1[edges[i]:edges[i + 1]])
            v, gr = g.dtc_logpdf_sharded(gp.MATERN52, th3, grad=True)
            v0, g0 = ctx.dtc_logpdf(gp.MATERN52, th3, grad=True)
            assert abs(v - v0) <= 1e-11 * abs(v0) and np.max(np.abs(gr - g0)) <= 1e-9 * np.max(np.abs(g0))
            # poorly conditioned cov(u) (EQ kernel, variance 900, jitter 1e-6): every member whitens its panels by L_u, the
            # whitened statistics are summed, member 0 finishes in whitened coordinates — as the one-device entry point does
            # (cond(cov(u)) ~ 6e10 here: two float64 evaluations with different summation orders agree to ~4e-9 / ~2e-5)
            thi3 = np.log([2.0, 30.0, 0.1])
            for vfe in (False, True):
                v, gr = g.dtc_logpdf_sharded(gp.EQ, thi3, vfe=vfe, jitter=1e-6, grad=True)
                v0, g0 = ctx.dtc_logpdf(gp.EQ, thi3, vfe=vfe, jitter=1e-6, grad=True)
                assert abs(v - v0) <= 1e-8 * abs(v0) and np.max(np.abs(gr - g0)) <= 1e-4 * np.max(np.abs(g0)), (vfe, v, v0, gr, g0)
                assert abs(g.dtc_logpdf_sharded(gp.EQ, thi3, vfe=vfe, jitter=1e-6) - ctx.dtc_logpdf(gp.EQ, thi3, vfe=vfe, jitter=1e-6)) <= 1e-8 * abs(v0)
    finally:
        g.close()


@pytest.mark.skipif(device_count() < 2, reason="needs two devices (gpurun --gpus 2)")
def test_group_scaled_dtc_row_sharded_devices(ctx):
    """The same over NCCL on every visible device: one all-gather + one all-reduce per evaluation."""
    import gpar_at_scale_b200 as gp
    rng = np.random.default_rng(52)
    nd = device_count()
    g = gp.Group(list(range(nd)))
    try:
        n, m, d = 200_003, 300, 2
        t = np.sort(rng.uniform(0, n / 30, n)); X = rng.normal(size=(n, d)); Z = rng.normal(size=(m, d)) * 1.5; y = rng.normal(size=n)
        th = np.array([0.2, 0.1, -0.3, 0.2, -1.0])
        ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(t); ctx.set_outputs(y); ctx.set_noise_vector(None)
        v0 = ctx.scaled_dtc(3, 3, th)
        lo = g.load_row_slices(X, Z, t, y)
        v = g.scaled_dtc_sharded(3, 3, th, lo)
        assert abs(v - v0) <= 1e-11 * abs(v0), (v, v0)
        vg0, g0 = ctx.scaled_dtc_grad(3, 3, th)
        vg, gr = g.scaled_dtc_sharded(3, 3, th, lo, grad=True)
        assert abs(vg - vg0) <= 1e-11 * abs(vg0) and np.max(np.abs(gr - g0)) <= 1e-9 * np.max(np.abs(g0)), (gr, g0)
        # q(u) and a short fit over NCCL
        params = np.array([0.9, 1.2, 1.4, 0.8, 0.12])
        m0, D0, U0 = ctx.compute_q_u(3, 3, params)
        m1, D1, U1 = g.compute_q_u_sharded(3, 3, params, lo)
        eps = np.finfo(float).eps
        tolq = max(1e-8, 100 * eps * np.linalg.cond(U0) ** 2); told = max(tolq, 100 * eps * np.linalg.cond(U0) * np.linalg.cond(D0))
        rel = lambda a, b_: np.max(np.abs(a - b_)) / np.max(np.abs(b_))
        assert rel(U1, U0) <= 1e-12 and rel(m1, m0) <= tolq and rel(D1, D0) <= told
        from gpar_at_scale_b200 import neldermead
        fmin, xmin, calls = g.fit_sharded(3, 3, lo, th, iterations=5)
        rp = neldermead.optimize(lambda p_: -g.scaled_dtc_sharded(3, 3, p_, lo), th, iterations=5)
        assert fmin == rp.minimum and np.array_equal(xmin, rp.minimizer) and calls == rp.f_calls
    finally:
        g.close()


def test_fit_through_the_row_sharded_objective(ctx):
    """api.get_optim_scaled_gpar_params(group=...): Nelder-Mead on the row-sharded objective (3 loopback members) reaches the
    optimum of the one-device fit (values differ in the last digits, so the simplices may part ways late: compare optima)."""
    from gpar_at_scale_b200 import api
    rng = np.random.default_rng(61)
    n, m = 6000, 30
    t = np.arange(n) / 30.0
    X = np.sin(0.3 * t)[:, None] + 0.1 * rng.normal(size=(n, 1)); Z = np.linspace(X.min(), X.max(), m)[:, None]
    y = np.cos(2.0 * X[:, 0]) + 0.3 * np.sin(0.05 * t) + 0.1 * rng.normal(size=n)
    kw = dict(i_log_time_l=1.0, i_log_time_var=0.0, i_log_out_l=0.0, i_log_out_var=0.0, i_log_noise_sigma=-1.0, iterations=60, return_result=True)
    p1, r1 = api.get_optim_scaled_gpar_params(X, Z, t, y, ctx=ctx, **kw)
    g = _loopback_group(3)
    try:
        p3, r3 = api.get_optim_scaled_gpar_params(X, Z, t, y, group=g, **kw)
    finally:
        g.close()
    assert abs(r3.minimum - r1.minimum) <= 1e-6 * abs(r1.minimum), (r1.minimum, r3.minimum)
    # L-BFGS on the sharded gradient against L-BFGS on one device
    kw["iterations"] = 15
    _, l1 = api.get_optim_scaled_gpar_params(X, Z, t, y, ctx=ctx, optimizer="lbfgs", **kw)
    g = _loopback_group(2)
    try:
        _, l2 = api.get_optim_scaled_gpar_params(X, Z, t, y, group=g, optimizer="lbfgs", **kw)
    finally:
        g.close()
    assert abs(l2.minimum - l1.minimum) <= 1e-6 * abs(l1.minimum), (l1.minimum, l2.minimum)
    # the same fits with the optimiser inside the library (gpar_group_fit_sharded): the C++ Nelder-Mead / L-BFGS are operation-for-
    # operation twins of the Python ones and the objective is the same call, so minima and evaluation counts are identical
    from gpar_at_scale_b200 import neldermead, lbfgs
    g = _loopback_group(2)
    try:
        lo = g.load_row_slices(X, Z, t, y)
        th0 = np.array([1.0, 0.0, 0.0, 0.0, -1.0])
        fmin, xmin, calls = g.fit_sharded(3, 3, lo, th0, iterations=25)
        rp = neldermead.optimize(lambda p_: -g.scaled_dtc_sharded(3, 3, p_, lo), th0, iterations=25)
        assert fmin == rp.minimum and np.array_equal(xmin, rp.minimizer) and calls == rp.f_calls
        fl, xl, cl = g.fit_sharded(3, 3, lo, th0, iterations=6, optimizer="lbfgs")

        def fg(p_):
            v, gr = g.scaled_dtc_sharded(3, 3, p_, lo, grad=True)
            return -v, -gr
        rl = lbfgs.optimize(fg, th0, iterations=6)
        assert abs(fl - rl.minimum) <= 1e-12 * abs(rl.minimum) and np.allclose(xl, rl.minimizer, rtol=0, atol=1e-10)
    finally:
        g.close()
    with pytest.raises(ValueError):
        api.get_optim_scaled_gpar_params(X, Z, t, y, group=object(), n_restarts=4)
    # prediction with the two N x M stages (fit skipped here, q(u)) on the sharded path: the same host draws (same rng seed) through
    # the one-device q(u) and through the sharded q(u) give the same predictive means / stds
    ts = t[:500] + 0.5 / 30.0; Xs = np.sin(0.3 * ts)[:, None]
    kwp = dict(opt_params=p1, nsamples=20, sampler="host", ctx=ctx)
    mean1, std1 = api.get_gpar_scaled_predictions(X, Z, t, y, ts, Xs, rng=np.random.default_rng(3), **kwp)
    kwd = dict(opt_params=p1, nsamples=20, sampler="device", seed=5, ctx=ctx)
    mean1d, std1d = api.get_gpar_scaled_predictions(X, Z, t, y, ts, Xs, **kwd)
    g = _loopback_group(3)
    try:
        mean3, std3 = api.get_gpar_scaled_predictions(X, Z, t, y, ts, Xs, rng=np.random.default_rng(3), group=g, **kwp)
        mean3d, std3d = api.get_gpar_scaled_predictions(X, Z, t, y, ts, Xs, group=g, **kwd)      # seeded device draws on member 0
    finally:
        g.close()
    assert np.max(np.abs(mean3 - mean1)) <= 1e-6 * np.max(np.abs(mean1)) and np.max(np.abs(std3 - std1)) <= 1e-6 * np.max(std1)
    assert np.max(np.abs(mean3d - mean1d)) <= 1e-6 * np.max(np.abs(mean1d)) and np.max(np.abs(std3d - std1d)) <= 1e-6 * np.max(std1d)


def test_scaled_slice_abi_for_one_process_per_gpu_hosts(ctx):
    """gpar_scaled_slice_* (the row-sharded objective for hosts that own their collectives): (a) three slices on three
    contexts of device 0, the all-gather / all-reduce done by hand on torch tensors; (b) parallel.scaled_dtc_row_sharded
    through a world_size-1 NCCL process group.  Value and gradient equal the one-device entry points."""
    import socket
    import torch
    import torch.distributed as dist
    import gpar_at_scale_b200 as gp
    from gpar_at_scale_b200 import parallel
    rng = np.random.default_rng(71)
    n, m, d = 9001, 130, 2
    t = np.sort(rng.uniform(0, n / 30, n)); X = rng.normal(size=(n, d)); Z = rng.normal(size=(m, d)); y = rng.normal(size=n)
    th = rng.uniform(-1.0, 0.3, 5)
    ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_times(t); ctx.set_outputs(y); ctx.set_noise_vector(None)
    v0 = ctx.scaled_dtc(3, 3, th); vg0, g0 = ctx.scaled_dtc_grad(3, 3, th)
    dev = torch.device("cuda", 0)
    kw = dict(dtype=torch.float64, device=dev)
    b = [0, 2048, 2060, n]
    engines = [gp.Context(0) for _ in range(3)]
    try:
        for i, e in enumerate(engines):
            e.set_times(t); e.set_outputs(y); e.set_pseudo(Z); e.set_inputs(np.ascontiguousarray(X[b[i]:b[i + 1]]))
        for grad in (False, True):
            counts = [e.scaled_slice_begin(3, 3, th, b[i], grad) for i, e in enumerate(engines)]
            sc, stc = counts[0]
            assert all(c == (sc, stc) for c in counts)
            summaries = [torch.empty(sc, **kw) for _ in engines]
            for e, sm in zip(engines, summaries):
                e.scaled_slice_summary(sm)
            gathered = torch.cat(summaries)
            stats = [torch.empty(stc, **kw) for _ in engines]
            for i, (e, st) in enumerate(zip(engines, stats)):
                e.scaled_slice_stats(gathered, i, st)
            total = torch.stack(stats).sum(0)
            if not grad:
                for e in engines:                                   # every rank can finish
                    assert abs(e.scaled_slice_value(total) - v0) <= 1e-11 * abs(v0)
                continue
            s2 = [torch.empty(3 * sc, **kw) for _ in engines]
            for e, x in zip(engines, s2):
                e.scaled_slice_tangent_summary(total, x)
            g2 = torch.cat(s2)
            s5 = sum(e.scaled_slice_grad_partial(g2, i) for i, e in enumerate(engines))
            vg, gr = engines[0].scaled_slice_grad_finish(s5)
            assert abs(vg - vg0) <= 1e-11 * abs(vg0) and np.max(np.abs(gr - g0)) <= 1e-9 * np.max(np.abs(g0)), (gr, g0)
        with pytest.raises(gp.GparError):                            # a value-mode call on a gradient-mode slice
            engines[0].scaled_slice_value(total)
        with pytest.raises(gp.GparError):                            # begin has not run on this context
            gp.Context(0).scaled_slice_summary(summaries[0])
        engines[1].scaled_slice_begin(3, 3, th, b[1], False)
        engines[1].lgssm_logpdf(gp.MATERN52, np.zeros(3))            # any other compute call ends the slice evaluation (shared scratch)
        with pytest.raises(gp.GparError):
            engines[1].scaled_slice_summary(summaries[1])
    finally:
        for e in engines:
            e.close()
    # (b) the torch.distributed driver, one rank
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        assert parallel.row_slice_bounds(n, 1) == [0, n]
        v = parallel.scaled_dtc_row_sharded(ctx, 3, 3, th, 0, device=dev)
        vg, gr = parallel.scaled_dtc_row_sharded(ctx, 3, 3, th, 0, grad=True, device=dev)
        assert abs(v - v0) <= 1e-11 * abs(v0) and abs(vg - vg0) <= 1e-11 * abs(vg0) and np.max(np.abs(gr - g0)) <= 1e-9 * np.max(np.abs(g0))
    finally:
        dist.destroy_process_group()


def test_group_abi_error_behaviour():
    """Status codes and messages instead of crashes: duplicate devices, missing resident result, incomplete task,
    unknown optimiser, member without data."""
    import gpar_at_scale_b200 as gp
    from gpar_at_scale_b200 import _ffi
    with pytest.raises(gp.GparError):
        gp.Group([0, 0])                                  # one member per device
    with pytest.raises(gp.GparError):
        gp.Group([])
    g = gp.Group([0])
    try:
        with pytest.raises(gp.GparError, match="resident result"):
            g.broadcast(0, n=10)                          # nothing smoothed / predicted yet
        with pytest.raises(gp.GparError, match="source member"):
            g.broadcast(3, values=np.zeros(4))
        with pytest.raises(gp.GparError, match="must be set|not set|inputs"):
            g.scaled_dtc(gp.MATERN52, gp.MATERN52, np.zeros((1, 5)))       # member has no data: the member's message surfaces
        t = np.arange(50) / 30.0
        bad = [{"X": np.zeros((50, 2)), "Z": None, "y": np.zeros(50), "theta0": np.zeros(5)}]
        lib = g._lib
        arr = (_ffi.FitTask * 1)()
        y = _ffi.as_f64(np.zeros(50)); X = _ffi.as_f64(np.zeros((50, 2)))
        arr[0].y = _ffi.dptr(y); arr[0].X = _ffi.dptr(X); arr[0].D = 2; arr[0].Z = None; arr[0].M = 0
        out = np.zeros(1); th = np.zeros(5)
        st = lib.gpar_group_fit(g._h, _ffi.dptr(_ffi.as_f64(t)), 50, ctypes.cast(arr, ctypes.c_void_p), 1, 3, 3, 0, 2, _ffi.dptr(out), _ffi.dptr(th), None, None)
        assert st == _ffi.GPAR_ERR_INVALID and b"incomplete" in lib.gpar_group_last_error(g._h)
        arr[0].X = None; arr[0].D = 0
        st = lib.gpar_group_fit(g._h, _ffi.dptr(_ffi.as_f64(t)), 50, ctypes.cast(arr, ctypes.c_void_p), 1, 3, 3, 7, 2, _ffi.dptr(out), _ffi.dptr(th), None, None)
        assert st == _ffi.GPAR_ERR_INVALID and b"optimizer" in lib.gpar_group_last_error(g._h)
        # a time-only task with valid arguments runs
        arr[0].theta0[0] = 0.1; arr[0].theta0[1] = 0.0; arr[0].theta0[2] = -1.0
        yy = _ffi.as_f64(np.sin(t)); arr[0].y = _ffi.dptr(yy)
        st = lib.gpar_group_fit(g._h, _ffi.dptr(_ffi.as_f64(t)), 50, ctypes.cast(arr, ctypes.c_void_p), 1, 3, 3, 0, 3, _ffi.dptr(out), _ffi.dptr(th), None, None)
        assert st == _ffi.GPAR_OK and np.isfinite(out[0]) and np.all(np.isfinite(th[:3])) and np.all(np.isnan(th[3:]))
    finally:
        g.close()


def chain_tasks(seed, n=2500, m=40, outputs=3, restarts=2):
    rng = np.random.default_rng(seed)
    t = np.cumsum(rng.exponential(1 / 30, n))
    Y = np.zeros((outputs, n))
    Y[0] = np.sin(t) + 0.1 * rng.normal(size=n)
    for o in range(1, outputs):
        Y[o] = np.cos(0.7 * t) + 0.5 * Y[o - 1] + 0.1 * rng.normal(size=n)
    tasks = []
    for o in range(outputs):
        for r in range(restarts):
            th0 = np.random.default_rng([seed, o, r]).random(3 if o == 0 else 5)
            if o == 0:
                tasks.append({"X": None, "Z": None, "y": Y[0], "theta0": th0})
            else:
                X = np.ascontiguousarray(Y[:o].T)
                tasks.append({"X": X, "Z": np.ascontiguousarray(X[:: n // m][:m]), "y": Y[o], "theta0": th0})
    return t, tasks


def python_fit(ctx, t, tk, iterations):
    import gpar_at_scale_b200 as gp
    from gpar_at_scale_b200 import neldermead
    ctx.set_times(t); ctx.set_outputs(tk["y"]); ctx.set_noise_vector(None)
    if tk["X"] is None:
        f = lambda th: -ctx.lgssm_logpdf(gp.MATERN52, th)[0]
    else:
        ctx.set_inputs(tk["X"]); ctx.set_pseudo(tk["Z"])

        def f(th):
            try:
                return -ctx.scaled_dtc(gp.MATERN52, gp.MATERN52, th)
            except gp.PosDefException:
                return np.inf
    return neldermead.optimize(f, tk["theta0"], iterations=iterations)


def test_group_fit_walks_the_same_simplices_as_the_python_mirror(ctx):
    """gpar_group_fit (C++ Nelder-Mead inside the library) against neldermead.py over the single-context ABI:
    same objective values in the same order, so minimum, minimiser and evaluation count are identical."""
    import gpar_at_scale_b200 as gp
    t, tasks = chain_tasks(11)
    g = gp.Group([0])
    try:
        minimum, minimizer, calls, member = g.fit(t, tasks, gp.MATERN52, gp.MATERN52, iterations=12)
    finally:
        g.close()
    for k, tk in enumerate(tasks):
        res = python_fit(ctx, t, tk, 12)
        p = 3 if tk["X"] is None else 5
        assert calls[k] == res.f_calls and member[k] == 0
        assert minimum[k] == pytest.approx(res.minimum, rel=1e-13)
        assert np.allclose(minimizer[k, :p], res.minimizer, rtol=1e-13, atol=0) and np.all(np.isnan(minimizer[k, p:]))


def test_group_fit_lbfgs_matches_the_python_mirror(ctx):
    """L-BFGS inside the library (analytic gradients) against lbfgs.py over the single-context ABI: the same
    algorithm (dot products may round differently), so the minima agree tightly and are below Nelder-Mead's."""
    import gpar_at_scale_b200 as gp
    from gpar_at_scale_b200 import lbfgs
    t, tasks = chain_tasks(13, outputs=2, restarts=1)
    g = gp.Group([0])
    try:
        minimum, minimizer, calls, _ = g.fit(t, tasks, gp.MATERN52, gp.MATERN52, iterations=10, optimizer="lbfgs")
        nm_minimum = g.fit(t, tasks, gp.MATERN52, gp.MATERN52, iterations=10)[0]
    finally:
        g.close()
    for k, tk in enumerate(tasks):
        ctx.set_times(t); ctx.set_outputs(tk["y"]); ctx.set_noise_vector(None)
        if tk["X"] is None:
            def fg(th):
                v, gr = ctx.lgssm_logpdf_grad(gp.MATERN52, th)
                return -v[0], -gr[0]
        else:
            ctx.set_inputs(tk["X"]); ctx.set_pseudo(tk["Z"])

            def fg(th):
                v, gr = ctx.scaled_dtc_grad(gp.MATERN52, gp.MATERN52, th)
                return -v, -gr
        res = lbfgs.optimize(fg, tk["theta0"], iterations=10)
        assert minimum[k] == pytest.approx(res.minimum, rel=1e-7)
        assert abs(int(calls[k]) - res.f_calls) <= 2
        assert minimum[k] <= nm_minimum[k] + 1e-9 * abs(nm_minimum[k])


@pytest.mark.skipif(device_count() < 2, reason="needs two devices (gpurun --gpus 2)")
def test_group_of_two_devices():
    import gpar_at_scale_b200 as gp
    g = gp.Group([0, 1]); g1 = gp.Group([0])
    try:
        probs = [problem(21), problem(22, n=4100, m=33, d=3)]
        for mbr, pr in zip(g.members, probs):
            load(mbr, *pr)
        th5 = np.array([[0.2, 0.1, -0.3, 0.2, -1.0], [-0.1, 0.3, 0.1, -0.2, -0.7]])
        v, gr, codes = g.scaled_dtc(gp.MATERN52, gp.MATERN52, th5, grad=True)
        for i, pr in enumerate(probs):
            load(g1.members[0], *pr)
            v0, g0, _ = g1.scaled_dtc(gp.MATERN52, gp.MATERN52, th5[i:i + 1], grad=True)
            assert codes[i] == 0 and v[i] == v0[0] and np.array_equal(gr[i], g0[0])
        # chain: member 1 smooths, its means reach member 0's inputs over NCCL
        _, mean, _ = g.members[1].lgssm_smooth(gp.MATERN52, np.array([0.0, 0.0, -1.0]))
        n0 = probs[0][0].shape[0]
        back = g.broadcast(1, n=n0)
        assert np.array_equal(back, mean[0][:n0])
        g.members[0].set_inputs_column(1)
        X2 = probs[0][1].copy(); X2[:, 1] = mean[0][:n0]
        load(g1.members[0], probs[0][0], X2, probs[0][2], probs[0][3])
        assert g.members[0].scaled_dtc(gp.MATERN52, gp.MATERN52, th5[0]) == g1.members[0].scaled_dtc(gp.MATERN52, gp.MATERN52, th5[0])
        # merged chain: member 1 smooths a merged train+test problem, the N* test means (test order) reach member 0
        tA, XA, ZA, yA = probs[0]
        rs = np.random.default_rng(5); ns = 900
        ts = np.sort(rs.uniform(tA[0], tA[-1], ns)); XsA = rs.normal(size=(ns, XA.shape[1]))
        g.members[1].set_merged(tA, yA, ts, 0.04)
        g.members[1].lgssm_smooth(gp.MATERN52, np.array([0.0, 0.0, -1.0]), keep_on_device=True)
        mt, _ = g.members[1].take_test()
        assert np.array_equal(g.broadcast(1, n=ns), mt)
        par = np.array([1.3, 0.9, 1.1, 0.8, 0.3]); Wm = rs.normal(size=(ZA.shape[0], 4))
        g.members[0].set_pseudo(ZA); g.members[0].set_merged(tA, yA, ts, par[4] ** 2, X=XA, Xs=XsA); g.members[0].set_merged_test_column(0)
        g.members[0].scaled_predict(gp.MATERN52, gp.MATERN52, par, Wm, keep_on_device=True)
        XsB = XsA.copy(); XsB[:, 0] = mt
        g1.members[0].set_pseudo(ZA); g1.members[0].set_merged(tA, yA, ts, par[4] ** 2, X=XA, Xs=XsB)
        g1.members[0].scaled_predict(gp.MATERN52, gp.MATERN52, par, Wm, keep_on_device=True)
        assert np.array_equal(g.members[0].take_test()[0], g1.members[0].take_test()[0])
        # fits: dynamic hand-out over two members, same results as one member
        t, tasks = chain_tasks(12)
        a = g.fit(t, tasks, gp.MATERN52, gp.MATERN52, iterations=6)
        b = g1.fit(t, tasks, gp.MATERN52, gp.MATERN52, iterations=6)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1], equal_nan=True) and np.array_equal(a[2], b[2])
        assert set(a[3].tolist()) == {0, 1}
    finally:
        g.close(); g1.close()
