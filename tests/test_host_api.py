"""CPU: host-side logic of the mirror of the reference's Julia interface (no GPU, no library calls)."""
import numpy as np
import pytest
import gpar_at_scale_b200 as gp
from gpar_at_scale_b200 import api, neldermead, chain, parallel
import toy_data as data


def test_to_colvecs_layout_matches_reference():
    # util.jl:16-31: a list of per-feature vectors -> D x N column-major == N records of D doubles
    X = api.to_ColVecs([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    assert X.shape == (3, 2) and X.flags["C_CONTIGUOUS"] and X.tolist() == [[1, 4], [2, 5], [3, 6]]
    assert api.to_ColVecs(np.arange(4.0)).shape == (4, 1)
    assert api.to_ColVecs([1.0, 2.0]).shape == (2, 1)
    Y = np.asfortranarray(np.arange(6.0).reshape(3, 2))
    assert api.to_ColVecs(Y).flags["C_CONTIGUOUS"]


def test_unpack_and_parse_params():
    assert api.unpack_gp([0.0, 0.0, 0.0]) == pytest.approx((1.001,) * 3)
    assert api.unpack_gpar(np.log(np.array([2.0, 3.0, 4.0, 5.0, 6.0]) - 1e-3)) == pytest.approx((2, 3, 4, 5, 6))
    rng = np.random.default_rng(0)
    p = api.parse_initial_gpar_params(None, 1.5, None, None, -3.0, rng=rng)      # util.jl:128-134: missing -> rand() in [0, 1)
    assert p[1] == 1.5 and p[4] == -3.0 and all(0 <= p[i] < 1 for i in (0, 2, 3))
    assert api.parse_initial_gp_params(0.1, 0.2, 0.3).tolist() == [0.1, 0.2, 0.3]


def test_masks():
    assert api.get_time_mask(3).tolist() == [1, 0, 0]
    assert api.get_output_mask(3).tolist() == [[0, 1, 0], [0, 0, 1]]
    with pytest.raises(ValueError):
        api.get_output_mask(1)


def test_kernel_objects_and_finite_gp():
    k = api.kernel(api.Matern52(), l=2.0, s=3.0)
    assert k.code == gp.MATERN52 and k.l == 2.0 and k.s == 3.0
    f = api.GP(k)(np.arange(5.0), 0.04)
    assert len(f) == 5 and f.noise == 0.04 and f.x.shape == (5, 1)
    with pytest.raises(ValueError):
        api.create_lgssm(np.arange(3.0), 1.0, 1.0, 0.1, api.EQ())      # EQ has no SDE form


def test_toy_data_shapes_follow_the_reference():
    # toy_data.jl:76-98: 10 000 samples minus 5 x 300 nuked minus the 4-point remainder -> 8 496; 100 000 true points
    x, y_obs, x_true, y_true = data.generate_big_dataset(np.random.default_rng(0))
    assert len(x) == 8496 and len(x_true) == 100000 and len(y_obs) == 3 and all(len(y) == 8496 for y in y_obs)
    assert x_true[-1] == pytest.approx(10000 / 30 + 50) and np.all(np.diff(x) > 0)
    gaps = np.diff(x)
    assert np.sum(gaps > 5) == 5                                    # five removed intervals
    x, y_obs, x_true, y_true = data.generate_small_dataset(np.random.default_rng(0))
    assert len(x) == 30 and len(x_true) == 1000
    # noise std is observation_noise^2 (toy_data.jl:29 quirk): 0.05^2 on the small set
    resid = y_obs[0] - data.f1_small(x)
    assert 0.3 * 0.0025 < np.std(resid) < 3 * 0.0025


def test_nelder_mead_matches_optim_defaults_on_rosenbrock():
    f = lambda v: (1 - v[0]) ** 2 + 100 * (v[1] - v[0] ** 2) ** 2
    res = neldermead.optimize(f, np.array([-1.2, 1.0]), iterations=2000)
    assert res.converged and res.minimum < 1e-8 and np.allclose(res.minimizer, [1, 1], atol=1e-3)   # g_tol = 1e-8 on the simplex spread
    # time limit and iteration budget are honoured
    res = neldermead.optimize(f, np.array([-1.2, 1.0]), iterations=5)
    assert res.iterations == 5 and not res.converged
    res = neldermead.optimize(lambda v: f(v), np.array([-1.2, 1.0]), time_limit=0.0)
    assert res.stopped_by_time


def test_chain_tasks_and_pseudo_inputs():
    tasks, costs = chain.make_tasks(8, 8)
    assert len(tasks) == 64 and tasks[0] == (0, 0) and costs[0] < costs[8] < costs[-1]
    X = np.arange(20.0).reshape(10, 2)
    Z = chain.strided_pseudo_inputs(X, 4)
    assert Z.shape == (4, 2) and Z[0].tolist() == X[0].tolist() and Z[-1].tolist() == X[-1].tolist()
    assert chain.strided_pseudo_inputs(X, 50).shape == (10, 2)
    best = parallel.best_per_output([(0, 0), (0, 1), (1, 0)], np.array([3.0, 2.0, 5.0]), np.zeros((3, 5)))
    assert best[0][0] == 2.0 and best[0][2] == 1 and best[1][0] == 5.0


def test_lbfgs_on_rosenbrock_and_failure_handling():
    from gpar_at_scale_b200 import lbfgs

    def fg(v):
        f = (1 - v[0]) ** 2 + 100 * (v[1] - v[0] ** 2) ** 2
        g = np.array([-2 * (1 - v[0]) - 400 * v[0] * (v[1] - v[0] ** 2), 200 * (v[1] - v[0] ** 2)])
        return f, g
    res = lbfgs.optimize(fg, np.array([-1.2, 1.0]), iterations=200, g_tol=1e-8, f_reltol=0.0)
    assert res.minimum < 1e-12 and np.allclose(res.minimizer, [1, 1], atol=1e-5) and res.f_calls < 120
    res = lbfgs.optimize(fg, np.array([-1.2, 1.0]), iterations=3)
    assert res.iterations == 3 and not res.converged

    # an objective that is +inf outside a box (a failed Cholesky reported as inf): the line search backs off
    def fg_box(v):
        if np.any(np.abs(v) > 2.0):
            return np.inf, np.zeros(2)
        return float(np.sum((v - 1.9) ** 2)), 2 * (v - 1.9)
    res = lbfgs.optimize(fg_box, np.array([-1.5, 0.0]), iterations=50)
    assert np.allclose(res.minimizer, [1.9, 1.9], atol=1e-4)


def test_nelder_mead_speculative_walks_the_same_simplices():
    """neldermead.optimize_speculative: the four candidate points of an iteration evaluated in one batched call, decisions
    taken with the values the sequential algorithm would have requested — identical minimiser, minimum, iterations and
    f_calls on Rosenbrock, a function with +inf regions, and a noisy quadratic in 5 dimensions."""
    from gpar_at_scale_b200 import neldermead
    fs = [lambda x: (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2,
          lambda x: np.inf if x[0] > 1.2 else float(np.sum((x - 0.5) ** 2) + np.sin(7 * x[1])),
          lambda x: float(np.sum((x - np.arange(5) / 5.0) ** 2 * (1 + np.arange(5))))]
    x0s = [np.array([-1.2, 1.0]), np.array([0.3, 0.9]), np.random.default_rng(0).random(5)]
    for f, x0 in zip(fs, x0s):
        calls = []
        a = neldermead.optimize(f, x0, iterations=150)
        b = neldermead.optimize_speculative(lambda P: (calls.append(len(P)), [f(p) for p in P])[1], x0, iterations=150)
        assert np.array_equal(a.minimizer, b.minimizer) and a.minimum == b.minimum
        assert a.iterations == b.iterations and a.f_calls == b.f_calls
        assert max(calls) <= max(4, len(x0) + 1) and len(calls) <= a.iterations + 8        # one batched call per iteration (+ shrinks)


def test_nelder_mead_lock_step_batch_equals_separate_runs():
    """neldermead.optimize_batch: several runs in lock-step on one batched objective — every run performs exactly the
    operations of a separate `optimize` (same minimiser, minimum, iterations, f_calls), including runs that converge
    early, +inf objective values and shrink steps; the batched objective sees at most (n + 1) points per run and round."""
    from gpar_at_scale_b200 import neldermead
    def f(x):
        if x[0] < -2.0:
            return np.inf
        return float((1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2 + 0.1 * np.sin(5 * x[2]) ** 2 + x[2] ** 2)
    rng = np.random.default_rng(4)
    X0 = rng.uniform(-1.5, 1.5, (7, 3)); X0[3] = [1.0, 1.0, 0.0]          # (one run starts at the optimum: converges at once)
    sizes = []
    def fbatch(P):
        sizes.append(len(P)); return [f(p) for p in P]
    runs = neldermead.optimize_batch(fbatch, X0, iterations=80)
    for k in range(7):
        one = neldermead.optimize(f, X0[k], iterations=80)
        assert np.array_equal(runs[k].minimizer, one.minimizer) and runs[k].minimum == one.minimum
        assert runs[k].iterations == one.iterations and runs[k].f_calls == one.f_calls and runs[k].converged == one.converged
    assert max(sizes) <= 7 * 4 and sizes[0] == 7 * 4


def test_row_sharded_fit_on_a_stand_in_group():
    """api.get_optim_scaled_gpar_params(group=...) is host logic over two Group methods (load_row_slices, scaled_dtc_sharded):
    with a stand-in group whose 'objective' is a quadratic, Nelder-Mead and L-BFGS reach its optimum, the slices are loaded
    once, and the options that need the batched single-device entry points are refused before anything is touched."""
    class FakeGroup:
        def __init__(self):
            self.loaded = 0; self.evals = 0
        def load_row_slices(self, X, Z, t, y):
            self.loaded += 1; self.shapes = (X.shape, Z.shape, len(t), len(y))
            return np.array([0, 1024], dtype=np.int64)
        def scaled_dtc_sharded(self, k_time, k_out, theta, row_lo, grad=False):
            self.evals += 1
            assert k_time == gp.MATERN52 and k_out == gp.MATERN52 and list(row_lo) == [0, 1024]
            d = np.asarray(theta) - np.array([0.5, -0.2, 0.1, 0.3, -1.0])
            v = -float(d @ d)                                   # the library maximises the objective: nlml = -v
            return (v, -2.0 * d) if grad else v
    n = 2000
    X = np.zeros((n, 2)); Z = np.zeros((7, 2)); t = np.arange(n, dtype=float); y = np.zeros(n)
    g = FakeGroup()
    p, r = api.get_optim_scaled_gpar_params(X, Z, t, y, group=g, iterations=400, return_result=True,
                                            i_log_time_l=0.0, i_log_time_var=0.0, i_log_out_l=0.0, i_log_out_var=0.0, i_log_noise_sigma=0.0)
    assert g.loaded == 1 and g.shapes == ((n, 2), (7, 2), n, n) and g.evals == r.f_calls
    assert np.allclose(r.minimizer, [0.5, -0.2, 0.1, 0.3, -1.0], atol=1e-4) and p == pytest.approx(api.unpack_gpar(r.minimizer))
    g2 = FakeGroup()
    _, r2 = api.get_optim_scaled_gpar_params(X, Z, t, y, group=g2, optimizer="lbfgs", iterations=50, return_result=True,
                                             i_log_time_l=0.0, i_log_time_var=0.0, i_log_out_l=0.0, i_log_out_var=0.0, i_log_noise_sigma=0.0)
    assert np.allclose(r2.minimizer, [0.5, -0.2, 0.1, 0.3, -1.0], atol=1e-6)
    for bad in (dict(n_restarts=4), dict(speculative=True)):
        g3 = FakeGroup()
        with pytest.raises(ValueError):
            api.get_optim_scaled_gpar_params(X, Z, t, y, group=g3, **bad)
        assert g3.loaded == 0
    assert parallel.row_slice_bounds(10_000, 4) == [0, 3072, 5120, 8192, 10_000]
    assert parallel.row_slice_bounds(100, 3, align=4) == [0, 36, 68, 100]
