"""CPU: the C++ optimisers that gpar_group_fit runs on every device thread (csrc/optim_host.h) against the Python host
mirror (neldermead.py, lbfgs.py — the restatements of Optim.jl's NelderMead defaults and of an L-BFGS with backtracking)
on analytic functions, including one that returns +inf (what a failed Cholesky becomes)."""
import os
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "optim_check")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", EXE, os.path.join(ROOT, "tools", "optim_check.cpp")])
    return EXE


def fval(name, x):
    x = np.asarray(x, dtype=np.float64)
    if name == "rosenbrock":
        a, b = 1.0 - x[0], x[1] - x[0] * x[0]
        return a * a + 100.0 * b * b, np.array([-2.0 * a - 400.0 * x[0] * b, 200.0 * b])
    if name == "quadratic5":
        w = np.arange(1, x.size + 1, dtype=np.float64)
        d = x - 0.1 * w
        f = 0.0
        for i in range(x.size):
            f += w[i] * d[i] * d[i]
        return f, 2.0 * w * d
    if x[0] <= -0.5:
        return np.inf, np.zeros_like(x)
    f = 0.0
    for i in range(x.size):
        f += (x[i] + 0.4) * (x[i] + 0.4)
    return f, 2.0 * (x + 0.4)


def run(exe, alg, fn, it, x0):
    out = subprocess.run([exe, alg, fn, str(it)] + ["%.17g" % v for v in x0], capture_output=True, text=True, check=True).stdout.split()
    return float(out[0]), int(out[1]), np.array([float(v) for v in out[2:]])


@pytest.mark.parametrize("fn,x0,it", [("rosenbrock", [-1.2, 1.0], 200), ("rosenbrock", [0.3, 0.7], 15), ("quadratic5", [0.9, 0.1, 0.5, 0.3, 0.7], 60),
                                      ("wall", [0.3, 0.2, 0.1], 50), ("wall", [-0.45, 0.2, 0.1], 30)])
def test_nelder_mead_twin_walks_the_same_simplices(exe, fn, x0, it):
    from gpar_at_scale_b200 import neldermead
    res = neldermead.optimize(lambda x: fval(fn, x)[0], np.array(x0), iterations=it)
    fb, calls, xb = run(exe, "nm", fn, it, x0)
    assert calls == res.f_calls
    assert fb == pytest.approx(res.minimum, rel=1e-13, abs=1e-300)
    assert np.allclose(xb, res.minimizer, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("fn,x0,it", [("rosenbrock", [-1.2, 1.0], 100), ("quadratic5", [0.9, 0.1, 0.5, 0.3, 0.7], 40), ("wall", [0.3, 0.2, 0.1], 30),
                                      ("wall", [-0.6, 0.2, 0.1], 10)])
def test_lbfgs_twin_matches_the_python_mirror(exe, fn, x0, it):
    from gpar_at_scale_b200 import lbfgs
    res = lbfgs.optimize(lambda x: fval(fn, x), np.array(x0), iterations=it)
    fb, calls, xb = run(exe, "lbfgs", fn, it, x0)
    if not np.isfinite(res.minimum):          # started outside the domain: both give up at once
        assert not np.isfinite(fb) and calls == res.f_calls == 1
        return
    assert abs(calls - res.f_calls) <= 1
    assert fb == pytest.approx(res.minimum, rel=1e-6, abs=1e-18)
    assert np.allclose(xb, res.minimizer, rtol=1e-6, atol=1e-9)
