"""CPU: the host-side work planner of the DMMA SYRK (csrc/syrk_plan.h) covers every
(tile-job, k-block) exactly once and balances the cost over the SMs, for ragged shapes."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "syrk_plan_check")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", EXE, os.path.join(ROOT, "tools", "syrk_plan_check.cpp")])
    return EXE


@pytest.mark.parametrize("T,NBK,with_h", [(8, 31250, 1), (8, 31250, 0), (16, 65536, 1), (16, 65536, 0), (1, 1, 0), (1, 1, 1), (1, 7, 1),
                                          (2, 3, 0), (3, 1000, 1), (2, 266, 0), (1, 31250, 0), (5, 149, 1), (8, 4, 1), (1, 148, 0), (1, 147, 1)])
def test_plan_covers_and_balances(exe, T, NBK, with_h):
    out = subprocess.run([exe, str(T), str(NBK), str(with_h), "148"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    tag, C, nseg, imbalance, aligned = out.stdout.split()
    assert tag == "ok" and 1 <= int(C) <= 148
    njobs = T * (T + 1) // 2 + (T * T if with_h else 0)
    if njobs * NBK >= 148 * 64:          # enough work: every SM busy, max load within 1.5 % of the mean
        assert int(C) == 148 and float(imbalance) <= 1.015, out.stdout


def test_headline_shape_is_phase_aligned(exe):
    # N = 1M, M = 1024 with gradient: the 92 regular jobs' first pieces share one k range
    out = subprocess.run([exe, "8", "31250", "1", "148"], capture_output=True, text=True).stdout.split()
    assert float(out[4]) >= 0.55


@pytest.mark.parametrize("T,NBK,with_h,ctas", [(1, 266, 0, 16), (1, 266, 1, 33), (1, 5, 0, 1), (2, 40, 1, 17), (1, 20, 0, 1), (2, 9, 0, 1), (3, 100, 1, 93)])
def test_plan_with_a_capped_cta_count(exe, T, NBK, with_h, ctas):
    """Small problems cap the CTA count (panel_syrk_run: at least 16 k-blocks per CTA): coverage must still be exact."""
    out = subprocess.run([exe, str(T), str(NBK), str(with_h), str(ctas)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    tag, C, nseg, imbalance, aligned = out.stdout.split()
    assert tag == "ok" and 1 <= int(C) <= ctas
