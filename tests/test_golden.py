"""CPU: the oracle reproduces the committed golden fixtures (guards the checker against drift)."""
import os
import numpy as np
import pytest
import oracle
from oracle import cport
from oracle.grad import dtc_diag_value_and_grad

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_dtc_golden():
    z = np.load(os.path.join(G, "dtc.npz"))
    for i in range(int(z["ncases"])):
        kind, vfe, jit = z[f"c{i}_meta"]
        val, grad = dtc_diag_value_and_grad(z[f"c{i}_theta"], z[f"c{i}_X"], z[f"c{i}_Z"], z[f"c{i}_y"], int(kind), bool(vfe), float(jit))
        assert val == pytest.approx(float(z[f"c{i}_val"]), rel=1e-10)
        assert np.allclose(grad, z[f"c{i}_grad"], rtol=1e-8, atol=1e-10)


def test_scaled_golden_with_c_port():
    z = np.load(os.path.join(G, "scaled.npz"))
    for i in range(int(z["ncases"])):
        kt, ko = (int(v) for v in z[f"c{i}_meta"])
        th = z[f"c{i}_theta"]
        tl, tv, ol, ov, ns = oracle.unpack_gpar(th)
        X, Z, t, y = z[f"c{i}_X"], z[f"c{i}_Z"], z[f"c{i}_t"], z[f"c{i}_y"]
        Cfu = oracle.pairwise(ko, X, Z, ol, ov ** 2)
        cu = oracle.pairwise(ko, Z, Z, ol, ov ** 2) + ns ** 2 * np.eye(len(Z))
        dtc, A = oracle.compute_gpar_dtc_objective(Cfu, cu, t, y, kt, tl, tv ** 2, ns ** 2, dense_logdet=False,
                                                   decorrelate=cport.kalman_decorrelate)
        assert dtc == pytest.approx(float(z[f"c{i}_dtc"]), rel=1e-10)
        assert np.allclose(A, z[f"c{i}_A"], atol=1e-9)


def test_lgssm_golden_with_c_port():
    z = np.load(os.path.join(G, "lgssm.npz"))
    for i in range(int(z["ncases"])):
        kind = int(z[f"c{i}_kind"]); t = z[f"c{i}_t"]; Y = z[f"c{i}_Y"]
        l, var, sig = oracle.unpack_gp(z[f"c{i}_theta"])
        rv = z[f"c{i}_rvec"] if f"c{i}_rvec" in z else None
        lml, alpha = cport.kalman_filter_batch(kind, t, Y, l, var ** 2, sig ** 2, rvec=rv, want_alpha=True)
        assert np.allclose(lml, z[f"c{i}_lml"], rtol=1e-12)
        assert np.allclose(alpha, z[f"c{i}_alpha"], atol=1e-11)
        _, mean, v = cport.kalman_smooth_batch(kind, t, Y, l, var ** 2, rv if rv is not None else sig ** 2)
        assert np.allclose(mean, z[f"c{i}_mean"], atol=1e-11) and np.allclose(v, z[f"c{i}_var"], rtol=1e-9)
