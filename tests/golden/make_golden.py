"""Generates tests/golden/*.npz: seeded inputs + oracle outputs for the hot-path entry points.

The reference cannot be executed here (no Julia) and ships no golden vectors (its data generators
use Julia's unseeded global RNG, src/data/toy_data.jl:34-36), so these fixtures come from the CPU
oracle (`oracle/`, parity unpinned — see oracle/__init__.py).  They guard the oracle against drift
(tests/test_golden.py, CPU) and are what the GPU parity tests compare against on the B200 box.

    python tests/golden/make_golden.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle                                   # noqa: E402
from oracle.grad import dtc_diag_value_and_grad  # noqa: E402
from oracle.dtc import scaled_gpar_objective     # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def toy_small(rng, n=30):
    x = np.linspace(0.0, n / 30.0, n)
    nz = lambda: rng.normal(0.0, 0.05 ** 2, n)
    y1 = -np.sin(10 * np.pi * (x + 1)) / (2 * x + 1) - x ** 4 + nz()
    y2 = np.cos(y1) ** 2 + np.sin(3 * x) + nz()
    y3 = y2 * y1 ** 2 + 3 * x + nz()
    return x, y1, y2, y3


def main():
    rng = np.random.default_rng(20201018)
    # ---- plain DTC / VFE (+ gradient) -----------------------------------------------------
    cases = {}
    for i, (kind, D, vfe, jit) in enumerate([(3, 1, 0, -1.0), (0, 2, 1, 1e-4), (2, 3, 0, -1.0), (1, 1, 1, -1.0)]):
        n, m = 257, 19
        X = rng.normal(size=(n, D)) * 1.5; Z = rng.normal(size=(m, D)) * 1.5; y = rng.normal(size=n)
        th = rng.uniform(-0.8, 0.4, 3)
        val, grad = dtc_diag_value_and_grad(th, X, Z, y, kind, bool(vfe), jit)
        cases.update({f"c{i}_X": X, f"c{i}_Z": Z, f"c{i}_y": y, f"c{i}_theta": th, f"c{i}_meta": np.array([kind, vfe, jit]),
                      f"c{i}_val": val, f"c{i}_grad": grad})
    np.savez(os.path.join(OUT, "dtc.npz"), ncases=4, **cases)

    # ---- scaled GPAR objective + q_u: the dtc_example.jl:8-64 protocol and two more ----------
    cases = {}
    x, y1, y2, y3 = toy_small(rng)
    specs = [(x, y1[:, None], y1[::3][:, None], y2, 3, 3, np.log(np.array([1.0, 1.0, 1.0, 1.0, 0.05]) - 1e-3))]
    n, m = 300, 25
    t = np.sort(rng.uniform(0, 10, n))
    specs.append((t, rng.normal(size=(n, 2)), rng.normal(size=(m, 2)), rng.normal(size=n), 3, 3, rng.uniform(-0.8, 0.3, 5)))
    specs.append((t, rng.normal(size=(n, 3)), rng.normal(size=(m, 3)), rng.normal(size=n), 2, 0, rng.uniform(-0.8, 0.3, 5)))
    for i, (t_, X, Z, y, kt, ko, th) in enumerate(specs):
        tl, tv, ol, ov, ns = oracle.unpack_gpar(th)
        Cfu = oracle.pairwise(ko, X, Z, ol, ov ** 2)
        cu = oracle.pairwise(ko, Z, Z, ol, ov ** 2) + ns ** 2 * np.eye(len(Z))
        dtc, A = oracle.compute_gpar_dtc_objective(Cfu, cu, t_, y, kt, tl, tv ** 2, ns ** 2, dense_logdet=True)
        cases.update({f"c{i}_t": t_, f"c{i}_X": X, f"c{i}_Z": Z, f"c{i}_y": y, f"c{i}_theta": th, f"c{i}_meta": np.array([kt, ko]),
                      f"c{i}_dtc": dtc, f"c{i}_A": A})
        if ko != 0:   # bare Cuu of an EQ kernel is not numerically PD here (Julia would throw PosDefException)
            m_e, Dinv, U_u = oracle.compute_q_u(Cfu, oracle.pairwise(ko, Z, Z, ol, ov ** 2), t_, y, kt, tl, tv ** 2, ns ** 2)
            cases.update({f"c{i}_m_e": m_e, f"c{i}_Dinv": Dinv, f"c{i}_U_u": U_u})
    np.savez(os.path.join(OUT, "scaled.npz"), ncases=len(specs), **cases)

    # ---- LGSSM logpdf / decorrelate / smooth ----------------------------------------------
    cases = {}
    i = 0
    for kind in (1, 2, 3):
        for usevec in (0, 1):
            n, b = 131, 3
            t = np.cumsum(rng.exponential(1 / 30, n)) if usevec else np.arange(n) / 30.0
            Y = rng.normal(size=(b, n))
            th = rng.uniform(-1.2, 0.4, 3)
            l, var, sig = oracle.unpack_gp(th)
            rv = np.where(rng.uniform(size=n) < 0.15, 1e10, sig ** 2) if usevec else None
            noise = rv if usevec else sig ** 2
            lml = np.zeros(b); alpha = np.zeros((b, n)); mean = np.zeros((b, n)); var_ = np.zeros((b, n))
            for j in range(b):
                lml[j], alpha[j] = oracle.kalman_decorrelate(kind, t, Y[j], l, var ** 2, noise)
                _, mean[j], var_[j] = oracle.kalman_smooth(kind, t, Y[j], l, var ** 2, noise)
            cases.update({f"c{i}_t": t, f"c{i}_Y": Y, f"c{i}_theta": th, f"c{i}_kind": kind, f"c{i}_lml": lml, f"c{i}_alpha": alpha,
                          f"c{i}_mean": mean, f"c{i}_var": var_})
            if usevec:
                cases[f"c{i}_rvec"] = rv
            i += 1
    np.savez(os.path.join(OUT, "lgssm.npz"), ncases=i, **cases)

    # ---- exact GP / GPAR ------------------------------------------------------------------
    cases = {}
    x, y1, y2, y3 = toy_small(rng)
    xs = np.linspace(0, 1.1, 50)
    # GP (ntheta=3): EQ on time
    th3 = np.array([np.log(0.2), np.log(1.0), np.log(0.05)])
    l, var, sig = oracle.unpack_gp(th3)
    K = oracle.pairwise(0, x[:, None], x[:, None], l, var ** 2)
    Ksf = oracle.pairwise(0, xs[:, None], x[:, None], l, var ** 2)
    lml = [oracle.exact_logpdf(K, sig ** 2, yy) for yy in (y1, y2)]
    pm = [oracle.exact_posterior(K, Ksf, np.full(50, var ** 2), sig ** 2, yy) for yy in (y1, y2)]
    cases.update(dict(gp_X=x[:, None], gp_Y=np.stack([y1, y2]), gp_theta=th3, gp_Xs=xs[:, None], gp_lml=np.array(lml),
                      gp_mean=np.stack([p[0] for p in pm]), gp_var=pm[0][1]))
    # GPAR (ntheta=5): EQ time + Matern52 outputs, inputs (x, y1, y2) -> y3
    th5 = np.array([np.log(0.3), 0.1, np.log(1.5), -0.2, np.log(0.05)])
    tl, tv, ol, ov, ns = oracle.unpack_gpar(th5)
    X3 = np.stack([x, y1, y2], axis=1)
    Xs3 = np.stack([xs, np.sin(xs), np.cos(xs)], axis=1)
    K = oracle.gpar_kernel_matrix(0, 3, X3, X3, tl, tv, ol, ov)
    Ksf = oracle.gpar_kernel_matrix(0, 3, Xs3, X3, tl, tv, ol, ov)
    pm = oracle.exact_posterior(K, Ksf, np.full(50, tv ** 2 + ov ** 2), ns ** 2, y3)
    cases.update(dict(gpar_X=X3, gpar_y=y3, gpar_theta=th5, gpar_Xs=Xs3, gpar_lml=oracle.exact_logpdf(K, ns ** 2, y3),
                      gpar_mean=pm[0], gpar_var=pm[1]))
    np.savez(os.path.join(OUT, "exact.npz"), **cases)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
