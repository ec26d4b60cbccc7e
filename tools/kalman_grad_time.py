"""LGSSM log-pdf + gradient on cfg 3 (1024 x 10k, own model each) and 1 x 10M: one-pass (default) vs three-phase (GPAR_KF_ONEPASS=0)."""
import os, sys, subprocess, json
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import gpar_at_scale_b200 as gp
    rng = np.random.default_rng(2)
    ctx = gp.Context(0)
    B, NK = 1024, 10000
    tk = np.cumsum(rng.exponential(1 / 30, NK)); Yk = rng.normal(size=(B, NK))
    ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
    def med(fn, n=7, skip=3):
        o = []
        for _ in range(n):
            r = fn(); o.append(ctx.last_timing()[0])
        return float(np.median(o[skip:])), int(ctx.last_timing()[1]), r
    ctx.set_times(tk); ctx.set_outputs(Yk)
    a, la, ra = med(lambda: ctx.lgssm_logpdf_grad(3, ths))
    v, lv, rv = med(lambda: ctx.lgssm_logpdf(3, ths))
    N10 = 10_000_000
    ctx.set_outputs(rng.normal(size=N10)); ctx.set_times(np.cumsum(rng.exponential(1 / 30, N10)))
    c, lc, rc = med(lambda: ctx.lgssm_logpdf_grad(3, np.log([1.0, 1.0, 0.1])))
    print(json.dumps({"onepass": os.environ.get("GPAR_KF_ONEPASS", "1"), "variant": os.environ.get("GPAR_KF1_VARIANT", "-"), "cfg3_grad_ms": round(a, 4), "cfg3_value_ms": round(v, 4), "cfg3_launches": la,
                      "1x10M_grad_ms": round(c, 4), "1x10M_launches": lc, "g0": [float(x) for x in np.ravel(ra[1])[:3]], "g10": [float(x) for x in np.ravel(rc[1])[:3]]}))
else:
    for op, v in (("1", None),):
        env = dict(os.environ); env["GPAR_KF_ONEPASS"] = op
        if v: env["GPAR_KF1_VARIANT"] = v
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print(p.stdout.strip() or p.stderr[-800:], flush=True)
