"""One-pass log-pdf (kf_chunk_element_kernel): variants x chunk lengths on cfg 3 (1024 x 10k, own model each) and 1 x 10M, irregular grid."""
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
    def med(fn, n=9, skip=3):
        o = []
        for _ in range(n):
            fn(); o.append(ctx.last_timing()[0])
        return float(np.median(o[skip:])), int(ctx.last_timing()[1])
    ctx.set_times(tk); ctx.set_outputs(Yk)
    a, la = med(lambda: ctx.lgssm_logpdf(3, ths))
    N10 = 10_000_000
    ctx.set_outputs(rng.normal(size=N10)); ctx.set_times(np.cumsum(rng.exponential(1 / 30, N10)))
    c, lc = med(lambda: ctx.lgssm_logpdf(3, np.log([1.0, 1.0, 0.1])))
    print(json.dumps({"onepass": os.environ.get("GPAR_KF_ONEPASS", "1"), "variant": os.environ.get("GPAR_KF1_VARIANT", "0"), "L": os.environ.get("GPAR_KF_L", "auto"), "pf": os.environ.get("GPAR_KF1_PREFETCH"),
                      "cfg3_filter_ms": round(a, 4), "cfg3_launches": la, "1x10M_irregular_ms": round(c, 4), "1x10M_launches": lc}))
else:
    runs = [("1", v, pf) for v in ("0", "1", "1") for pf in ("0",)]
    for op, v, L in runs:
        env = dict(os.environ); env["GPAR_KF_ONEPASS"] = op; env["GPAR_KF1_VARIANT"] = v
        env["GPAR_KF1_PREFETCH"] = L
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print(p.stdout.strip() or p.stderr[-800:], flush=True)
