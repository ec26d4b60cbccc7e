"""Chunk-length sweep of the general Kalman path (GPAR_KF_L) on cfg 3 (1024 x 10k, own model each) and 1 x 10M, irregular grid."""
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
            fn(); o.append(ctx.last_timing()[0])
        return float(np.median(o[skip:]))
    ctx.set_times(tk); ctx.set_outputs(Yk)
    a = med(lambda: ctx.lgssm_logpdf(3, ths))
    b = med(lambda: ctx.lgssm_smooth(3, ths[0], keep_on_device=True), 5, 2)
    N10 = 10_000_000
    ctx.set_outputs(rng.normal(size=N10)); ctx.set_times(np.cumsum(rng.exponential(1 / 30, N10)))
    c = med(lambda: ctx.lgssm_logpdf(3, np.log([1.0, 1.0, 0.1])))
    print(json.dumps({"L": os.environ.get("GPAR_KF_L", "default"), "cfg3_filter_ms": a, "cfg3_smooth_shared_ms": b, "1x10M_irregular_ms": c}))
else:
    for L in ["default", "16", "32", "64", "128", "256"]:
        env = dict(os.environ)
        if L != "default":
            env["GPAR_KF_L"] = L
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print(p.stdout.strip() or p.stderr[-500:])
