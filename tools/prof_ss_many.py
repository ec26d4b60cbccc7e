"""1024 x 10k (and 256 x 40k) regular grid: two-pass steady-state path vs the single-pass scheme with relaxed guards."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(2)
ctx = gp.Context(0)
for B, NK in ((1024, 10000), (256, 40000), (64, 160000)):
    Y = rng.normal(size=(B, NK))
    ths = np.stack([np.log(rng.uniform(0.05, 5, B)), np.log(rng.uniform(0.3, 3, B)), np.log(rng.uniform(0.01, 1, B))], axis=1)
    ctx.set_outputs(Y); ctx.set_times_range(0.0, 1 / 30, NK)
    ref = None
    for mode in ("general", "two-pass", "single-pass"):
        for k in ("GPAR_KF_STEADY", "GPAR_SS_LONG_MIN_N", "GPAR_SS_LONG_MAX_BATCH"):
            os.environ.pop(k, None)
        if mode == "general":
            os.environ["GPAR_KF_STEADY"] = "0"
        elif mode == "single-pass":
            os.environ["GPAR_SS_LONG_MIN_N"] = "1"; os.environ["GPAR_SS_LONG_MAX_BATCH"] = "100000"
        ms = []
        for i in range(7):
            v = ctx.lgssm_logpdf(3, ths); ms.append(ctx.last_timing()[0])
        if ref is None:
            ref = v.copy()
        print("%d x %d %-11s median %.1f us (%.1f G steps/s) launches %d  max rel dlml %.2e" % (B, NK, mode, np.median(ms[2:]) * 1e3, B * NK / np.median(ms[2:]) / 1e6, ctx.last_timing()[1], np.max(np.abs(v - ref) / np.abs(ref))), flush=True)
