"""Single-pass steady-state Kalman (regular grid): variants of the non-head pass, timed by the library's own
CUDA events (ctx.last_timing) on 1 x 10M and 8 x 10M; lml / alpha compared with variant 0 and the general scan."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp

rng = np.random.default_rng(2)
ctx = gp.Context(0)
N = 10_000_000
th = np.log([1.0, 1.0, 0.1])
for batch in (1, 8):
    Y = rng.normal(size=(batch, N))
    ctx.set_outputs(Y); ctx.set_times_range(0.0, 1 / 30, N)
    ref = None
    for var in ("general", "0", "1", "2", "3"):
        if var == "general":
            os.environ["GPAR_KF_STEADY"] = "0"; os.environ["GPAR_FILTER_SHARED"] = "0"
        else:
            os.environ.pop("GPAR_KF_STEADY", None); os.environ["GPAR_FILTER_SHARED"] = "0"; os.environ["GPAR_SS3_VARIANT"] = var
        ths = np.tile(th, (batch, 1))
        ms = []
        for i in range(7):
            v = ctx.lgssm_logpdf(3, ths); ms.append(ctx.last_timing()[0])
        if batch == 1:
            msd = []
            for i in range(4):
                _, a = ctx.lgssm_decorrelate(3, th); msd.append(ctx.last_timing()[0])
        else:
            a = None; msd = [float("nan")]
        if ref is None:
            ref = (v.copy(), a)
        da = 0.0 if a is None else float(np.max(np.abs(a - ref[1])))
        med = float(np.median(ms[2:]))
        print("batch %d variant %-7s logpdf best %.1f us median %.1f us (%.2f TB/s of y) decorrelate %.1f us | rel dlml %.2e  max dalpha %.2e"
              % (batch, var, min(ms) * 1e3, med * 1e3, batch * N * 8 / (med * 1e-3) / 1e12, min(msd) * 1e3,
                 float(np.max(np.abs(v - ref[0]) / np.abs(ref[0]))), da), flush=True)

# ticket-fused finish (bit 0) and programmatic dependent launch of the main pass (bit 1), default variant per batch
os.environ.pop("GPAR_SS3_VARIANT", None)
for batch in (1, 8, 16):
    Y = rng.normal(size=(batch, N))
    ctx.set_outputs(Y); ctx.set_times_range(0.0, 1 / 30, N)
    ths = np.tile(th, (batch, 1))
    ref = None
    for fused in ("0", "1", "3"):
        os.environ["GPAR_SS3_FUSED"] = fused
        ms = []
        for i in range(9):
            v = ctx.lgssm_logpdf(3, ths); ms.append(ctx.last_timing()[0])
        if ref is None:
            ref = v.copy()
        med = float(np.median(ms[2:]))
        print("batch %2d fused %s logpdf best %.1f us median %.1f us (%.2f TB/s of y) launches %d | max rel dlml %.2e"
              % (batch, fused, min(ms) * 1e3, med * 1e3, batch * N * 8 / (med * 1e-3) / 1e12, ctx.last_timing()[1],
                 float(np.max(np.abs(v - ref) / np.abs(ref)))), flush=True)
os.environ.pop("GPAR_SS3_FUSED", None)
