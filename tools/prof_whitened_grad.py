"""Times gpar_scaled_dtc_grad at N = 1M, M = 1024 in its three forms (collapsed analytic, whitened coordinates, value only with
the whitened panel) — device ms via gpar_last_timing — and checks that the two analytic forms agree."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpar_at_scale_b200 as gp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rng = np.random.default_rng(0)
t = np.arange(N) / 30.0
x = rng.uniform(0, 100, N); z = np.linspace(0, 100, M)
y = np.sin(x) + 0.5 * np.sin(0.05 * t) + 0.1 * rng.normal(size=N)
th = np.log([2.0, 0.5, 1.0, 1.0, 0.1])
ctx = gp.Context(0)
ctx.set_inputs(x); ctx.set_pseudo(z); ctx.set_times(t); ctx.set_outputs(y)
def run(label, f, reps=3):
    for it in range(reps):
        t0 = time.perf_counter(); r = f(); t1 = time.perf_counter()
        ms, L = ctx.last_timing()
    print("%-34s %.2f ms wall, %.2f ms device, %d launches" % (label, (t1 - t0) * 1e3, ms, L), flush=True)
    return r
run("value (collapsed)", lambda: ctx.scaled_dtc(3, 3, th))
os.environ["GPAR_ROBUST_COND"] = "0"
run("value (whitened panel)", lambda: ctx.scaled_dtc(3, 3, th))
del os.environ["GPAR_ROBUST_COND"]
va, ga = run("value + gradient (collapsed)", lambda: ctx.scaled_dtc_grad(3, 3, th))
os.environ["GPAR_GRAD_WHITENED"] = "1"
vw, gw = run("value + gradient (whitened)", lambda: ctx.scaled_dtc_grad(3, 3, th))
del os.environ["GPAR_GRAD_WHITENED"]
print("values", va, vw, "gradient difference %.2e" % (np.max(np.abs(gw - ga)) / np.max(np.abs(ga))))
print(ga); print(gw)
# plain DTC with gradient
th3 = np.log([1.0, 1.0, 0.1])
va, ga = run("plain DTC + gradient (collapsed)", lambda: ctx.dtc_logpdf(3, th3, grad=True))
os.environ["GPAR_GRAD_WHITENED"] = "1"
vw, gw = run("plain DTC + gradient (whitened)", lambda: ctx.dtc_logpdf(3, th3, grad=True))
del os.environ["GPAR_GRAD_WHITENED"]
print("values", va, vw, "gradient difference %.2e" % (np.max(np.abs(gw - ga)) / np.max(np.abs(ga))))
