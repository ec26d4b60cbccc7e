"""Times gpar_scaled_dtc_grad at N = 1M, M = 1024 (device ms via gpar_last_timing) next to the value-only call."""
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
for it in range(4):
    t0 = time.perf_counter(); v = ctx.scaled_dtc(3, 3, th); t1 = time.perf_counter()
    ms, L = ctx.last_timing()
    print("value: %.2f ms wall, %.2f ms device, %d launches" % ((t1 - t0) * 1e3, ms, L))
for it in range(4):
    t0 = time.perf_counter(); v, g = ctx.scaled_dtc_grad(3, 3, th); t1 = time.perf_counter()
    ms, L = ctx.last_timing()
    print("value+grad: %.2f ms wall, %.2f ms device, %d launches" % ((t1 - t0) * 1e3, ms, L), g)
