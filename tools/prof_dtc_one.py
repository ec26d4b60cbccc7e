"""Two gpar_dtc_logpdf(grad) calls at N = 1M, M = 1024 (BASELINE configs[1]) — for ncu captures of the SYRK / producer."""
import sys
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
rng = np.random.default_rng(1)
N, M = 1_000_000, 1024
x = rng.uniform(0, 100, N); z = np.linspace(x.min(), x.max(), M)
y = np.sin(x) + 0.3 * np.cos(3.1 * x) + 0.1 * rng.normal(size=N)
ctx = gp.Context(0)
ctx.set_inputs(x); ctx.set_pseudo(z); ctx.set_outputs(y)
for i in range(2):
    print(ctx.dtc_logpdf(3, np.log([1.0, 1.0, 0.1]), grad=True), ctx.last_timing())
