"""Per-evaluation latency at the reference's own problem sizes (N = 8 496, M = 50 / 81; EEG N = M = 156)."""
import sys, os, time
import numpy as np
sys.path.insert(0, ".")
import gpar_at_scale_b200 as gp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples'))
import toy_data as data
rng = np.random.default_rng(0)
x, y_obs, x_true, y_true = data.generate_big_dataset(rng, true_samples=1000)
ctx = gp.Context(0)
th5 = np.log(np.array([1.0, 1.0, 1.0, 1.0, 0.6])); th3 = np.log(np.array([1.0, 1.0, 0.6]))
def bench(f, n=200):
    for _ in range(5): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
ctx.set_times(x); ctx.set_outputs(y_obs[0])
print("lgssm_logpdf N=8496: %.3f ms/eval (device %.3f)" % (bench(lambda: ctx.lgssm_logpdf(3, th3)), ctx.last_timing()[0]))
X = y_obs[0][:, None]; Z = np.linspace(X.min(), X.max(), 50)[:, None]
ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y_obs[1])
print("scaled_dtc N=8496 M=50 D=1: %.3f ms/eval (device %.3f, launches %d)" % (bench(lambda: ctx.scaled_dtc(3, 3, th5)), *ctx.last_timing()))
X = np.stack(y_obs[:2], axis=1); d1 = np.linspace(X[:, 0].min(), X[:, 0].max(), 9); d2 = np.linspace(X[:, 1].min(), X[:, 1].max(), 9)
Z = np.array([[a, b] for b in d2 for a in d1])
ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y_obs[2])
print("scaled_dtc N=8496 M=81 D=2: %.3f ms/eval (device %.3f, launches %d)" % (bench(lambda: ctx.scaled_dtc(3, 3, th5)), *ctx.last_timing()))
Xe = rng.normal(size=(156, 5)); ctx.set_inputs(Xe); ctx.set_pseudo(Xe); ctx.set_times(np.arange(156) / 256.0); ctx.set_outputs(rng.normal(size=156))
print("scaled_dtc EEG N=M=156 D=5: %.3f ms/eval (device %.3f)" % (bench(lambda: ctx.scaled_dtc(3, 3, th5)), ctx.last_timing()[0]))
print("exact_logpdf GPAR N=156 D=5: %.3f ms/eval (device %.3f)" % (bench(lambda: ctx.exact_logpdf(3, 3, th5)), ctx.last_timing()[0]))
# batched candidates (gpar_scaled_dtc_batch) at the reference's own size
X = y_obs[0][:, None]; Z = np.linspace(X.min(), X.max(), 50)[:, None]
ctx.set_times(x); ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y_obs[1])
ths = np.tile(th5, (64, 1)) + 0.05 * rng.normal(size=(64, 5))
ms = bench(lambda: ctx.scaled_dtc_batch(3, 3, ths), n=20)
print("scaled_dtc_batch N=8496 M=50, 64 candidates, fused small-problem path: %.3f ms per batch = %.4f ms per candidate (device launches %d)" % (ms, ms / 64, ctx.last_timing()[1]))
for lc in ():
    os.environ["GPAR_SS_LC"] = str(lc)
    ms = bench(lambda: ctx.scaled_dtc_batch(3, 3, ths), n=20)
    print("   whitening chunk length %4d: %.3f ms per batch" % (lc, ms))
os.environ.pop("GPAR_SS_LC", None)
Xe = rng.normal(size=(156, 5)); ctx.set_inputs(Xe); ctx.set_pseudo(Xe); ctx.set_times(np.arange(156) / 256.0); ctx.set_outputs(rng.normal(size=156))
ms = bench(lambda: ctx.scaled_dtc_batch(3, 3, ths), n=20)
print("scaled_dtc_batch EEG N=M=156 D=5, 64 candidates, fused small-problem path: %.3f ms per batch = %.4f ms per candidate" % (ms, ms / 64))
ctx.set_times(x); ctx.set_inputs(X); ctx.set_pseudo(Z); ctx.set_outputs(y_obs[1])
os.environ["GPAR_SCALED_SMALL"] = "0"
for lanes in (1, 4, 16):
    os.environ["GPAR_LANES"] = str(lanes)
    ms = bench(lambda: ctx.scaled_dtc_batch(3, 3, ths), n=10)
    print("scaled_dtc_batch N=8496 M=50, 64 candidates, lane path, %2d lanes: %.3f ms per batch = %.4f ms per candidate" % (lanes, ms, ms / 64))
os.environ.pop("GPAR_LANES"); os.environ.pop("GPAR_SCALED_SMALL")
