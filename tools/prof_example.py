"""The reference's live example (examples/gpar_scaled_example.py: 3 outputs, N = 8 496, 100 000 prediction points) timed warm:
plain Nelder-Mead, speculative Nelder-Mead (candidate points of an iteration in one batched call), 8 lock-step restarts."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(1, "examples")
import importlib.util
spec = importlib.util.spec_from_file_location("ex", "examples/gpar_scaled_example.py")
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
import io, contextlib
def run(**kw):
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter(); nrmse, _ = mod.main(iterations=150, quiet=True, **kw); dt = time.perf_counter() - t0
    return dt, nrmse
run(); run(speculative=True); run(n_restarts=8)          # warm-up: context, allocations, graph capture
for name, kw in (("plain", {}), ("speculative", {"speculative": True}), ("8 restarts", {"n_restarts": 8})):
    dts = [run(**kw) for _ in range(3)]
    print("%-12s fit + predict %.3f s (best of 3), normalised RMSE %s" % (name, min(d for d, _ in dts), ["%.3f" % v for v in dts[0][1]]))
