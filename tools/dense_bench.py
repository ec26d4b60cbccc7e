"""Device time of the hand-written dense routines (dense_la.cu) at the orders the path uses."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpar_at_scale_b200 as gp
ctx = gp.Context(0)
for n in (50, 81, 156, 512, 1024, 2048):
    r = ctx.dense_bench(n)
    fl = 2.0 * n ** 3
    print(n, {k: round(v, 4) for k, v in r.items()}, "full gemm TFLOP/s %.2f" % (fl / r["gemm_full"] / 1e9))
