"""Strong scaling of ONE DTC logpdf+grad evaluation with the rows sharded over the devices of a group
(gpar_group_dtc_logpdf_sharded): N total fixed, each member holds N / ndev rows, one NCCL all-reduce per evaluation.

    python tools/bench_sharded_dtc.py --devices 2 --npoints 2000000 --pseudo 1024
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpar_at_scale_b200 as gp

ap = argparse.ArgumentParser()
ap.add_argument("--devices", type=int, default=1); ap.add_argument("--npoints", type=int, default=2_000_000)
ap.add_argument("--pseudo", type=int, default=1024); ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
rng = np.random.default_rng(1)
x = rng.uniform(0.0, 100.0, a.npoints); z = np.linspace(x.min(), x.max(), a.pseudo)
y = np.sin(x) + 0.3 * np.cos(3.1 * x) + 0.1 * rng.normal(size=a.npoints)
th = np.log([1.0, 1.0, 0.1])
g = gp.Group(list(range(a.devices)))
per = (a.npoints + a.devices - 1) // a.devices
for i, m in enumerate(g.members):
    sl = slice(i * per, min(a.npoints, (i + 1) * per))
    m.set_inputs(x[sl]); m.set_pseudo(z); m.set_outputs(y[sl])
for _ in range(3):
    v, gr = g.dtc_logpdf_sharded(gp.MATERN52, th, grad=True)
t0 = time.perf_counter()
for _ in range(a.steps):
    v, gr = g.dtc_logpdf_sharded(gp.MATERN52, th, grad=True)
dt = (time.perf_counter() - t0) / a.steps
print(json.dumps({"metric": "sharded DTC logpdf+grad ms per evaluation", "value": dt * 1e3, "unit": "ms", "n_gpus": a.devices, "scaling": "strong",
                  "config": {"workload": "one dtc_logpdf_grad, N=%d rows sharded over the devices, M=%d" % (a.npoints, a.pseudo)},
                  "logpdf": v, "grad": gr.tolist()}))
g.close()
