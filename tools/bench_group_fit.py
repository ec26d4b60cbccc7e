"""GPAR fit at scale through the single-process C-ABI group (gpar_group_fit): the same problem, task list, start
points and Nelder-Mead budget as tools/bench_gpar_fit.py (one process per GPU over torch.distributed), but ONE
process drives all visible devices — what the Julia package would do with one `ccall`.

    python tools/bench_group_fit.py --devices 8 --npoints 2097152 --pseudo 2048 --outputs 8 --restarts 2 --iterations 8
Prints one JSON line: {"metric": "GPAR fit s", ...}.
"""
import argparse
import json
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_gpar_fit import synth          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=1)
    ap.add_argument("--npoints", type=int, default=2097152); ap.add_argument("--pseudo", type=int, default=2048)
    ap.add_argument("--outputs", type=int, default=8); ap.add_argument("--restarts", type=int, default=2)
    ap.add_argument("--iterations", type=int, default=8); ap.add_argument("--optimizer", default="neldermead", choices=["neldermead", "lbfgs"])
    a = ap.parse_args()
    import gpar_at_scale_b200 as gp
    from gpar_at_scale_b200 import chain
    t, Y = synth(a.outputs, a.npoints)
    tasks = []
    for o in range(a.outputs):
        X = np.ascontiguousarray(Y[:o].T) if o else None
        Z = chain.strided_pseudo_inputs(X, a.pseudo) if o else None
        for r in range(a.restarts):
            th0 = np.random.default_rng([4, o, r]).random(3 if o == 0 else 5)      # the start points of chain.fit_chain(seed=4)
            tasks.append({"X": X, "Z": Z, "y": Y[o], "theta0": th0, "output": o})
    g = gp.Group(list(range(a.devices)))
    # warm-up: one objective evaluation per member allocates the panels
    for m in g.members:
        m.set_inputs(tasks[-1]["X"][:, :1]); m.set_pseudo(chain.strided_pseudo_inputs(tasks[-1]["X"][:, :1], a.pseudo)); m.set_times(t); m.set_outputs(Y[1])
    g.scaled_dtc(gp.MATERN52, gp.MATERN52, np.zeros((a.devices, 5)))
    t0 = time.perf_counter()
    minimum, minimizer, calls, member = g.fit(t, tasks, gp.MATERN52, gp.MATERN52, a.iterations, optimizer=a.optimizer)
    dt = time.perf_counter() - t0
    best = {}
    for k, tk in enumerate(tasks):
        o = tk["output"]
        if o not in best or minimum[k] < best[o]:
            best[o] = float(minimum[k])
    per_member = [int(calls[member == i].sum()) for i in range(a.devices)]
    print(json.dumps({"metric": "GPAR fit s", "value": dt, "unit": "s", "n_gpus": a.devices, "higher_is_better": False, "scaling": "strong",
                      "config": {"workload": "gpar_group_fit (one process) outputs=%d N=%d M=%d restarts=%d %s_iterations=%d"
                                 % (a.outputs, a.npoints, a.pseudo, a.restarts, a.optimizer, a.iterations)},
                      "objective_evals_per_member": per_member, "tasks": len(tasks), "best_nlml": {str(o): best[o] for o in sorted(best)}}))
    g.close()


if __name__ == "__main__":
    main()
