"""ONE output whose rows are sharded over several GPUs (NEW; the reference has one CPU process): the scaled-GPAR fit of
src/gp/dtc.jl:11-77 and the q(u) of src/gp/gpar_scaled_inference.jl:141-196 run on row slices — every device holds the full (t, y)
and its slice of the inputs, two small NCCL collectives per evaluation — and the Monte-Carlo prediction (:74-135), which needs no
N x M array, runs on one device.

    python examples/sharded_output_example.py --devices 2            # two GPUs of one box
    GPAR_GROUP_LOOPBACK=1 python examples/sharded_output_example.py --devices 3 --same-device     # 1-GPU box: three members on device 0
"""
import argparse
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, os.path.dirname(os.path.abspath(__file__)))
import gpar_at_scale_b200 as gp
from gpar_at_scale_b200 import api
import toy_data as data


def main(devices=2, same_device=False, iterations=60, seed=0, true_samples=20000, optimizer="neldermead", quiet=False):
    rng = np.random.default_rng(seed)
    x, y_obs, x_true, y_true = data.generate_big_dataset(rng, true_samples=true_samples)
    y1, y2, _ = y_obs
    test_y1, test_y2, _ = y_true
    pseudo_y1 = np.linspace(test_y1.min(), test_y1.max(), 50)
    g = gp.Group([0] * devices if same_device else list(range(devices)))
    try:
        t0 = time.perf_counter()
        y2_out, y2_std = api.get_gpar_scaled_predictions([y1], [pseudo_y1], x, y2, x_true, [test_y1], iterations=iterations, rng=rng,
                                                         debug=not quiet, group=g)
        dt = time.perf_counter() - t0
    finally:
        g.close()
    inside = x_true <= x.max()
    rmse = float(np.sqrt(np.mean((y2_out[inside] - test_y2[inside]) ** 2)))
    if not quiet:
        print("fit + q(u) on %d row slices, prediction on one device: %.2f s; RMSE of y2 vs the true function on the training span %.3f" % (devices, dt, rmse))
    return rmse, dt


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=2); ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--iterations", type=int, default=60)
    a = ap.parse_args()
    main(a.devices, a.same_device, a.iterations)
