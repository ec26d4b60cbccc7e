"""The reference's live example, `big_synthetic_dataset()` of examples/GPAR_scaled_examples.jl:86-216,
without the plots: 3-output synthetic GPAR, N = 8 496 training points (10 000 minus the nuked
intervals), 100 000 prediction points; y1 by the time-only state-space GP, y2 (M = 50 pseudo-points
on a line) and y3 (9 x 9 grid) by scaled GPAR.  The reference runs the two fits under 170 s / 250 s
Nelder-Mead time limits on the CPU; here an iteration budget is used.

    python examples/gpar_scaled_example.py [--iterations 150]
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


def main(iterations=150, seed=0, true_samples=100000, quiet=False, speculative=False, n_restarts=1):
    rng = np.random.default_rng(seed)
    x, y_obs, x_true, y_true = data.generate_big_dataset(rng, true_samples=true_samples)
    y1, y2, y3 = y_obs
    test_y1, test_y2, test_y3 = y_true
    t0 = time.perf_counter()
    _, (y1_out, y1_var) = api.get_sde_predictions(x, y1, x_true, kernel_structure=api.Matern52(), i_log_time_l=-3, i_log_time_var=0.2,
                                                  i_log_noise_sigma=-10, debug=not quiet, return_arrays=True, speculative=speculative)          # :102-111
    pseudo_y1 = np.linspace(test_y1.min(), test_y1.max(), 50)                                                         # :115
    y2_out, y2_std = api.get_gpar_scaled_predictions([y1], [pseudo_y1], x, y2, x_true, [test_y1], iterations=iterations,
                                                     debug=not quiet, rng=rng, speculative=speculative, n_restarts=n_restarts)                                        # :132-142
    d1 = np.linspace(test_y1.min(), test_y1.max(), 9); d2 = np.linspace(test_y2.min(), test_y2.max(), 9)
    pseudo_y3 = np.array([[a, b] for b in d2 for a in d1])                                                            # Iterators.product order, :145-148
    y3_out, y3_std = api.get_gpar_scaled_predictions([y1, y2], pseudo_y3, x, y3, x_true, [test_y1, y2_out], iterations=iterations,
                                                     debug=not quiet, rng=rng, speculative=speculative, n_restarts=n_restarts)                                        # :165-175
    dt = time.perf_counter() - t0
    inside = x_true <= x.max()
    pairs = ((y1_out, test_y1), (y2_out, test_y2), (y3_out, test_y3))
    rmse = [float(np.sqrt(np.mean((o[inside] - tr[inside]) ** 2))) for o, tr in pairs]
    nrmse = [r / float(np.std(tr[inside])) for r, (o, tr) in zip(rmse, pairs)]      # relative to the spread of the true function
    if not quiet:
        print("fit + predict: %.1f s; RMSE vs true functions on the training span: y1 %.3f  y2 %.3f  y3 %.3f (normalised %.3f %.3f %.3f)"
              % (dt, *rmse, *nrmse))
    return nrmse, dt


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--iterations", type=int, default=150)
    ap.add_argument("--speculative", action="store_true", help="candidate points of a Nelder-Mead iteration in one batched call")
    ap.add_argument("--restarts", type=int, default=1, help="lock-step Nelder-Mead restarts per output")
    a = ap.parse_args()
    main(a.iterations, speculative=a.speculative, n_restarts=a.restarts)
