"""Synthetic 3-output GPAR data — host-side mirror of src/data/toy_data.jl (the only file under the
reference's src/data).  NOT part of the product package: SURVEY 2.1 #7 keeps the data generators in Julia; this
copy only feeds the example driver and the tests.  O(N) element-wise work: stays on the host, as in the reference; it is the
shape source for the benchmark inputs.  The reference draws from Julia's unseeded global RNG
(toy_data.jl:34-36); here a numpy Generator can be passed for reproducibility."""
import numpy as np

START = 0.0
STEP_SIZE = 1.0 / 30.0
NOISE_MU = 0.0


def nuke(x, nr_nuked_intervals, nuked_per_interval):
    """toy_data.jl:42-57: removes `nuked_per_interval` points after each of the interval boundaries."""
    if nr_nuked_intervals == 0:
        return x, 0
    kept = len(x) // (nr_nuked_intervals + 1)
    acc = [x[:kept]]
    for i in range(1, nr_nuked_intervals + 1):
        acc.append(x[i * kept + nuked_per_interval:(i + 1) * kept])
    out = np.concatenate(acc)
    return out, len(x) - len(out)


def _generate_toy_data(data_samples, true_samples, f1, f2, f3, observation_noise=0.05, extended_true_period=0.0,
                       nr_nuked_intervals=0, nuked_per_interval=0, rng=None):
    """toy_data.jl:9-40.  NB the noise std is observation_noise^2 (Distributions' Normal takes a std;
    the reference passes the square — quirk kept, toy_data.jl:29)."""
    rng = rng or np.random.default_rng()
    stop = STEP_SIZE * data_samples
    x_true = np.linspace(START, stop + extended_true_period, true_samples)
    y1_true = f1(x_true); y2_true = f2(x_true, y1_true); y3_true = f3(x_true, y1_true, y2_true)
    x = np.linspace(START, stop, data_samples)
    x, removed = nuke(x, nr_nuked_intervals, nuked_per_interval)
    n = data_samples - removed
    sd = observation_noise ** 2
    y1 = f1(x) + rng.normal(NOISE_MU, sd, n)
    y2 = f2(x, y1) + rng.normal(NOISE_MU, sd, n)
    y3 = f3(x, y1, y2) + rng.normal(NOISE_MU, sd, n)
    return x, [y1, y2, y3], x_true, [y1_true, y2_true, y3_true]


def f1_small(x): return -np.sin(10 * np.pi * (x + 1)) / (2 * x + 1) - x ** 4
def f2_small(x, y1): return np.cos(y1) ** 2 + np.sin(3 * x)
def f3_small(x, y1, y2): return y2 * y1 ** 2 + 3 * x


def generate_small_dataset(rng=None):
    """toy_data.jl:59-74: N = 30, 1000 true points."""
    return _generate_toy_data(30, 1000, f1_small, f2_small, f3_small, observation_noise=0.05, rng=rng)


def f1_big(x): return 3 + -np.sin(np.pi / 10 * (x + 1)) - x ** 0.3
def f2_big(x, y1): return np.cos(y1) ** 2 + np.sin(np.pi / 20 * x)
def f3_big(x, y1, y2): return y2 * y1 ** 2 + 0.1 * x


def generate_big_dataset(rng=None, data_samples=10000, true_samples=100000):
    """toy_data.jl:76-98: N = 10 000 (8 496 after `nuke`), 100 000 true points on [0, 333.3 + 50]."""
    return _generate_toy_data(data_samples, true_samples, f1_big, f2_big, f3_big, observation_noise=0.8,
                              extended_true_period=50.0, nr_nuked_intervals=5, nuked_per_interval=300, rng=rng)
