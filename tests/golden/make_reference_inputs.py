"""Seeded inputs for the REFERENCE-side golden generator (julia/make_reference_golden.jl).

The reference cannot run in this image (no Julia; Stheno / TemporalGPs un-vendored), so the oracle is
"parity unpinned" (DESIGN 2).  This script writes the inputs of a fixed set of small cases twice — as
`reference_inputs.json` (read by tests/test_reference_golden.py) and as `reference_inputs.jl` (a Julia literal,
`include`d by the generator so that it needs no JSON package; floats are printed with repr = exact round trip).
Running `julia --project=<GPAR-at-scale checkout> julia/make_reference_golden.jl` where Stheno 0.6 / TemporalGPs
0.1-0.2 / Optim are installed produces tests/golden/reference_outputs.json; the tests then compare the oracle
(CPU) and the CUDA library (GPU) against the reference's OWN numbers.

    python tests/golden/make_reference_inputs.py
"""
import json
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    rng = np.random.default_rng(20261018)
    c = {}
    c["theta3"] = [0.3, -0.2, -1.7]
    c["theta5"] = [0.4, -0.1, 0.2, -0.3, -1.2]
    # kernels / pairwise (util.jl:57-96 docstring numbers included)
    c["kern_X"] = rng.normal(size=(12, 3)).tolist()
    c["kern_Z"] = rng.normal(size=(5, 3)).tolist()
    c["kern_l"] = 1.3
    c["kern_var"] = 0.8
    c["mask_X"] = [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]]        # rows = points (the docstring's ColVecs columns)
    c["mask_Y"] = [[1.5, 2.5, 3.5], [4.5, 5.5, 6.5], [7.5, 8.5, 9.5]]
    # exact GP / GPAR (optimized.jl:28-36,132-154,94,236)
    n = 30
    x = np.linspace(0.0, 1.0, n)
    y1 = -np.sin(10 * np.pi * (x + 1)) / (2 * x + 1) - x ** 4 + 0.05 * rng.normal(size=n)
    y2 = np.cos(y1) ** 2 + np.sin(3 * x) + 0.05 * rng.normal(size=n)
    y3 = y2 * y1 ** 2 + 3 * x + 0.05 * rng.normal(size=n)
    c["exact_x"] = x.tolist(); c["exact_y1"] = y1.tolist(); c["exact_y2"] = y2.tolist(); c["exact_y3"] = y3.tolist()
    xs = np.linspace(-0.1, 1.1, 11)
    c["exact_xs"] = xs.tolist()
    c["exact_xs_y1"] = np.sin(xs).tolist(); c["exact_xs_y2"] = np.cos(xs).tolist()
    # state-space (temporal_gp_inference.jl:15-39,78,109): irregular grid with exact duplicates, noise-vector variant
    nt = 200
    t = np.cumsum(rng.exponential(1 / 30, nt)); t[50] = t[49]; t[120] = t[119]; t[121] = t[119]
    c["lgssm_t"] = t.tolist()
    c["lgssm_y"] = (np.sin(3 * t) + 0.2 * rng.normal(size=nt)).tolist()
    rv = np.full(nt, (np.exp(c["theta3"][2]) + 1e-3) ** 2); rv[rng.choice(nt, 20, replace=False)] = 1e10
    c["lgssm_noise_vector"] = rv.tolist()
    c["lgssm_t_regular"] = (np.arange(nt) / 30.0).tolist()
    # scaled GPAR objective and q(u) (dtc.jl:83-128; gpar_scaled_inference.jl:141-196)
    ns, m = 150, 12
    ts_ = np.cumsum(rng.exponential(1 / 30, ns))
    X = rng.normal(size=(ns, 2))
    c["scaled_t"] = ts_.tolist(); c["scaled_X"] = X.tolist(); c["scaled_Z"] = X[:: ns // m][:m].tolist()
    c["scaled_y"] = (np.sin(ts_) + 0.5 * X[:, 0] + 0.1 * rng.normal(size=ns)).tolist()
    # plain DTC self-check sizes (dtc_example.jl:26-50): N = 30, M = 10, sigma_obs = 0.05, sigma_t = 0.04
    c["selfcheck_Z"] = np.linspace(y1.min(), y1.max(), 10).tolist()
    # prediction protocol (gpar_scaled_inference.jl:74-135 deterministic parts; temporal_gp_inference.jl:55-66,93-112)
    c["pred_ts"] = np.sort(rng.uniform(ts_[0] - 0.2, ts_[-1] + 0.3, 40)).tolist()
    c["pred_Xs"] = rng.normal(size=(40, 2)).tolist()
    # Optim.NelderMead defaults
    c["nm_x0_2d"] = [-1.2, 1.0]
    c["nm_x0_5d"] = [0.3, -0.4, 0.5, 0.1, -0.2]
    return c


def to_julia(v, indent=0):
    if isinstance(v, float) or isinstance(v, int):
        return repr(float(v))
    if isinstance(v, list) and v and isinstance(v[0], list):           # matrix: rows = points
        return "[" + "; ".join(" ".join(repr(float(e)) for e in row) for row in v) + "]"
    if isinstance(v, list):
        return "[" + ", ".join(repr(float(e)) for e in v) + "]"
    raise TypeError(type(v))


def main():
    c = cases()
    with open(os.path.join(HERE, "reference_inputs.json"), "w") as f:
        json.dump(c, f, indent=0, sort_keys=True)
    with open(os.path.join(HERE, "reference_inputs.jl"), "w") as f:
        f.write("# written by tests/golden/make_reference_inputs.py — do not edit; matrices: one ROW per point\n")
        f.write("const IN = Dict{String, Any}(\n")
        for k in sorted(c):
            f.write('    "%s" => %s,\n' % (k, to_julia(c[k])))
        f.write(")\n")


if __name__ == "__main__":
    main()
