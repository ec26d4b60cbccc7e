# make_reference_golden.jl — REFERENCE-SIDE golden vectors for the B200 library's parity tests.
#
# The oracle under oracle/ is a restatement of the reference's algorithm and of its un-vendored dependencies
# (Stheno ~0.6, TemporalGPs ~0.1-0.2, Optim); nothing the reference itself computed is in the test loop because
# Julia is not installed where the library is built ("parity unpinned", DESIGN.md 2).  This script closes that gap:
# run it ONCE on any machine where the reference's environment instantiates,
#
#     julia --project=/path/to/GPAR-at-scale /path/to/this/repo/julia/make_reference_golden.jl
#
# and commit the file it writes, tests/golden/reference_outputs.json.  tests/test_reference_golden.py then compares
# the CPU oracle (pytest -m "not gpu") AND the CUDA library through its C ABI (pytest -m gpu) with the reference's
# own numbers at 1e-8 relative (skipped while the file is absent).
#
# Every case calls the reference's exported functions or, where the reference only has a closure, the same
# Stheno / TemporalGPs calls the closure makes (cited).  Inputs: tests/golden/reference_inputs.jl (seeded, written by
# tests/golden/make_reference_inputs.py; matrices hold one ROW per point).  No package beyond the reference's own
# dependencies is needed (the JSON is emitted by hand, floats via repr = exact round trip).
using GPARatScale
using Stheno
using Stheno: GPC, EQ, Matern12, Matern32, Matern52, ColVecs, pairwise, stretch, kernel
using TemporalGPs
using TemporalGPs: to_sde, smooth, SArrayStorage, decorrelate
using Optim
using Distributions
using LinearAlgebra

const REPO = normpath(joinpath(@__DIR__, ".."))
include(joinpath(REPO, "tests", "golden", "reference_inputs.jl"))        # const IN

# ---- tiny JSON emitter --------------------------------------------------------------------------------------
jval(x::Bool) = x ? "true" : "false"
jval(x::Integer) = string(x)
jval(x::Real) = isfinite(x) ? repr(Float64(x)) : "null"
jval(x::AbstractString) = "\"" * x * "\""
jval(x::AbstractVector) = "[" * join(map(jval, x), ",") * "]"
jval(x::Tuple) = jval(collect(x))
jval(x::AbstractMatrix) = "[" * join([jval(collect(x[i, :])) for i in 1:size(x, 1)], ",") * "]"      # list of ROWS
jval(d::AbstractDict) = "{" * join(["\"$(k)\":" * jval(d[k]) for k in sort(collect(keys(d)))], ",") * "}"

const OUT = Dict{String, Any}()
kern_of(code) = code == 0 ? EQ() : code == 1 ? Matern12() : code == 2 ? Matern32() : Matern52()
const KNAMES = Dict(0 => "eq", 1 => "matern12", 2 => "matern32", 3 => "matern52")
colvecs(X::AbstractMatrix, D = size(X, 2)) = to_ColVecs([collect(X[:, d]) for d in 1:D])      # util.jl:16-22

OUT["versions"] = Dict("julia" => string(VERSION),
                       "note" => "record `Pkg.status()` of Stheno / TemporalGPs / Optim next to this file when committing it")

# ---- A. parameter transforms (util.jl:36-55) ----------------------------------------------------------------
OUT["unpack_gp"] = collect(unpack_gp(IN["theta3"]))
OUT["unpack_gpar"] = collect(unpack_gpar(IN["theta5"]))

# ---- B. kernels, pairwise, masks (dtc.jl:31,104; gpar_scaled_inference.jl:89,156-157; util.jl:57-123) -------
let X = IN["kern_X"], Z = IN["kern_Z"], l = IN["kern_l"], s = IN["kern_var"]^2
    for code in 0:3, D in 1:3
        k = kernel(kern_of(code), l = l, s = s)                                   # dtc.jl:31
        OUT["pairwise_$(KNAMES[code])_D$(D)"] = pairwise(k, colvecs(X, D), colvecs(Z, D))
    end
    MX = colvecs(IN["mask_X"]); MY = colvecs(IN["mask_Y"])
    OUT["mask_time"] = pairwise(stretch(EQ(), get_time_mask(3)), MX, MY)          # util.jl:63-70
    OUT["mask_out"] = pairwise(stretch(EQ(), get_output_mask(3)), MX, MY)         # util.jl:72-79
end

# ---- C/D. exact GP and exact GPAR: log-pdf and posterior marginals (optimized.jl:28-36,132-154,94,236) ------
let x = IN["exact_x"], y1 = IN["exact_y1"], y2 = IN["exact_y2"], y3 = IN["exact_y3"], xs = IN["exact_xs"]
    l, pv, ns = unpack_gp(IN["theta3"])
    for code in (0, 3)
        k = pv^2 * stretch(kern_of(code), 1 / l)                                  # optimized.jl:30
        f = GP(k, GPC())
        OUT["exact_gp_logpdf_$(KNAMES[code])"] = logpdf(f(x, ns^2), y1)           # optimized.jl:34
        gp = GP(kernel(kern_of(code), l = l, s = pv^2), GPC())                    # optimized.jl:55-56
        post = gp | (gp(x, ns^2) ← y1)                                            # optimized.jl:94
        ms = marginals(post(xs))                                                  # plot_examples.jl:106-108
        OUT["exact_gp_post_mean_$(KNAMES[code])"] = mean.(ms)
        OUT["exact_gp_post_std_$(KNAMES[code])"] = std.(ms)
    end
    tl, tv, ol, ov, sg = unpack_gpar(IN["theta5"])
    for (name, cols, yy, cols_s) in (("D2", [x, y1], y2, [xs, IN["exact_xs_y1"]]),
                                     ("D3", [x, y1, y2], y3, [xs, IN["exact_xs_y1"], IN["exact_xs_y2"]]))
        inp = to_ColVecs(cols); L = length(cols)
        for (ct, co) in ((3, 3), (0, 0), (3, 0))
            kt = kernel(stretch(kern_of(ct), get_time_mask(L)), l = tl, s = tv^2)        # optimized.jl:133-136
            ko = kernel(stretch(kern_of(co), get_output_mask(L)), l = ol, s = ov^2)      # optimized.jl:138-140
            f = GP(kt + ko, GPC())
            tag = "$(name)_$(KNAMES[ct])_$(KNAMES[co])"
            OUT["exact_gpar_logpdf_$(tag)"] = logpdf(f(inp, sg^2), yy)                    # optimized.jl:152
            post = f | (f(inp, sg^2) ← yy)                                               # optimized.jl:236
            ms = marginals(post(to_ColVecs(cols_s)))                                      # eeg.jl:185-188
            OUT["exact_gpar_post_mean_$(tag)"] = [m.μ for m in ms]
            OUT["exact_gpar_post_std_$(tag)"] = [m.σ for m in ms]
        end
    end
end

# ---- E. state-space GP: logpdf, decorrelate, smooth (temporal_gp_inference.jl:15-39,78,109; dtc.jl:106) ------
let y = IN["lgssm_y"], rv = IN["lgssm_noise_vector"]
    l, pv, ns = unpack_gp(IN["theta3"])
    for code in 1:3, (gname, t) in (("irregular", IN["lgssm_t"]), ("regular", IN["lgssm_t_regular"]))
        tag = "$(KNAMES[code])_$(gname)"
        lg = create_lgssm(t, l, pv, ns, kern_of(code))                                   # temporal_gp_inference.jl:15-39
        OUT["lgssm_logpdf_$(tag)"] = logpdf(lg, y)                                        # :78
        lml, alpha = decorrelate(lg, y)                                                   # dtc.jl:106
        OUT["lgssm_decorrelate_lml_$(tag)"] = lml
        OUT["lgssm_decorrelate_alpha_$(tag)"] = collect(alpha)
        _, ysm, _ = smooth(lg, y)                                                         # :109
        OUT["lgssm_smooth_mean_$(tag)"] = [g.m[1] for g in ysm]                           # GPAR_scaled_examples.jl:111
        OUT["lgssm_smooth_var_$(tag)"] = [g.P[1] for g in ysm]                            # GPAR_scaled_examples.jl:129 (a variance)
        lgv = create_lgssm(t, l, pv, ns, kern_of(code), noise_vector = rv)                # :32-38 (the 1e10 trick)
        OUT["lgssm_noisevec_logpdf_$(tag)"] = logpdf(lgv, y)
        _, ysv, _ = smooth(lgv, y)
        OUT["lgssm_noisevec_smooth_mean_$(tag)"] = [g.m[1] for g in ysv]
        OUT["lgssm_noisevec_smooth_var_$(tag)"] = [g.P[1] for g in ysv]
    end
    # the regular grid as an AbstractRange (toy_data.jl:6; TemporalGPs RegularSpacing path)
    tr = range(0.0, step = 1 / 30, length = length(y))
    OUT["lgssm_logpdf_matern52_range"] = logpdf(create_lgssm(tr, l, pv, ns, Matern52()), y)
end

# ---- F. the scaled-GPAR objective (dtc.jl:29-48,83-128) and the reference's own self-check -------------------
let X = IN["scaled_X"], Z = IN["scaled_Z"], t = IN["scaled_t"], y = IN["scaled_y"]
    tl, tv, ol, ov, sg = unpack_gpar(IN["theta5"])
    for (ct, co) in ((3, 3), (3, 0), (2, 3), (1, 3)), D in 1:2
        gp_prior = GP(kernel(kern_of(co), l = ol, s = ov^2), GPC())                       # dtc.jl:31-32
        f = gp_prior(colvecs(X, D), sg^2); u = gp_prior(colvecs(Z, D), sg^2)              # dtc.jl:34-35
        tk = kernel(kern_of(ct), l = tl, s = tv^2)                                        # dtc.jl:37
        dtc, A = compute_gpar_dtc_objective(f, u, t, y; time_kernel = tk, temporal_noise_sigma = sg)
        tag = "$(KNAMES[ct])_$(KNAMES[co])_D$(D)"
        OUT["scaled_dtc_$(tag)"] = dtc
        OUT["scaled_A_fro_$(tag)"] = norm(A)
        if ct == 3 && co == 3
            OUT["scaled_A_$(tag)"] = A                                                     # M x N, the second return value
            # compute_q_u with kernels built as gpar_scaled_inference.jl:58-59 builds them
            q_u, U_u = GPARatScale.compute_q_u(colvecs(X, D), colvecs(Z, D), t, y;
                                               out_kernel = kernel(kern_of(co), l = ol, s = ov^2), time_kernel = tk,
                                               temporal_noise_sigma = sg)
            OUT["q_u_mean_$(tag)"] = collect(mean(q_u))
            OUT["q_u_cov_$(tag)"] = Matrix(cov(q_u))
            OUT["q_u_U_$(tag)"] = Matrix(U_u)
        end
    end
    # examples/dtc_example.jl:8-64, on the seeded small set: LGSSM-whitened DTC vs Stheno-style dense DTC
    x = IN["exact_x"]; y1 = IN["exact_y1"]; y2 = IN["exact_y2"]; z = IN["selfcheck_Z"]
    gp_prior = GP(Matern52(), GPC())
    f = gp_prior(y1, 0.05^2); u = gp_prior(z, 0.05^2)
    dtc, A = compute_gpar_dtc_objective(f, u, x, y2; time_kernel = Matern52(), temporal_noise_sigma = 0.04)
    noise_matrix = cov(GP(Matern52(), GPC())(x, 0.04^2))
    chol = cholesky(noise_matrix)
    A2 = cholesky(Symmetric(cov(u))).U' \ (chol.U' \ cov(f, u))'
    Le = cholesky(Symmetric(A2 * A2' + I)); d = chol.U' \ (y2 - mean(f))
    tmp = logdet(chol) + logdet(Le) + sum(abs2, d) - sum(abs2, Le.U' \ (A2 * d))
    OUT["selfcheck_dtc_lgssm"] = dtc
    OUT["selfcheck_dtc_dense"] = -(length(y2) * log(2π) + tmp) / 2
    OUT["selfcheck_A_maxabsdiff"] = maximum(abs.(A - A2))
end

# ---- G. prediction protocol, deterministic parts (temporal_gp_inference.jl:55-66,93-112) ---------------------
let t = IN["scaled_t"], y = IN["scaled_y"], ts = IN["pred_ts"]
    l, pv, ns = unpack_gp(IN["theta3"])
    latent = vcat(t, ts); outputs = vcat(y, repeat([0], length(ts)))
    perm = sortperm(latent); rev = sortperm(perm)
    nv = vcat(repeat([ns^2], length(t)), repeat([1e10], length(ts)))[perm]
    lg = create_lgssm(latent[perm], l, pv, ns, Matern52(), noise_vector = nv)
    _, ysm, _ = smooth(lg, outputs[perm])
    obs = ysm[rev][length(y)+1:end]
    OUT["sde_pred_mean"] = [g.m[1] for g in obs]
    OUT["sde_pred_var"] = [g.P[1] for g in obs]
end

# ---- H. Optim.NelderMead with its default options (dtc.jl:58-61; optimized.jl:45,164) ------------------------
rosen2(x) = (1.0 - x[1])^2 + 100.0 * (x[2] - x[1]^2)^2
rosen5(x) = sum(100.0 * (x[i+1] - x[i]^2)^2 + (1.0 - x[i])^2 for i in 1:4)
for (name, f, x0, iters) in (("nm_rosen2", rosen2, IN["nm_x0_2d"], 1000), ("nm_rosen5", rosen5, IN["nm_x0_5d"], 1000),
                             ("nm_rosen5_60it", rosen5, IN["nm_x0_5d"], 60))
    r = Optim.optimize(f, copy(x0), NelderMead(), Optim.Options(iterations = iters, store_trace = true))
    OUT[name] = Dict("minimizer" => Optim.minimizer(r), "minimum" => Optim.minimum(r), "iterations" => Optim.iterations(r),
                     "f_calls" => Optim.f_calls(r), "converged" => Optim.converged(r),
                     "trace_f" => [tr.value for tr in Optim.trace(r)])
end

open(joinpath(REPO, "tests", "golden", "reference_outputs.json"), "w") do io
    write(io, jval(OUT)); write(io, "\n")
end
println("wrote tests/golden/reference_outputs.json with $(length(OUT)) entries")
