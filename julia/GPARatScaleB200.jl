# GPARatScaleB200.jl — Julia-side binding of libgpar_b200.so (include/gpar_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The same symbols, argument
# order and memory layouts are exercised through Python ctypes (gpar-at-scale_b200/_ffi.py) by the
# test-suite; this file is what a maintainer of GPAR-at-scale adds so that the package's exported
# functions keep their names and signatures while their numeric cores run on the B200.
#
# Usage inside the reference package (src/GPARatScale.jl):
#     include("GPARatScaleB200.jl"); using .GPARatScaleB200
# and replace the bodies shown in INTEGRATION.md.
module GPARatScaleB200

using LinearAlgebra: PosDefException

const LIB = get(ENV, "GPAR_B200_LIB", "libgpar_b200.so")
const GPAR_OK, GPAR_ERR_INVALID, GPAR_ERR_CUDA, GPAR_ERR_NOT_POSDEF, GPAR_ERR_NOMEM = 0, 1, 2, 3, 4
@enum KernelCode EQ_K = 0 MATERN12_K = 1 MATERN32_K = 2 MATERN52_K = 3

mutable struct Ctx
    h::Ptr{Cvoid}
    function Ctx(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        st = ccall((:gpar_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r)
        st == GPAR_OK || error("gpar_ctx_create failed with status $st (no CUDA device?)")
        c = new(r[])
        finalizer(x -> ccall((:gpar_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), c)
        return c
    end
    Ctx(h::Ptr{Cvoid}, ::Val{:borrowed}) = new(h)      # member of a Group: the group owns the handle, no finalizer
end

function check(c::Ctx, st::Integer)
    st == GPAR_OK && return nothing
    msg = unsafe_string(ccall((:gpar_last_error, LIB), Cstring, (Ptr{Cvoid},), c.h))
    # what `cholesky` throws in the reference (src/gp/dtc.jl:119-120) and Optim propagates
    st == GPAR_ERR_NOT_POSDEF && throw(PosDefException(1))
    error("libgpar_b200 status $st: $msg")
end

# Stheno kernel structure -> code.  (Stheno.EQ, Matern12, Matern32, Matern52)
kernel_code(k) = let n = string(nameof(typeof(k)))
    n == "EQ" ? 0 : n == "Matern12" ? 1 : n == "Matern32" ? 2 : n == "Matern52" ? 3 :
        throw(ArgumentError("kernel $n has no libgpar_b200 code"))
end

# ---- resident data.  ColVecs.X is a D x N column-major Matrix{Float64} (src/util.jl:16-31):
#      exactly the N-records-of-D-doubles layout the ABI expects, so the pointer is passed as is.
set_inputs!(c::Ctx, X::Matrix{Float64}) =
    check(c, ccall((:gpar_set_inputs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64), c.h, X, size(X, 1), size(X, 2)))
set_pseudo!(c::Ctx, Z::Matrix{Float64}) =
    check(c, ccall((:gpar_set_pseudo, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64), c.h, Z, size(Z, 1), size(Z, 2)))
set_times!(c::Ctx, t::Vector{Float64}) =
    check(c, ccall((:gpar_set_times, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, t, length(t)))
# a range keeps TemporalGPs' RegularSpacing behaviour: constant transition matrix (toy_data.jl:6)
set_times!(c::Ctx, t::AbstractRange) =
    check(c, ccall((:gpar_set_times_range, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Int64), c.h, first(t), step(t), length(t)))
set_outputs!(c::Ctx, y::VecOrMat{Float64}) =      # N or N x batch (column-major: sequence b contiguous)
    check(c, ccall((:gpar_set_outputs, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int32), c.h, y, size(y, 1), size(y, 2)))
set_noise_vector!(c::Ctx, r::Union{Nothing, Vector{Float64}}) = r === nothing ?
    check(c, ccall((:gpar_set_noise_vector, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, C_NULL, 0)) :
    check(c, ccall((:gpar_set_noise_vector, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), c.h, r, length(r)))

# ---- compute ------------------------------------------------------------------------------
"DTC / VFE log-pdf (and gradient) for Sigma_y = sigma^2 I; theta = raw (log l, log var, log sigma)."
function dtc_logpdf(c::Ctx, k, theta::Vector{Float64}; vfe::Bool = false, jitter::Float64 = -1.0, grad::Bool = false)
    val = Ref{Float64}(0.0)
    g = grad ? zeros(3) : Float64[]
    check(c, ccall((:gpar_dtc_logpdf, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Float64, Ref{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k), theta, vfe, jitter, val, grad ? pointer(g) : C_NULL))
    return grad ? (val[], g) : val[]
end

"DTC / VFE objective with d/dtheta (3) and d/dZ (D x M, the ColVecs layout of Z) — pseudo-input optimisation."
function dtc_logpdf_zgrad(c::Ctx, k, theta::Vector{Float64}, D::Integer, M::Integer; vfe::Bool = false, jitter::Float64 = -1.0)
    val = Ref{Float64}(0.0); g = zeros(3); gz = zeros(D, M)
    check(c, ccall((:gpar_dtc_logpdf_zgrad, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Float64, Ref{Float64}, Ptr{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k), theta, vfe ? 1 : 0, jitter, val, g, gz))
    return val[], g, gz
end

"compute_gpar_dtc_objective (src/gp/dtc.jl:83-128): returns (dtc, A) like the reference."
function scaled_dtc(c::Ctx, k_time, k_out, theta::Vector{Float64}, N::Integer, M::Integer; return_A::Bool = false)
    val = Ref{Float64}(0.0)
    A = return_A ? zeros(M, N) : zeros(0, 0)
    check(c, ccall((:gpar_scaled_dtc, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k_time), kernel_code(k_out), theta, val, return_A ? pointer(A) : C_NULL))
    return val[], A
end

"compute_gpar_dtc_objective for the columns of `thetas` (5 x ncand) at once, on the resident data (simplex vertices x restarts
of the loop dtc.jl:58-61): (values, codes); a candidate whose Cholesky fails has value NaN and a non-zero code."
function scaled_dtc_batch(c::Ctx, k_time, k_out, thetas::Matrix{Float64})
    n = size(thetas, 2); vals = zeros(n); codes = zeros(Int32, n)
    check(c, ccall((:gpar_scaled_dtc_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Int32}),
                   c.h, kernel_code(k_time), kernel_code(k_out), thetas, n, vals, codes))
    return vals, codes
end

"compute_gpar_dtc_objective with d/d theta (5) — for Optim.LBFGS in place of NelderMead (dtc.jl:58-61)."
function scaled_dtc_grad(c::Ctx, k_time, k_out, theta::Vector{Float64})
    val = Ref{Float64}(0.0); grad = zeros(5)
    check(c, ccall((:gpar_scaled_dtc_grad, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ref{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k_time), kernel_code(k_out), theta, val, grad))
    return val[], grad
end

"compute_q_u (src/gp/gpar_scaled_inference.jl:141-196): (m_e, inv(D), U_u); params are positive values."
function compute_q_u(c::Ctx, k_time, k_out, params::Vector{Float64}, M::Integer)
    m_e = zeros(M); Dinv = zeros(M, M); U_u = zeros(M, M)
    check(c, ccall((:gpar_compute_q_u, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k_time), kernel_code(k_out), params, m_e, Dinv, U_u))
    return m_e, Dinv, U_u
end

"vcat / sortperm / gather / 1e10 noise vector on the device (temporal_gp_inference.jl:55-66,93-97; gpar_scaled_inference.jl:75-87,100-103)."
function set_merged!(c::Ctx, t::Vector{Float64}, y::Vector{Float64}, ts::Vector{Float64}, sigma2::Float64;
                     X::Matrix{Float64} = zeros(0, 0), Xs::Matrix{Float64} = zeros(0, 0))
    D = size(X, 1)
    check(c, ccall((:gpar_set_merged, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64),
                   c.h, t, y, D > 0 ? pointer(X) : C_NULL, length(t), ts, D > 0 ? pointer(Xs) : C_NULL, length(ts), Int32(D), sigma2))
end
"result[reverse_perm][N+1:end] of the last smoother / prediction (temporal_gp_inference.jl:111-112; gpar_scaled_inference.jl:132-133)."
function take_test(c::Ctx, Ns::Integer)
    a = zeros(Ns); b = zeros(Ns)
    check(c, ccall((:gpar_take_test, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), c.h, a, b))
    return a, b
end

"Seeded device draws from q_u (replaces rand(q_u) and U_u \\ eps, gpar_scaled_inference.jl:94-96); W stays resident."
function sample_q_u(c::Ctx, k_time, k_out, params::Vector{Float64}, seed::Integer, S::Integer, M::Integer; return_host::Bool = false)
    W = return_host ? zeros(M, S) : zeros(0, 0); E = return_host ? zeros(M, S) : zeros(0, 0)
    check(c, ccall((:gpar_sample_q_u, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, UInt64, Int32, Ptr{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k_time), kernel_code(k_out), params, UInt64(seed), Int32(S),
                   return_host ? pointer(W) : C_NULL, return_host ? pointer(E) : C_NULL))
    return W, E
end

"logpdf(lgssm, y) for every resident sequence; theta is 3 x batch_theta (column per model)."
# (with ONE resident sequence and several columns in `theta`, the columns are hyper-parameter candidates: pass batch = size(theta, 2))
function lgssm_logpdf(c::Ctx, k, theta::VecOrMat{Float64}, batch::Integer)
    lml = zeros(batch)
    check(c, ccall((:gpar_lgssm_logpdf, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int32, Ptr{Float64}),
                   c.h, kernel_code(k), theta, size(theta, 2), lml))
    return lml
end

"logpdf and d logpdf / d theta (3 x batch) — for gradient-based optimisers (Optim.LBFGS) in place of NelderMead."
function lgssm_logpdf_grad(c::Ctx, k, theta::VecOrMat{Float64}, batch::Integer)
    lml = zeros(batch); grad = zeros(3, batch)
    check(c, ccall((:gpar_lgssm_logpdf_grad, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k), theta, size(theta, 2), lml, grad))
    return lml, grad
end

function lgssm_decorrelate(c::Ctx, k, theta::Vector{Float64}, N::Integer, batch::Integer = 1)
    alpha = zeros(N, batch); lml = zeros(batch)
    check(c, ccall((:gpar_lgssm_decorrelate, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k), theta, alpha, lml))
    return lml, alpha
end

function lgssm_smooth(c::Ctx, k, theta::Vector{Float64}, N::Integer, batch::Integer = 1)
    m = zeros(N, batch); v = zeros(N, batch); lml = zeros(batch)
    check(c, ccall((:gpar_lgssm_smooth, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k), theta, m, v, lml))
    return lml, m, v
end

function exact_logpdf(c::Ctx, k_time, k_out, theta::Vector{Float64}, batch::Integer = 1)
    lml = zeros(batch)
    check(c, ccall((:gpar_exact_logpdf, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int32, Ptr{Float64}),
                   c.h, kernel_code(k_time), kernel_code(k_out), theta, length(theta), lml))
    return lml
end

"`ncand` hyper-parameter candidates (columns of `thetas`, 3 or 5 rows) of the exact log-pdf in one launch — the simplex
vertices x restarts of optimized.jl:45,164: (lml (batch x ncand), codes); a failed Cholesky gives NaN and a non-zero code."
function exact_logpdf_batch(c::Ctx, k_time, k_out, thetas::Matrix{Float64}, batch::Integer = 1)
    n = size(thetas, 2); lml = zeros(batch, n); codes = zeros(Int32, n)
    check(c, ccall((:gpar_exact_logpdf_batch, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Int32}),
                   c.h, kernel_code(k_time), kernel_code(k_out), thetas, size(thetas, 1), n, lml, codes))
    return lml, codes
end

function exact_posterior(c::Ctx, k_time, k_out, theta::Vector{Float64}, Xs::Matrix{Float64}, batch::Integer = 1)
    Ns = size(Xs, 2); mean = zeros(Ns, batch); var = zeros(Ns)
    check(c, ccall((:gpar_exact_posterior, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int32, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                   c.h, kernel_code(k_time), kernel_code(k_out), theta, length(theta), Xs, Ns, mean, var))
    return mean, var
end

# ---- several devices from this one Julia process (gpar_group_*): one member context per device -------------
struct FitTask                       # struct gpar_fit_task (include/gpar_b200.h); isbits, C layout
    X::Ptr{Float64}; D::Int32
    Z::Ptr{Float64}; M::Int64
    y::Ptr{Float64}
    theta0::NTuple{5, Float64}
end

mutable struct Group
    h::Ptr{Cvoid}
    members::Vector{Ctx}
    function Group(devices::Vector{<:Integer})
        r = Ref{Ptr{Cvoid}}(C_NULL); devs = Int32.(devices)
        st = ccall((:gpar_group_create, LIB), Cint, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), devs, length(devs), r)
        st == GPAR_OK || error("gpar_group_create failed with status $st (devices / NCCL unavailable?)")
        ms = [Ctx(ccall((:gpar_group_ctx, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Int32), r[], i - 1), Val(:borrowed)) for i in 1:length(devs)]
        g = new(r[], ms)
        finalizer(x -> ccall((:gpar_group_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), g)
        return g
    end
end

function gcheck(g::Group, st::Integer)
    st == GPAR_OK && return nothing
    error("libgpar_b200 group status $st: " * unsafe_string(ccall((:gpar_group_last_error, LIB), Cstring, (Ptr{Cvoid},), g.h)))
end

# member i evaluates thetas[:, i] on its resident data, all members at once -> (vals, grads, codes)
function group_scaled_dtc(g::Group, k_time, k_out, thetas::Matrix{Float64}; grad::Bool = false)
    n = length(g.members); vals = zeros(n); codes = zeros(Int32, n); grads = grad ? zeros(5, n) : nothing
    gcheck(g, ccall((:gpar_group_scaled_dtc, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                    g.h, kernel_code(k_time), kernel_code(k_out), thetas, vals, grad ? grads : C_NULL, codes))
    return vals, grads, codes
end
function group_dtc_logpdf(g::Group, k, thetas::Matrix{Float64}; vfe::Bool = false, jitter::Float64 = -1.0, grad::Bool = false)
    n = length(g.members); vals = zeros(n); codes = zeros(Int32, n); grads = grad ? zeros(3, n) : nothing
    gcheck(g, ccall((:gpar_group_dtc_logpdf, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                    g.h, kernel_code(k), thetas, vfe, jitter, vals, grad ? grads : C_NULL, codes))
    return vals, grads, codes
end

# ONE DTC / VFE objective whose rows are sharded over the members (same Z everywhere): all-reduce of the statistics
function group_dtc_logpdf_sharded(g::Group, k, theta::Vector{Float64}; vfe::Bool = false, jitter::Float64 = -1.0, grad::Bool = false)
    val = Ref{Float64}(0.0); gr = grad ? zeros(3) : nothing
    gcheck(g, ccall((:gpar_group_dtc_logpdf_sharded, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Float64, Ref{Float64}, Ptr{Float64}),
                    g.h, kernel_code(k), theta, vfe, jitter, val, grad ? gr : C_NULL))
    return grad ? (val[], gr) : val[]
end

# ONE scaled-GPAR objective (compute_gpar_dtc_objective, dtc.jl:83-128) whose rows are sharded over the members: every member
# holds the full (t, y) and Z and rows row_lo[i]+1 : row_lo[i]+N_i of the inputs (row_lo zero-based, multiples of 4)
function group_scaled_dtc_sharded(g::Group, k_time, k_out, theta::Vector{Float64}, row_lo::Vector{Int64}; grad::Bool = false)
    val = Ref{Float64}(0.0); gr = grad ? zeros(5) : nothing
    gcheck(g, ccall((:gpar_group_scaled_dtc_sharded, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Int64}, Ref{Float64}, Ptr{Float64}),
                    g.h, kernel_code(k_time), kernel_code(k_out), theta, row_lo, val, grad ? gr : C_NULL))
    return grad ? (val[], gr) : val[]
end

# compute_q_u (gpar_scaled_inference.jl:141-196) on the row slices resident on the members -> (m_e, inv(D), U_u)
function group_compute_q_u_sharded(g::Group, k_time, k_out, params::Vector{Float64}, row_lo::Vector{Int64}, M::Integer)
    m_e = zeros(M); Dinv = zeros(M, M); U_u = zeros(M, M)
    gcheck(g, ccall((:gpar_group_compute_q_u_sharded, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    g.h, kernel_code(k_time), kernel_code(k_out), params, row_lo, m_e, Dinv, U_u))
    return m_e, Dinv, U_u
end

# S seeded draws W = U_u \\ eps_j from the q(u) of the row-sharded evaluation (device sampler on member 0)
function group_sample_q_u_sharded(g::Group, k_time, k_out, params::Vector{Float64}, row_lo::Vector{Int64}, M::Integer, seed::Integer, S::Integer)
    W = zeros(M, S); E = zeros(M, S)
    gcheck(g, ccall((:gpar_group_sample_q_u_sharded, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Int64}, UInt64, Int32, Ptr{Float64}, Ptr{Float64}),
                    g.h, kernel_code(k_time), kernel_code(k_out), params, row_lo, seed, S, W, E))
    return W, E
end

# one whole fit (dtc.jl:58-61) on the row-sharded objective: every device works on every evaluation
function group_fit_sharded(g::Group, k_time, k_out, row_lo::Vector{Int64}, theta0::Vector{Float64}; iterations::Integer = 200, optimizer::Symbol = :neldermead)
    fmin = Ref{Float64}(0.0); xmin = zeros(5); calls = Ref{Int32}(0)
    gcheck(g, ccall((:gpar_group_fit_sharded, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Float64}, Int32, Int32, Ref{Float64}, Ptr{Float64}, Ref{Int32}),
                    g.h, kernel_code(k_time), kernel_code(k_out), row_lo, theta0, optimizer == :lbfgs ? 1 : 0, iterations, fmin, xmin, calls))
    return fmin[], xmin, Int(calls[])
end

# whole Nelder-Mead fits of the chain's conditional GPs; Xs[k] (D x N) / Zs[k] (D x M) are `nothing` for a time-only task
function group_fit(g::Group, t::Vector{Float64}, Xs, Zs, ys::Vector{Vector{Float64}}, theta0s::Vector{Vector{Float64}}, k_time, k_out; iterations::Integer = 200, optimizer::Symbol = :neldermead)
    nt = length(ys)
    GC.@preserve Xs Zs ys begin
        tasks = [FitTask(Xs[k] === nothing ? C_NULL : pointer(Xs[k]), Xs[k] === nothing ? 0 : size(Xs[k], 1),
                         Zs[k] === nothing ? C_NULL : pointer(Zs[k]), Zs[k] === nothing ? 0 : size(Zs[k], 2), pointer(ys[k]),
                         ntuple(j -> j <= length(theta0s[k]) ? theta0s[k][j] : 0.0, 5)) for k in 1:nt]
        minimum = fill(NaN, nt); minimizer = fill(NaN, 5, nt); calls = zeros(Int32, nt); member = fill(Int32(-1), nt)
        gcheck(g, ccall((:gpar_group_fit, LIB), Cint,
                        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{FitTask}, Int32, Cint, Cint, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                        g.h, t, length(t), tasks, nt, kernel_code(k_time), kernel_code(k_out), optimizer === :lbfgs ? 1 : 0, iterations, minimum, minimizer, calls, member))
    end
    return minimum, minimizer, calls, member
end

# posterior means down the chain: resident result of member `src` (or `values`) -> every member's chain buffer
function group_broadcast(g::Group, src::Integer, n::Integer; values::Union{Nothing, Vector{Float64}} = nothing)
    out = zeros(n)
    gcheck(g, ccall((:gpar_group_broadcast, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64, Ptr{Float64}),
                    g.h, src, values === nothing ? C_NULL : values, n, out))
    return out
end
set_inputs_column!(c::Ctx, d::Integer, col::Union{Nothing, Vector{Float64}} = nothing) =
    check(c, ccall((:gpar_set_inputs_column, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), c.h, d, col === nothing ? C_NULL : col))
# merged train+test problem (set_merged!): feature d (0-based) of the inputs AT THE TEST LOCATIONS <- col (N* values, test
# order) or the chain buffer filled by group_broadcast — `[test_y1, y2_out]`, GPAR_scaled_examples.jl:172
set_merged_test_column!(c::Ctx, d::Integer, col::Union{Nothing, Vector{Float64}} = nothing) =
    check(c, ccall((:gpar_set_merged_test_column, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), c.h, d, col === nothing ? C_NULL : col))
# roofline denominators of the context's device: (dmma_tflops, dfma_tflops, hbm_copy_gbs)
function measure_peaks(c::Ctx)
    a = Ref{Float64}(0); b = Ref{Float64}(0); h = Ref{Float64}(0)
    check(c, ccall((:gpar_measure_peaks, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}, Ref{Float64}), c.h, a, b, h))
    return (dmma_tflops = a[], dfma_tflops = b[], hbm_copy_gbs = h[])
end

end # module
