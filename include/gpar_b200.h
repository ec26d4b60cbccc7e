/* gpar_b200.h — C ABI of libgpar_b200.so: the B200 (sm_100a) GP linear-algebra hot path of
 * GPAR-at-scale.  This is the drop-in boundary: the reference's Julia functions keep their names
 * and signatures and `ccall` these entry points instead of calling Stheno.jl / TemporalGPs.jl /
 * LinearAlgebra (INTEGRATION.md shows the Julia stubs; the Python ctypes binding in
 * gpar-at-scale_b200/_ffi.py binds the same symbols with the same layouts).
 *
 * Conventions
 *  - every function returns an int status (GPAR_OK = 0); no exception or abort crosses the ABI;
 *    gpar_last_error(ctx) returns the message of the last failing call on that context.
 *  - all pointers in the signatures are HOST pointers owned by the caller; inputs are copied to
 *    the device by the gpar_set_* calls and stay resident across optimiser evaluations; outputs
 *    are written synchronously (the call blocks until its stream has drained).
 *  - matrices are column-major Float64, exactly the memory of a Julia Array{Float64}.  A Stheno
 *    `ColVecs` (src/util.jl:16-31) is a D x N column-major matrix = N records of D contiguous
 *    doubles; X and Z are passed in that layout.
 *  - `theta` are the RAW optimiser parameters; the library applies exp(theta)+1e-3 and the
 *    squaring of variances / noise itself (src/util.jl:36-55; dtc.jl:31-37).
 *  - kernel codes: GPAR_EQ, GPAR_MATERN12, GPAR_MATERN32, GPAR_MATERN52 (Stheno EQ / Matern12 /
 *    Matern32 / Matern52).  State-space entry points accept the three Matern kinds only.
 *  - a context is bound to one device and is not thread-safe; distinct contexts may be used
 *    concurrently from distinct host threads.
 */
#ifndef GPAR_B200_H
#define GPAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPAR_ABI_VERSION 1

enum gpar_status {
  GPAR_OK = 0,
  GPAR_ERR_INVALID = 1,      /* bad argument / missing data (message says which) */
  GPAR_ERR_CUDA = 2,         /* CUDA runtime / cuBLAS / cuSOLVER failure */
  GPAR_ERR_NOT_POSDEF = 3,   /* a Cholesky factorisation failed (Julia: PosDefException) */
  GPAR_ERR_NOMEM = 4
};

enum gpar_kernel { GPAR_EQ = 0, GPAR_MATERN12 = 1, GPAR_MATERN32 = 2, GPAR_MATERN52 = 3 };

typedef struct gpar_ctx gpar_ctx;

/* ---- context ------------------------------------------------------------------------------ */
int gpar_abi_version(void);
int gpar_ctx_create(int device, gpar_ctx** out);
int gpar_ctx_destroy(gpar_ctx* ctx);
const char* gpar_last_error(const gpar_ctx* ctx);
/* Device time (ms, CUDA events on the context's stream) of the kernels of the last compute call,
 * and how many of the library's own kernels it launched. */
int gpar_last_timing(const gpar_ctx* ctx, double* device_ms, int64_t* kernel_launches);

/* Device time of the phases of the last pseudo-point call (gpar_dtc_logpdf / gpar_scaled_dtc /
 * gpar_compute_q_u): phase_ms[0] = panel producers (Kuf evaluation, whitening), [1] = the DMMA
 * panel-SYRK kernel alone, [2] = everything after it (partial-tile reduction, M x M tail). n >= 3. */
int gpar_last_profile(const gpar_ctx* ctx, double* phase_ms, int32_t n);

/* Roofline denominators measured on this context's device (bench.py, SURVEY 8d "measure the FP64 DMMA peak on the box"):
 * issue peaks of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) and of DFMA from register-only loops, and the HBM copy bandwidth
 * (1 GiB each way, read + write bytes).  Any pointer may be NULL.  ~60 ms. */
int gpar_measure_peaks(gpar_ctx* ctx, double* dmma_tflops, double* dfma_tflops, double* hbm_copy_gbs);

/* Diagnostics: device time (ms) of the library's dense M x M routines at order n on a synthetic SPD matrix — out7 =
 * [potrf, trtri, triangular gemm, lower-tile gemm, full gemm, trsv, transposed trsv].  Measurement support only. */
int gpar_dense_bench(gpar_ctx* ctx, int32_t n, double* out7);

/* ---- resident data (host -> device copies) ------------------------------------------------ */
/* ColVecs inputs, src/gp/dtc.jl:26-27, gpar_scaled_inference.jl:38-40 (to_ColVecs, util.jl:16-31) */
int gpar_set_inputs(gpar_ctx* ctx, const double* X, int32_t D, int64_t N);
int gpar_set_pseudo(gpar_ctx* ctx, const double* Z, int32_t D, int64_t M);
/* time locations, ascending (callers sort: temporal_gp_inference.jl:61-66) */
int gpar_set_times(gpar_ctx* ctx, const double* t, int64_t N);
/* The regular grid t_k = t0 + k dt, k = 0..N-1 — the Julia shim calls this for an AbstractRange
 * (range(0, step = 1/30, length = N), toy_data.jl:6).  TemporalGPs gives a range constant A, Q
 * (RegularSpacing); so does the library: the transition matrix is built once per sequence, not per step. */
int gpar_set_times_range(gpar_ctx* ctx, double t0, double dt, int64_t N);
/* outputs: `batch` sequences of N values, sequence b at y + b*N */
int gpar_set_outputs(gpar_ctx* ctx, const double* y, int64_t N, int32_t batch);
/* per-step observation noise R_k (the 1e10 trick, temporal_gp_inference.jl:93-97); NULL clears it */
int gpar_set_noise_vector(gpar_ctx* ctx, const double* r, int64_t N);

/* The merge / sort / un-sort protocol of get_sde_predictions (temporal_gp_inference.jl:55-66,93-97,111-112)
 * and get_gpar_scaled_predictions (gpar_scaled_inference.jl:75-87,100-103,132-133) on the device:
 *   times = vcat(t, ts)[perm] with perm = sortperm (stable: at equal times the training point comes first),
 *   outputs = vcat(y, zeros(Ns))[perm], noise vector = vcat(fill(sigma2, N), fill(1e10, Ns))[perm] and, when
 *   D > 0, inputs = vcat(X, Xs)[perm] become the resident times / outputs / noise vector / inputs.
 * X, Xs: N x D and Ns x D records (NULL with D = 0 for the time-only model).  After a smoother or a
 * prediction on the merged problem (their host outputs may then be NULL), gpar_take_test returns the
 * entries at the Ns test locations in the order of `ts` (result[reverse_perm][N+1:end]). */
int gpar_set_merged(gpar_ctx* ctx, const double* t, const double* y, const double* X, int64_t N,
                    const double* ts, const double* Xs, int64_t Ns, int32_t D, double sigma2);
/* a_test, b_test (Ns each; b_test nullable): (mean, var) of the last gpar_lgssm_smooth (batch 1) or
 * (mean, std) of the last gpar_scaled_predict, at the test locations. */
int gpar_take_test(gpar_ctx* ctx, double* a_test, double* b_test);

/* ---- pseudo-point approximation, diagonal noise: Stheno dtc / elbo --------------------------
 * Replaces Stheno's `dtc(f(X, sigma^2), y, u)` / `elbo(...)` (restated at
 * examples/dtc_example.jl:10-23) for f = GP(kernel(k; l, s = var^2)), theta = (log l, log var,
 * log sigma) as unpack_gp (util.jl:36-43).  cov(u) = Kuu + jitter*I; jitter < 0 means "use
 * sigma^2", the reference's own convention (u = gp_prior(Z, noise_sigma^2), dtc.jl:35,
 * dtc_example.jl:46-47) and then the gradient flows through it.
 * vfe = 0: DTC; vfe = 1: Titsias bound (DTC - 1/2 (tr(Kff)/sigma^2 - ||A||_F^2)).
 * grad (nullable): d val / d theta[0..2].  Needs set_inputs, set_pseudo, set_outputs(batch=1). */
int gpar_dtc_logpdf(gpar_ctx* ctx, int kernel, const double theta[3], int vfe, double jitter,
                    double* val, double* grad);

/* The same objective with, additionally, its gradient with respect to the pseudo-inputs: grad_Z is M x D
 * records like Z (NEW — the reference keeps Z fixed, dtc_example.jl:82-90; pseudo-input optimisation is what the
 * Titsias bound vfe = 1 is for).  Matern-1/2 is not differentiable where a pseudo-input coincides with an
 * input: that pair contributes 0 (a subgradient). */
int gpar_dtc_logpdf_zgrad(gpar_ctx* ctx, int kernel, const double theta[3], int vfe, double jitter,
                          double* val, double* grad_theta, double* grad_Z);

/* ---- scaled GPAR objective: compute_gpar_dtc_objective, src/gp/dtc.jl:83-128 ---------------
 * theta = unpack_gpar parameters (util.jl:45-55).  Returns dtc (the nlml closure dtc.jl:29-47
 * returns -dtc).  A_or_null: optional M x N column-major output of the `A` the reference also
 * returns (dtc.jl:119,127).  Needs set_inputs, set_pseudo, set_times, set_outputs(batch=1). */
int gpar_scaled_dtc(gpar_ctx* ctx, int k_time, int k_out, const double theta[5], double* dtc,
                    double* A_or_null);

/* `ncand` hyper-parameter candidates of the same objective on the SAME resident data (NEW): thetas is
 * 5 x ncand column-major — the simplex vertices x restarts that the Nelder-Mead loop dtc.jl:58-61 evaluates one after
 * the other.  When the M x M tail fits shared memory (M <= ~110: the reference's own sizes) all candidates share one fused
 * launch sequence (values agree with gpar_scaled_dtc to ~1e-12 relative); otherwise they run concurrently on up to 16
 * internal lanes (streams + scratch of their own) of this device, bit-identical to gpar_scaled_dtc.  codes (nullable, ncand): 0 or GPAR_ERR_NOT_POSDEF (value NaN); with
 * codes == NULL a failed Cholesky fails the call.  gpar_last_timing reports the wall-clock ms of the batch. */
int gpar_scaled_dtc_batch(gpar_ctx* ctx, int k_time, int k_out, const double* thetas, int32_t ncand, double* dtc, int32_t* codes);

/* The same objective with its gradient d dtc / d theta[0..4] (NEW — the reference optimises it with
 * Nelder-Mead, dtc.jl:58-61; a gradient lets Optim.LBFGS replace it). */
int gpar_scaled_dtc_grad(gpar_ctx* ctx, int k_time, int k_out, const double theta[5], double* dtc, double* grad);

/* compute_q_u, src/gp/gpar_scaled_inference.jl:141-196: m_e (M), inv(D) (M x M), U_u (M x M upper).
 * params are the POSITIVE (already unpacked) values (time_l, time_var, out_l, out_var, noise_sigma),
 * as the reference passes kernels built from opt_params (gpar_scaled_inference.jl:57-73). */
int gpar_compute_q_u(gpar_ctx* ctx, int k_time, int k_out, const double params[5], double* m_e,
                     double* Dinv, double* U_u);

/* Seeded draws from q_u on the device (replaces `rand(q_u)` + `U_u \ eps`, gpar_scaled_inference.jl:94-96,
 * whose RNG is Julia's unseeded global one): S draws eps_j ~ MvNormal(m_e, inv(D)) from a Philox4x32-10
 * stream keyed by `seed`, W[:, j] = U_u \ eps_j.  Needs the TRAINING data resident, like gpar_compute_q_u.
 * W stays on the device for a following gpar_scaled_predict(..., W = NULL, S, ...); W_out / eps_out
 * (M x S column-major, nullable) receive host copies. */
int gpar_sample_q_u(gpar_ctx* ctx, int k_time, int k_out, const double params[5], uint64_t seed, int32_t S,
                    double* W_out, double* eps_out);

/* Monte-Carlo prediction of one scaled-GPAR output, get_gpar_scaled_predictions
 * (gpar_scaled_inference.jl:89-135), batched: resident data are the MERGED, time-SORTED train+test
 * inputs X* (set_inputs), times (set_times), outputs with 0 at test points (set_outputs, batch 1),
 * noise vector sigma^2 / 1e10 (set_noise_vector) and Z.  W (M x S column-major): column j =
 * U_u \ eps_j for the caller's draws eps_j ~ q_u (:91-97; the RNG stays with the caller).
 * Outputs (length N+N*, still in sorted order): sample mean and corrected std over the S draws of
 * f*_j = fx_j + smooth(y* - fx_j).m[1] (:113-125).  params: positive values as gpar_compute_q_u.
 * W == NULL: use the device-resident weights of the last gpar_sample_q_u (same M and S).
 * mean = std = NULL: results stay on the device (gpar_take_test). */
int gpar_scaled_predict(gpar_ctx* ctx, int k_time, int k_out, const double params[5], const double* W,
                        int32_t S, double* mean, double* std);

/* ---- state-space approximation: TemporalGPs logpdf / decorrelate / smooth ------------------
 * Model of create_lgssm (temporal_gp_inference.jl:15-39): GP(kernel(k; l, s = var^2)) -> SDE ->
 * LGSSM on the resident times with noise sigma^2, or the resident noise vector if one is set.
 * theta = unpack_gp parameters (log l, log var, log sigma).
 * gpar_lgssm_logpdf: `batch_theta` parameter sets (theta + 3*b); with batch_theta == outputs batch,
 *   sequence b uses theta b (independent models); with batch_theta == 1 all sequences share it.
 *   lml: one value per sequence (temporal_gp_inference.jl:78).
 *   With ONE resident sequence and batch_theta > 1 the parameter sets are hyper-parameter CANDIDATES (simplex vertices x
 *   restarts of the Nelder-Mead loop, temporal_gp_inference.jl:82) evaluated on that sequence in one pass: lml[batch_theta]. */
int gpar_lgssm_logpdf(gpar_ctx* ctx, int kernel, const double* theta, int32_t batch_theta, double* lml);
/* The same log-pdf with its gradient (NEW — the reference's optimisers are derivative-free,
 * temporal_gp_inference.jl:82): grad + 3*b = d lml[b] / d theta of the parameter set sequence b uses
 * (its own, or the shared one).  With a resident noise vector d/d theta[2] = 0 (sigma is unused). */
int gpar_lgssm_logpdf_grad(gpar_ctx* ctx, int kernel, const double* theta, int32_t batch_theta, double* lml, double* grad);
/* decorrelate (dtc.jl:106,115): alpha (N x batch) and lml (batch) for every resident sequence. */
int gpar_lgssm_decorrelate(gpar_ctx* ctx, int kernel, const double theta[3], double* alpha, double* lml);
/* smooth (temporal_gp_inference.jl:109; gpar_scaled_inference.jl:117): mean = m_s[1], var = P_s[1,1]
 * per step and sequence (N x batch each), lml (batch, nullable).  mean = var = NULL: results stay on the
 * device (gpar_take_test). */
int gpar_lgssm_smooth(gpar_ctx* ctx, int kernel, const double theta[3], double* mean, double* var, double* lml);

/* ---- exact GP / GPAR (dense), src/gp/optimized.jl ------------------------------------------
 * ntheta = 3: GP on the resident inputs (any D) with kernel k_time (optimized.jl:28-36);
 * ntheta = 5: GPAR kernel time_var^2 k_time(|dx_1|/time_l) + out_var^2 k_out(||dx_2:D||/out_l)
 * (optimized.jl:132-154).  logpdf per resident output sequence. */
int gpar_exact_logpdf(gpar_ctx* ctx, int k_time, int k_out, const double* theta, int32_t ntheta, double* lml);
/* `ncand` hyper-parameter candidates at once (NEW): thetas is ntheta x ncand column-major — e.g. the vertices
 * of a Nelder-Mead simplex times the restarts of optimized.jl:45,164 — evaluated on the SAME resident data in one launch
 * (one CTA per candidate when N <= 200).  lml[c * batch + b]; codes (nullable, ncand): 0, or non-zero where the candidate's
 * Cholesky failed (its lml entries are NaN) — with codes == NULL such a failure returns GPAR_ERR_NOT_POSDEF. */
int gpar_exact_logpdf_batch(gpar_ctx* ctx, int k_time, int k_out, const double* thetas, int32_t ntheta, int32_t ncand,
                            double* lml, int32_t* codes);
/* posterior marginals at Xs (D x Ns ColVecs): mean and variance (optimized.jl:94,236; eeg.jl:185-208) */
int gpar_exact_posterior(gpar_ctx* ctx, int k_time, int k_out, const double* theta, int32_t ntheta,
                         const double* Xs, int64_t Ns, double* mean, double* var);

/* ---- several devices of one box from ONE host process --------------------------
 * The shards are the reference's own: per-output conditional GPs (examples/GPAR_scaled_examples.jl:132-175 fits
 * output i on the OBSERVED outputs < i, so the fits are independent) and hyper-parameter restarts (util.jl:128-134).
 * A group owns one context per device; load each member's data with the gpar_set_* calls on gpar_group_ctx(g, i).
 * No collective touches the data path of these task-sharded calls (the *_sharded entry points below shard the ROWS of one
 * objective instead and exchange slice summaries / M x M statistics): NCCL (bound at run time from libnccl.so.2) all-gathers the scalars —
 * (status, value, gradient) or (minimum, minimiser) — so that every device holds the table, and broadcasts posterior
 * means down the GPAR chain (GPAR_scaled_examples.jl:172).  Group calls are blocking and not thread-safe. */
typedef struct gpar_group gpar_group;
int gpar_group_create(const int32_t* devices, int32_t ndev, gpar_group** out);   /* distinct device ordinals */
int gpar_group_destroy(gpar_group* g);
int32_t gpar_group_size(const gpar_group* g);
gpar_ctx* gpar_group_ctx(gpar_group* g, int32_t member);                         /* borrowed, owned by the group */
const char* gpar_group_last_error(const gpar_group* g);

/* Member i evaluates thetas[:, i] (3 x ndev / 5 x ndev column-major) on ITS resident data, all members concurrently
 * (one host thread per device).  vals[ndev]; grads (nullable) 3 x ndev / 5 x ndev; codes (nullable) receives each
 * member's status — GPAR_ERR_NOT_POSDEF of a member does not fail the call (an optimiser reads it as +Inf). */
int gpar_group_dtc_logpdf(gpar_group* g, int kernel, const double* thetas, int vfe, double jitter,
                          double* vals, double* grads, int32_t* codes);             /* gpar_dtc_logpdf per member */
int gpar_group_scaled_dtc(gpar_group* g, int k_time, int k_out, const double* thetas,
                          double* vals, double* grads, int32_t* codes);             /* gpar_scaled_dtc(_grad) per member */

/* ONE plain DTC / VFE objective (gpar_dtc_logpdf) whose data rows are sharded over the members — SURVEY 8e's optional
 * intra-output N-sharding: load slice i of (X, y) and the SAME pseudo-inputs on member i; each member evaluates the
 * sufficient statistics of its slice, one ncclAllReduce (8 (2 M^2 + 2 M + 1) bytes) sums them over NVLink, member 0
 * runs the M x M tail.  (The row-sharded entry points are the only ones with a data-path collective.) */
int gpar_group_dtc_logpdf_sharded(gpar_group* g, int kernel, const double theta[3], int vfe, double jitter, double* val, double* grad);

/* ONE scaled-GPAR objective (gpar_scaled_dtc; compute_gpar_dtc_objective, src/gp/dtc.jl:83-128) whose ROWS are sharded
 * over the members: member i holds the FULL times and outputs (gpar_set_times / gpar_set_outputs, N_full each — 16 bytes
 * per step; the 1 x N filter of dtc.jl:106 is cheap and every member runs it), the same pseudo-inputs, and rows
 * [row_lo[i], row_lo[i] + N_i) of the inputs (gpar_set_inputs with that slice; consecutive slices covering N_full, every
 * row_lo[i] a multiple of 4).  The N x M work — kernel panel, the M whitening passes of dtc.jl:110-117, A A' of :120 — is
 * sliced; the filter carry of the M columns crosses the slice boundaries in ONE all-gather of slice summaries
 * (D x D transition product + D x M exit state per member), and ONE all-reduce sums (beta'beta, beta'alpha) before member 0
 * runs the M x M tail.  A poorly conditioned cov(u) is handled as on one device (every member whitens its panel by L_u).
 * GPAR_GROUP_LOOPBACK=1 lets gpar_group_create put several members on ONE device (collectives become device copies).
 * grad (nullable): the five derivatives of gpar_scaled_dtc_grad — forward-mode tangents per slice, a second all-gather of three
 * tangent summaries; a poorly conditioned cov(u) switches every slice to whitened coordinates (a whitened copy of its panel). */
int gpar_group_scaled_dtc_sharded(gpar_group* g, int k_time, int k_out, const double theta[5], const int64_t* row_lo, double* val, double* grad);

/* The same row-sharded evaluation for hosts that run ONE PROCESS PER DEVICE and own the collectives (torch.distributed over
 * NCCL — gpar-at-scale_b200/parallel.py; MPI; Julia `Distributed` workers): the context holds the full (t, y), the pseudo-inputs
 * and rows [row_lo, row_lo + N) of the inputs, exactly as a member of gpar_group_scaled_dtc_sharded.  Every pointer named *_dev
 * is a DEVICE pointer on the context's device; every call returns after its work is complete.
 *   begin(theta, row_lo, want_grad) -> counts;  summary(summary_dev[summary_count])
 *   [all-gather the summaries, rank-major -> gathered_dev[world * summary_count]]
 *   stats(gathered_dev, member = rank, stats_dev[stats_count])       [all-reduce (sum) stats_dev]
 *   value(stats_dev, &val)                                           — any (every) rank
 * gradient (want_grad != 0) instead of value():
 *   tangent_summary(stats_dev, summary2_dev[3 * summary_count])      [all-gather -> gathered2_dev[world * 3 * summary_count]]
 *   grad_partial(gathered2_dev, member, s5[5] HOST)                  [sum the five numbers over the ranks]
 *   grad_finish(s5_total, &val, grad[5]) */
int gpar_scaled_slice_begin(gpar_ctx* ctx, int k_time, int k_out, const double theta[5], int64_t row_lo, int32_t want_grad,
                            int64_t* summary_count, int64_t* stats_count);
int gpar_scaled_slice_summary(gpar_ctx* ctx, double* summary_dev);
int gpar_scaled_slice_stats(gpar_ctx* ctx, const double* gathered_dev, int32_t member, double* stats_dev);
int gpar_scaled_slice_value(gpar_ctx* ctx, const double* stats_dev, double* val);
int gpar_scaled_slice_tangent_summary(gpar_ctx* ctx, const double* stats_dev, double* summary2_dev);
int gpar_scaled_slice_grad_partial(gpar_ctx* ctx, const double* gathered2_dev, int32_t member, double s5[5]);
int gpar_scaled_slice_grad_finish(gpar_ctx* ctx, const double s5_total[5], double* val, double grad[5]);

/* One conditional-GP fit of the chain: inputs X (D x N ColVecs = the observed earlier outputs; D = 0: a time-only
 * state-space GP with 3 parameters, temporal_gp_inference.jl:69-82), pseudo-inputs Z (D x M), outputs y (N),
 * start point theta0 (the first 3 or 5 entries are used). */
typedef struct gpar_fit_task {
  const double* X; int32_t D;
  const double* Z; int64_t M;
  const double* y;
  double theta0[5];
} gpar_fit_task;
/* Fits (Nelder-Mead: src/gp/dtc.jl:58-61 with an iteration budget instead of a time limit) of all tasks on the common
 * sorted time grid t (N): members take tasks dynamically, longest first.  minimum[ntasks] (of the NEGATED objective,
 * as the reference minimises), minimizer 5 x ntasks (NaN-padded), f_calls / member_of (nullable) per task. */
enum gpar_optimizer { GPAR_OPT_NELDER_MEAD = 0, GPAR_OPT_LBFGS = 1 };   /* L-BFGS uses the library's analytic gradients */
int gpar_group_fit(gpar_group* g, const double* t, int64_t N, const gpar_fit_task* tasks, int32_t ntasks, int k_time, int k_out,
                   int32_t optimizer, int32_t iterations, double* minimum, double* minimizer, int32_t* f_calls, int32_t* member_of);

/* compute_q_u (src/gp/gpar_scaled_inference.jl:141-196; gpar_compute_q_u) with the rows sharded over the members exactly like
 * gpar_group_scaled_dtc_sharded (same resident slices, same two collectives; bare Cuu — every member whitens its panel by L_u when
 * it is poorly conditioned).  m_e[M], Dinv[M x M], U_u[M x M] column-major.  A single output whose N x M panel exceeds one device can
 * then be fitted (gpar_group_fit_sharded) and predicted: sampling q(u) and gpar_scaled_predict need no N x M array. */
int gpar_group_compute_q_u_sharded(gpar_group* g, int k_time, int k_out, const double params[5], const int64_t* row_lo,
                                   double* m_e, double* Dinv, double* U_u);
/* ... and S seeded draws from that q(u) (gpar_sample_q_u's device sampler on member 0): W = U_u \ eps_j (M x S column-major, host
 * copy; nullable) for gpar_scaled_predict, eps_out (nullable). */
int gpar_group_sample_q_u_sharded(gpar_group* g, int k_time, int k_out, const double params[5], const int64_t* row_lo, uint64_t seed, int32_t S,
                                  double* W_out, double* eps_out);

/* ONE fit with every device working on every evaluation: the optimiser of gpar_group_fit on the ROW-SHARDED scaled objective
 * (gpar_group_scaled_dtc_sharded; the slices — full (t, y), Z, the member's rows of X — are already resident).  For a single output
 * too large for one device; replaces the loop src/gp/dtc.jl:58-61 for that output.  minimum (of the NEGATED objective), minimizer[5],
 * f_calls (nullable). */
int gpar_group_fit_sharded(gpar_group* g, int k_time, int k_out, const int64_t* row_lo, const double theta0[5], int32_t optimizer, int32_t iterations,
                           double* minimum, double* minimizer, int32_t* f_calls);

/* n doubles from member src into EVERY member's chain buffer (ncclBroadcast over NVLink): `host` if given, else the
 * resident result of its last gpar_lgssm_smooth / gpar_scaled_predict — for a merged train+test problem
 * (gpar_set_merged) the posterior means AT THE N* TEST LOCATIONS IN TEST ORDER (what gpar_take_test returns first;
 * n must be N*), otherwise the first n values of the result.  out (nullable) receives a host copy read from a
 * receiving member. */
int gpar_group_broadcast(gpar_group* g, int32_t src, const double* host, int64_t n, double* out);
/* Column d of the resident inputs X <- col (host, N values) or, with col == NULL, the context's chain buffer: the
 * predicted means of an earlier output become an input feature of the later outputs without a host round trip. */
int gpar_set_inputs_column(gpar_ctx* ctx, int32_t d, const double* col);
/* The same for a MERGED problem (gpar_set_merged with D > 0): column d of the inputs at the N* test locations <- col
 * (host, N* values in test order) or the chain buffer — the chain step `[test_y1, y2_out]` of
 * GPAR_scaled_examples.jl:172 device to device: gpar_group_broadcast(g, owner, NULL, N*, NULL) after the earlier
 * output's prediction, then gpar_set_merged (any placeholder in column d of Xs) + this call on the later output. */
int gpar_set_merged_test_column(gpar_ctx* ctx, int32_t d, const double* col);

#ifdef __cplusplus
}
#endif
#endif /* GPAR_B200_H */
