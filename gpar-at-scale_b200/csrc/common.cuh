// common.cuh — context, device buffers and error plumbing shared by the kernels of libgpar_b200.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/gpar_b200.h"

#define GPAR_TILE 128          // M-tile of the panel layout and of the DMMA kernel
#define GPAR_KT 32             // n-steps per pipeline stage of the DMMA kernel

// Growable device buffer (never shrinks; freed with the context).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool borrowed = false;          // a view of another context's resident buffer (lanes of the batched entry points): never freed here
  void borrow(const DevBuf& o) { if (!borrowed) release(); p = o.p; cap = o.cap; borrowed = true; }
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (borrowed) return cudaErrorInvalidValue;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct gpar_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;     // side stream: the G-independent part of the M x M tail overlaps the SYRK
  cudaEvent_t ev_fork = nullptr, ev_side = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t pev[4] = {nullptr, nullptr, nullptr, nullptr};   // phase marks: producers done, main kernel start/end
  bool phase_valid = false;
  int num_sms = 148;
  std::string err;
  double last_ms = 0.0;
  int64_t last_launches = 0;
  int64_t launches = 0;      // running counter, reset at the start of every compute call

  // resident data
  DevBuf X, Z, t, y, rvec;
  int32_t D = 0, Dz = 0, ybatch = 0;
  int64_t N = 0, M = 0, Nt = 0, Ny = 0, Nr = 0;
  bool has_rvec = false;
  bool y_broadcast = false;           // transient: the running lgssm call evaluates parameter candidates on ONE resident sequence
  bool ss_deferred_ok = false;        // the running entry point checks lgssm_steady_failed() after its final synchronisation
  bool ss_pending = false;            // the single-pass steady-state path ran: its flags are on their way to ctx->pinned
  int ss_skip = 0;                    // calls for which the steady-state Kalman path stays off after a non-converged hand-over
  double t_reg_dt = 0.0;              // > 0: the resident times are the regular grid t0 + k dt (gpar_set_times_range)

  // scratch (grown on demand)
  DevBuf panelK, panelD, panelB, panelA, kal_f, partial, segs, jobs, gpart, scal, dense, tailws, info;
  DevBuf kal_a, kal_b, kal_c, kal_d, kal_e;
  // merged train+test problem (gpar_set_merged): staging, position of every test point in the sorted arrays,
  // and the device arrays of the last smoother / prediction result (for gpar_take_test)
  DevBuf mrg, test_pos; int64_t merged_N = 0, merged_Ns = 0; const double* res_a = nullptr; const double* res_b = nullptr; int64_t res_len = 0;
  DevBuf shbuf;                       // shared-model smoother: tables, chunk states, filtered means (smooth_shared.cu)
  bool dla_optin = false;
  DevBuf dla_ws, dla_ws_side, dla_ws2;                // scratch of the dense M x M routines (dense_la.cu): main / side stream, transposes
  DevBuf qW; int64_t qW_M = 0; int32_t qW_S = 0;     // resident W = U_u \ eps of the last gpar_sample_q_u
  DevBuf chain; int64_t chain_n = 0;                  // values passed down the GPAR chain (gpar_group_broadcast / gpar_set_inputs_column)
  void* pinned = nullptr; size_t pinned_cap = 0;
  // pinned staging of the fused small-problem sequence (LGSSM parameters, candidates, results): fixed host addresses, so
  // that the sequence can be replayed as a CUDA graph; param_staging (when set) is where kalman.cu packs (l, s, noise)
  void* small_pin = nullptr; size_t small_pin_cap = 0;
  double* param_staging = nullptr;
  // the captured sequence of the last (shape, pointers) key and its bookkeeping (scaled_small.cu)
  struct SmallGraph {
    unsigned long long key[20] = {0}; bool have_key = false;      // key of `exec` / of the last plain run
    int warm = 0;                                                // plain runs seen with this key (capture on the second)
    bool failed = false;                                         // capture or instantiation failed for this key: stay plain
    cudaGraphExec_t exec = nullptr;
  } sgraph;
  // lanes: worker contexts on the same device (own streams and scratch, resident data borrowed from this context) that
  // evaluate hyper-parameter candidates concurrently (gpar_scaled_dtc_batch)
  std::vector<gpar_ctx*> lanes;
  // one row slice of a scaled objective sharded over a group (scaled.cu: scaled_slice_phase1 / 2 / finish)
  struct SliceState {
    int D = 0, Mpad = 0, nch = 0, whg = 0, CT = 1; int64_t lo = 0, Npad = 0; bool robust = false;
    const double *table = nullptr, *alpha = nullptr; double *sums = nullptr, *resp = nullptr, *psi = nullptr, *gp = nullptr;
    double *summary = nullptr, *init = nullptr, *G = nullptr, *g = nullptr; size_t summary_count = 0, stats_count = 0;
    // gradient mode (scaled_slice_grad_*): tangent tables of the slice, the tangent summaries / entering tangent states, and the
    // host-side pieces of the final assembly
    bool grad = false, begun = false, qu = false; int64_t nfull = 0; int k_out = 0;
    const double *dtable = nullptr, *dalpha = nullptr; double *evec = nullptr, *panelD = nullptr, *summary2 = nullptr, *init2 = nullptr;
    double pv[5] = {0}, ex[5] = {0}, val = 0.0, raw[8 + 20] = {0}, fsums[6] = {0}, wq[4] = {0};      // wq: tr Q, <W,Q>, <X,Q>, c'c (whitened tail)
  } slice;
  // SYRK plan cache: the (tiles, k-blocks, with_h) of the plan currently resident in `segs`/`jobs`
  int plan_T = -1, plan_h = -1, plan_C = 0, plan_J = 0; int64_t plan_NBK = -1; size_t plan_nseg = 0;
};

int gpar_fail(gpar_ctx* c, int code, const char* fmt, ...);

// Brackets the KERNELS of one compute call with CUDA events on the context's stream; call stop()
// before the device->host copies of the results so that gpar_last_timing reports device work only.
struct CallTimer {
  gpar_ctx* c; bool stopped = false;
  explicit CallTimer(gpar_ctx* ctx) : c(ctx) { c->launches = 0; c->slice.begun = false; cudaEventRecord(c->ev0, c->stream); }      // any other compute call ends a row-slice evaluation (shared scratch)
  void stop() { if (!stopped) { cudaEventRecord(c->ev1, c->stream); stopped = true; } }
  ~CallTimer() {
    stop();
    if (cudaEventSynchronize(c->ev1) == cudaSuccess) { float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); c->last_ms = ms; }
    c->last_launches = c->launches;
  }
};

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return gpar_fail(ctx, e_ == cudaErrorMemoryAllocation ? GPAR_ERR_NOMEM : GPAR_ERR_CUDA, \
                       "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CHK(call) do { int r_ = (call); if (r_ != GPAR_OK) return r_; } while (0)

// Every kernel of this library is launched through this macro so that launches are counted.
#define LAUNCH(ctx, kern, grid, block, smem, ...)                                          \
  do {                                                                                    \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                          \
    (ctx)->launches++;                                                                    \
    CU(cudaGetLastError());                                                               \
  } while (0)

// Programmatic dependent launch: the kernel may be scheduled while its predecessor on the stream still runs; it must
// execute pdl_wait() before it touches anything the predecessor writes.  A predecessor that calls pdl_trigger() early
// lets the dependent's CTAs become resident (and run their prologue) underneath it.
#define LAUNCH_PDL(ctx, kern, grid, block, smem, ...)                                       \
  do {                                                                                    \
    cudaLaunchConfig_t cfg_ = {};                                                         \
    cfg_.gridDim = (grid); cfg_.blockDim = (block); cfg_.dynamicSmemBytes = (smem); cfg_.stream = (ctx)->stream; \
    cudaLaunchAttribute at_[1];                                                           \
    at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at_[0].val.programmaticStreamSerializationAllowed = 1; \
    cfg_.attrs = at_; cfg_.numAttrs = 1;                                                  \
    CU(cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__));                                     \
    (ctx)->launches++;                                                                    \
  } while (0)
__device__ __forceinline__ void pdl_wait() {
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;");
#endif
}

struct GpParams { double l, var, s, sigma, noise; double dl, ds_dv, dn; };
// exp(theta)+1e-3 (src/util.jl:36-43), s = var^2, noise = sigma^2; d*/dtheta chain factors.
static inline GpParams unpack_gp3(const double* th) {
  GpParams p;
  p.l = exp(th[0]) + 1e-3; p.var = exp(th[1]) + 1e-3; p.sigma = exp(th[2]) + 1e-3;
  p.s = p.var * p.var; p.noise = p.sigma * p.sigma;
  p.dl = exp(th[0]); p.ds_dv = 2.0 * p.var * exp(th[1]); p.dn = 2.0 * p.sigma * exp(th[2]);
  return p;
}

// ---- device helpers -----------------------------------------------------------------------
// Stheno base kernels as a function of d2 = ||x - z||^2 / l^2 (src: oracle/kernels.py).
// Returns kappa(r); *dkdl_l receives  l * d kappa / d l  (= -r kappa'(r)), used for gradients.
template <int KIND, bool GRAD>
__device__ __forceinline__ double base_kernel_dev(double d2, double& l_dkdl) {
  if (KIND == GPAR_EQ) {
    double k = exp(-0.5 * d2);
    if (GRAD) l_dkdl = d2 * k;                      // d/dl exp(-r0^2/(2 l^2)) = r0^2/l^3 k
    return k;
  } else if (KIND == GPAR_MATERN12) {
    double r = sqrt(d2), k = exp(-r);
    if (GRAD) l_dkdl = r * k;
    return k;
  } else if (KIND == GPAR_MATERN32) {
    double a = sqrt(3.0 * d2), e = exp(-a);
    if (GRAD) l_dkdl = a * a * e;                   // -a d/da[(1+a)e^-a] = a^2 e^-a
    return (1.0 + a) * e;
  } else {
    double a = sqrt(5.0 * d2), e = exp(-a);
    if (GRAD) l_dkdl = a * a * (1.0 + a) * e * (1.0 / 3.0);  // -a d/da[(1+a+a^2/3)e^-a] = a^2(1+a)e^-a/3
    return (1.0 + a + a * a * (1.0 / 3.0)) * e;
  }
}

// The same kernels from r = ||x - z|| / l directly: with one input dimension r = |x - z| / l needs no FP64 square
// root (~15 of the ~60 FP64 issue slots of a Matern-5/2 element), which the panel producers exploit for D = 1.
template <int KIND, bool GRAD>
__device__ __forceinline__ double base_kernel_from_r(double r, double& l_dkdl) {
  if (KIND == GPAR_EQ) {
    double k = exp(-0.5 * r * r);
    if (GRAD) l_dkdl = r * r * k;
    return k;
  } else if (KIND == GPAR_MATERN12) {
    double k = exp(-r);
    if (GRAD) l_dkdl = r * k;
    return k;
  } else if (KIND == GPAR_MATERN32) {
    double a = 1.7320508075688772935274463415059 * r, e = exp(-a);
    if (GRAD) l_dkdl = a * a * e;
    return (1.0 + a) * e;
  } else {
    double a = 2.2360679774997896964091736687313 * r, e = exp(-a);
    if (GRAD) l_dkdl = a * a * (1.0 + a) * e * (1.0 / 3.0);
    return (1.0 + a + a * a * (1.0 / 3.0)) * e;
  }
}

// deterministic block reduction of one double (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < nw ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in warp 0
}

// ---- module entry points (host) -----------------------------------------------------------
// kuf_panel.cu
int launch_kuf_panels(gpar_ctx* ctx, int kind, bool grad, double l, double s,
                      double* panelK, double* panelD, double* gpart, int nsplit, int64_t Npad, int Mpad);
int launch_reduce_gh(gpar_ctx* ctx, const double* gpart, int nsplit, int Mpad, int nvec, double* out);
// panel_syrk.cu
int panel_syrk_run(gpar_ctx* ctx, const double* panelK, const double* panelD, int64_t Npad, int Mpad,
                   int M, bool with_h, double* G, double* H);
// kalman.cu: filter / smoother on device buffers.  hl/hs/hn: host arrays (nparam = 1 or batch) of
// positive (l, s, noise).  table (nullable, batch == 1): per-step [Phi (D*D), K (D), HA (D), 1/sqrt(S)].
// sums (nullable): per sequence (sum log S, sum alpha^2).
int lgssm_run(gpar_ctx* ctx, int kind, const double* hl, const double* hs, const double* hn, int nparam, int batch, int64_t N,
              const double* t, const double* y, const double* rvec, double* d_alpha, double* d_lml, double* d_mean, double* d_var,
              double* d_table, double* d_sums, double* d_table_fwd = nullptr);      // (smoother runs: d_table = backward table, d_table_fwd = filter table)
// forward-mode variant: values and two tangents (dirs = tangent slot 0/1/-1 of l, s, noise) in one pass
int lgssm_run_tangent(gpar_ctx* ctx, int kind, const double* hl, const double* hs, const double* hn, int nparam, int batch, int64_t N,
                      const double* t, const double* y, const double* rvec, const int dirs[3],
                      double* d_alpha, double* d_lml, double* d_dlml, double* d_sums, double* d_dalpha, double* d_table, double* d_dtable);
// scaled.cu: e = a - panel w (panel in the operand layout)
int launch_panel_residual(gpar_ctx* ctx, const double* panel, const double* w, const double* a, int64_t N, int64_t NB4, int T, int M, double* e);
// panel_gemm.cu: out[n, m'] = sum_m C[m', m] in[n, m] on DMMA (C pre-arranged by launch_dense_to_operand / launch_tri_operand)
int launch_dense_to_operand(gpar_ctx* ctx, const double* Q, int M, int Mpad, double* Aop);
int panel_gemm_run(gpar_ctx* ctx, const double* Aop, int Mpad, const double* in, int64_t in_groups, double* out, int64_t out_groups,
                   int64_t g_lo, int64_t ng, int64_t out_g0, int mt_lo, int n_mt, bool triangular, const double* in_diag = nullptr);
int launch_tri_operand(gpar_ctx* ctx, const double* L, int M, int Mpad, double* Yd, double* Aop);
int panel_tri_solve_run(gpar_ctx* ctx, const double* Aop, int Mpad, double* panel, int64_t groups, int64_t g_lo, int64_t ng, const double* src = nullptr);
// smooth_shared.cu: smoothed means of Sp (multiple of 128) TIME-MAJOR sequences yt[n][Sp] that share one model
int lgssm_smooth_shared(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, const double* t, const double* y1,
                        const double* rvec, const double* yt, int Sp, double* mean_t);
int lgssm_smooth_shared_seqmajor(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, int batch, const double* t,
                                 const double* y, const double* rvec, double* d_mean, double* d_var, double* d_lml);
int lgssm_filter_shared_seqmajor(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, int batch, const double* t,
                                 const double* y, const double* rvec, double* d_alpha, double* d_lml);
// merge.cu: result of the last smoother / prediction on a merged problem, gathered at the test locations (device arrays)
int merged_gather_test(gpar_ctx* ctx, double* dst_a, double* dst_b);
// Entry points that overwrite the buffers the resident result lives in call this first, so that a later gpar_take_test /
// gpar_group_broadcast fails loudly instead of reading stale memory.
static inline void gpar_drop_result(gpar_ctx* c) { c->res_a = nullptr; c->res_b = nullptr; c->res_len = 0; c->slice.begun = false; }      // (a running row-slice evaluation points into the same scratch)
// abi.cu: sufficient statistics of the plain DTC objective over the context's resident data slice (async on its stream)
// whiten_vfe >= 0: the panels are whitened by the L_u this context's dtc_tail_prepare(vfe = whiten_vfe) computed, before the SYRK
int dtc_slice_stats(gpar_ctx* ctx, int kernel, const GpParams& p, bool want_grad, double** stats, size_t* count, int whiten_vfe = -1);
// dense_la.cu: hand-written dense linear algebra on column-major FP64 device matrices (no cuBLAS / cuSOLVER anywhere in
// the library), enqueued on ctx->stream.  Triangular-operand flags restrict the k range of a product and mask the other
// triangle at load time.
enum { DLA_LOWER_TILES = 1,   // compute only the 64 x 64 tiles on / below the diagonal of C
       DLA_A_LOWER = 2, DLA_A_UPPER = 4,     // op(A) is lower / upper triangular
       DLA_B_UPPER = 8, DLA_B_LOWER = 16 };  // op(B) is upper / lower triangular
int dla_gemm(gpar_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int flags = 0);
int dla_gemm_batched(gpar_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, long long sA,
                     const double* B, int ldb, long long sB, double beta, double* C, int ldc, long long sC, int batch, int flags = 0);
int dla_potrf(gpar_ctx* ctx, int n, double* A, int lda, int* dinfo);                 // lower; *dinfo = 0 or failing minor (1-based)
int dla_potrf_batched(gpar_ctx* ctx, int n, double* A, int lda, long long sA, int batch, int* dinfo);
int dla_trtri(gpar_ctx* ctx, int n, const double* L, int ldl, double* V, int ldv);   // V = L^-1 (upper triangle zeroed)
int dla_trsm_left(gpar_ctx* ctx, bool trans, int n, int nrhs, const double* L, int ldl, double* B, int ldb);
int dla_trsv(gpar_ctx* ctx, bool trans, int n, const double* L, int ldl, double* x, double scale = 1.0);
int dla_dot(gpar_ctx* ctx, int n, const double* x, const double* y, double* out_dev);
int dla_scal(gpar_ctx* ctx, long long n, double a, double* x);
int dla_symmetrize(gpar_ctx* ctx, int n, double* A, int lda);                         // mirror the lower triangle up
// dense_tail.cu
struct TailBufs {   // M x M scratch of the tail inside ctx->dense
  double *Kj, *Lu, *Bm, *dKu, *V, *Kinv, *R, *Pm, *Tm, *Cm, *cvec, *wvec, *sc;
  int* dinfo;
};
constexpr int GPAR_NTR = 20;      // number of trace scalars produced by the tail (after 8 header scalars)
int tail_layout(gpar_ctx* ctx, bool want_grad, int vfe, TailBufs* b);
int dtc_tail_prepare(gpar_ctx* ctx, int kind, const GpParams& p, int vfe, double jitter, bool want_grad);
// raw_out (nullable): host array of 8 + GPAR_NTR doubles receiving [trB, logdetLambda, c'c, -, ..., traces]
int dtc_tail(gpar_ctx* ctx, int kind, const GpParams& p, int vfe, double jitter, int64_t N,
             const double* G, const double* H, const double* g, const double* h, double yy,
             double* val, double* grad, double* raw_out = nullptr, bool whitened_G = false);
// The tail in whitened coordinates (ill-conditioned cov(u)): statistics of the panels whitened by L_u before the SYRK.
// Cop = L_u^-T Lambda^-1 (dense M x M, the operand of S = A'(Lambda^-1 L_u^-1)), wt = Lambda^-1 A alpha, w = L_u^-T wt — device;
// trQ, WQ, XQ: tr Q, <L_u^-1 L_u^-T, Q>, <L_u^-1 (l dKuu/dl) L_u^-T, Q> with Q = Lambda^-1 - I + wt wt'.
struct WhitenedTail { const double *Cop, *wt, *w; double trQ, WQ, XQ, cc; };
int dtc_tail_whitened(gpar_ctx* ctx, const GpParams& p, int vfe, double jitter, int64_t N,
                      const double* Gw, const double* Hw, const double* g, const double* h, double yy,
                      double* val, double* grad, WhitenedTail* out);
// scaled.cu: a row slice of the scaled objective (the context holds the full (t, y) and rows [lo, lo + N) of X); group.cu
// all-gathers ctx->slice.summary between phase 1 and 2 and all-reduces ctx->slice.G (stats_count doubles) before finish
int scaled_slice_phase1(gpar_ctx* ctx, int k_time, int k_out, const double theta[5], int64_t lo, bool grad = false, bool qu = false);   // qu: theta = the positive parameters of compute_q_u, bare Cuu
int scaled_slice_qu_finish(gpar_ctx* ctx, double* m_e, double* Dinv, double* U_u);
int scaled_slice_qu_sample(gpar_ctx* ctx, uint64_t seed, int32_t S, double* W_out, double* eps_out);
// gradient mode: after the all-reduce every member runs the tail (P, w) and the zero-start tangent responses of its chunks and
// leaves 3 tangent summaries (ctx->slice.summary2, 3 x summary_count doubles) for a second all-gather; phase 4 finishes the
// tangent pass from the gathered summaries and returns the member's five partial sums; finish assembles on one member
int scaled_slice_grad_phase3(gpar_ctx* ctx);
int scaled_slice_grad_phase4(gpar_ctx* ctx, const double* gathered2, int member, double s5[5]);
int scaled_slice_grad_finish(gpar_ctx* ctx, const double s5[5], double* dtc, double* grad);
int scaled_slice_phase2(gpar_ctx* ctx, const double* gathered, int member);
int scaled_slice_finish(gpar_ctx* ctx, double* dtc);
// scaled.cu: conditioning decision and the panel whitening by L_u (see gpar_needs_whitened_panel)
int panel_left_solve(gpar_ctx* ctx, double* panel, int64_t Npad, int Mpad, int M, const double* Lu, const double* src = nullptr);   // src: out of place
bool gpar_needs_whitened_panel(const double minmax[2]);
// Gradient of a value entry point by the 4-point central stencil (h = 1e-2 in the raw log-space parameters; env
// GPAR_FD_STEP): the fallback of the analytic-gradient entry points when cov(u) is too poorly conditioned for the
// collapsed statistic ("Conditioning") — the value path whitens the panel by L_u and keeps 1e-8.
template <class ValFn>
int gpar_fd_gradient(ValFn f, const double* theta, int n, double* grad) {
  double h = 1e-2;      // log-space parameters: the h^4 truncation stays below the round-off noise of the value / h
  if (const char* e = getenv("GPAR_FD_STEP")) { const double v = atof(e); if (v > 0.0) h = v; }
  double th[8];
  for (int i = 0; i < n; i++) th[i] = theta[i];
  for (int i = 0; i < n; i++) {
    double v[4];
    const double off[4] = {-2.0 * h, -h, h, 2.0 * h};
    for (int q = 0; q < 4; q++) { th[i] = theta[i] + off[q]; CHK(f(th, &v[q])); }
    th[i] = theta[i];
    grad[i] = (v[0] - 8.0 * v[1] + 8.0 * v[2] - v[3]) / (12.0 * h);
  }
  return GPAR_OK;
}
int launch_diag_minmax(gpar_ctx* ctx, const double* L, int M, double* out2);
