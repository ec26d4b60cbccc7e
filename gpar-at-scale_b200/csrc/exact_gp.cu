// exact_gp.cu — exact (dense) GP / GPAR log-marginal likelihood and posterior marginals.
//
// Replaces Stheno's `logpdf(f(x, sigma^2), y)` (src/gp/optimized.jl:34,152) and
// `gp | (gp(x, sigma^2) <- y)` + `marginals` / `mean` (optimized.jl:94,236; eeg.jl:185-208;
// GPAR_examples/toy_example.jl:118-134) for the small-N path (N = 30 ... a few hundred).
// ntheta = 3: kernel(k_time; l, s = var^2) on all input features (optimized.jl:28-36);
// ntheta = 5: GPAR kernel time_var^2 k_t(|dx_1|/time_l) + out_var^2 k_o(||dx_2:D||/out_l)
// (create_gpar_kernel, optimized.jl:132-144).  One factorisation serves every resident output
// sequence (batch of right-hand sides).  Latency-bound: kernel-matrix evaluation and reductions are
// kernels of this file; the N^3/3 factorisation, the blocked triangular solves and the products are the hand-written
// routines of dense_la.cu (no library call).
#include "common.cuh"
#include <algorithm>

namespace {

__device__ __forceinline__ double eval_kind(int kind, double d2) {
  double dummy;
  switch (kind) {
    case GPAR_EQ: return base_kernel_dev<GPAR_EQ, false>(d2, dummy);
    case GPAR_MATERN12: return base_kernel_dev<GPAR_MATERN12, false>(d2, dummy);
    case GPAR_MATERN32: return base_kernel_dev<GPAR_MATERN32, false>(d2, dummy);
    default: return base_kernel_dev<GPAR_MATERN52, false>(d2, dummy);
  }
}

struct ExactKernel { int k_time, k_out, gpar; double time_il2, time_s, out_il2, out_s; };

// K[a + b*lda] = k(A_a, B_b) (+ noise on the diagonal when A == B and a == b)
__global__ void exact_kernel_matrix(const double* __restrict__ A, int na, const double* __restrict__ B, int nb, int D,
                                    ExactKernel ek, double diag_add, double* __restrict__ K, int lda) {
  int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= na || b >= nb) return;
  const double* xa = A + (int64_t)a * D; const double* xb = B + (int64_t)b * D;
  double v;
  if (ek.gpar) {
    double dt = xa[0] - xb[0], d2 = 0.0;
    for (int d = 1; d < D; d++) { double df = xa[d] - xb[d]; d2 = fma(df, df, d2); }
    v = ek.time_s * eval_kind(ek.k_time, dt * dt * ek.time_il2) + ek.out_s * eval_kind(ek.k_out, d2 * ek.out_il2);
  } else {
    double d2 = 0.0;
    for (int d = 0; d < D; d++) { double df = xa[d] - xb[d]; d2 = fma(df, df, d2); }
    v = ek.time_s * eval_kind(ek.k_time, d2 * ek.time_il2);
  }
  if (A == B && a == b) v += diag_add;
  K[(int64_t)a + (int64_t)b * lda] = v;
}

// out[b] = sum_i W[i + b*n]^2   (one block per column)
__global__ void colsumsq_kernel(const double* __restrict__ W, int n, double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { double v = W[(int64_t)i + (int64_t)blockIdx.x * n]; acc = fma(v, v, acc); }
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = r;
}
__global__ void logdet_chol_kernel(const double* L, int n, double* out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += log(L[(int64_t)i * n + i]);
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = 2.0 * r;
}

int setup_kernel(gpar_ctx* ctx, int k_time, int k_out, const double* theta, int ntheta, ExactKernel* ek, double* noise, double* kss) {
  if (ntheta == 3) {
    GpParams p = unpack_gp3(theta);
    *ek = ExactKernel{k_time, k_time, 0, 1.0 / (p.l * p.l), p.s, 0.0, 0.0};
    *noise = p.noise; *kss = p.s;
  } else if (ntheta == 5) {
    if (ctx->D < 2) return gpar_fail(ctx, GPAR_ERR_INVALID, "exact GPAR kernel needs D >= 2 input features (time + previous outputs), got D=%d", ctx->D);  // util.jl:112-117
    double pv[5];
    for (int i = 0; i < 5; i++) pv[i] = exp(theta[i]) + 1e-3;
    *ek = ExactKernel{k_time, k_out, 1, 1.0 / (pv[0] * pv[0]), pv[1] * pv[1], 1.0 / (pv[2] * pv[2]), pv[3] * pv[3]};
    *noise = pv[4] * pv[4]; *kss = pv[1] * pv[1] + pv[3] * pv[3];
  } else {
    return gpar_fail(ctx, GPAR_ERR_INVALID, "ntheta must be 3 (GP) or 5 (GPAR), got %d", ntheta);
  }
  if (k_time < GPAR_EQ || k_time > GPAR_MATERN52 || k_out < GPAR_EQ || k_out > GPAR_MATERN52)
    return gpar_fail(ctx, GPAR_ERR_INVALID, "unknown kernel code");
  return GPAR_OK;
}

// factorises K + sigma^2 I into `L` (n x n) and solves W = L^{-1} Y (n x batch)
int factor_and_whiten(gpar_ctx* ctx, const ExactKernel& ek, double noise, double* L, double* W, int* dinfo) {
  const int n = (int)ctx->N, batch = ctx->ybatch;
  const double* X = ctx->X.as<double>();
  dim3 grid((n + 127) / 128, n);
  LAUNCH(ctx, exact_kernel_matrix, grid, 128, 0, X, n, X, n, ctx->D, ek, noise, L, n);
  CHK(dla_potrf(ctx, n, L, n, dinfo));
  CU(cudaMemcpyAsync(W, ctx->y.p, (size_t)n * batch * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return dla_trsm_left(ctx, false, n, batch, L, n, W, n);
}

int check_exact(gpar_ctx* ctx, const char* who) {
  if (ctx->N < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: inputs not set", who);
  if (ctx->Ny != ctx->N || ctx->ybatch < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: outputs length %lld != N %lld", who, (long long)ctx->Ny, (long long)ctx->N);
  if (ctx->N > 32768) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: exact path is dense O(N^2) memory; N=%lld is too large (use the pseudo-point or state-space path)", who, (long long)ctx->N);
  return GPAR_OK;
}

}  // namespace

extern "C" {

int gpar_exact_logpdf(gpar_ctx* ctx, int k_time, int k_out, const double* theta, int32_t ntheta, double* lml) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!theta || !lml) return gpar_fail(ctx, GPAR_ERR_INVALID, "exact_logpdf: NULL argument");
  CHK(check_exact(ctx, "exact_logpdf"));
  ExactKernel ek; double noise, kss;
  CHK(setup_kernel(ctx, k_time, k_out, theta, ntheta, &ek, &noise, &kss));
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
  const int n = (int)ctx->N, batch = ctx->ybatch;
  CU(ctx->dense.reserve(((size_t)n * n + (size_t)n * batch + batch + 8) * sizeof(double)));
  CU(ctx->info.reserve(4 * sizeof(int)));
  double* L = ctx->dense.as<double>(); double* W = L + (size_t)n * n; double* q = W + (size_t)n * batch; double* ld = q + batch;
  CHK(factor_and_whiten(ctx, ek, noise, L, W, ctx->info.as<int>()));
  LAUNCH(ctx, colsumsq_kernel, batch, 256, 0, W, n, q);
  LAUNCH(ctx, logdet_chol_kernel, 1, 256, 0, L, n, ld);
  timer.stop();
  std::vector<double> h(batch + 1); int hinfo = 0;
  CU(cudaMemcpyAsync(h.data(), q, (batch + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(&hinfo, ctx->info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (hinfo != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(K + sigma^2 I) failed: leading minor %d is not positive definite", hinfo);
  for (int b = 0; b < batch; b++) lml[b] = -0.5 * ((double)n * 1.8378770664093454835606594728112 + h[batch] + h[b]);
  return GPAR_OK;
}

int gpar_exact_posterior(gpar_ctx* ctx, int k_time, int k_out, const double* theta, int32_t ntheta,
                         const double* Xs, int64_t Ns, double* mean, double* var) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!theta || !Xs || !mean || !var || Ns < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "exact_posterior: NULL argument or Ns < 1");
  CHK(check_exact(ctx, "exact_posterior"));
  ExactKernel ek; double noise, kss;
  CHK(setup_kernel(ctx, k_time, k_out, theta, ntheta, &ek, &noise, &kss));
  CU(cudaSetDevice(ctx->device));
  const int n = (int)ctx->N, batch = ctx->ybatch, D = ctx->D;
  CU(ctx->kal_b.reserve(((size_t)Ns * D + (size_t)n * Ns + (size_t)Ns * batch + Ns) * sizeof(double)));
  double* dXs = ctx->kal_b.as<double>(); double* V = dXs + (size_t)Ns * D; double* dmean = V + (size_t)n * Ns; double* dq = dmean + (size_t)Ns * batch;
  CU(cudaMemcpyAsync(dXs, Xs, (size_t)Ns * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
  CU(ctx->dense.reserve(((size_t)n * n + (size_t)n * batch + 8) * sizeof(double)));
  CU(ctx->info.reserve(4 * sizeof(int)));
  double* L = ctx->dense.as<double>(); double* W = L + (size_t)n * n;
  CHK(factor_and_whiten(ctx, ek, noise, L, W, ctx->info.as<int>()));
  // V = L^{-1} K_{f*}  (n x Ns);  mean = V^T W;  var = k** - colsum(V^2) + 1e-18
  dim3 grid((n + 127) / 128, (unsigned)Ns);
  LAUNCH(ctx, exact_kernel_matrix, grid, 128, 0, ctx->X.as<double>(), n, dXs, (int)Ns, D, ek, 0.0, V, n);
  CHK(dla_trsm_left(ctx, false, n, (int)Ns, L, n, V, n));
  CHK(dla_gemm(ctx, true, false, (int)Ns, batch, n, 1.0, V, n, W, n, 0.0, dmean, (int)Ns));
  LAUNCH(ctx, colsumsq_kernel, (int)Ns, 128, 0, V, n, dq);
  timer.stop();
  std::vector<double> hq(Ns); int hinfo = 0;
  CU(cudaMemcpyAsync(mean, dmean, (size_t)Ns * batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hq.data(), dq, (size_t)Ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(&hinfo, ctx->info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (hinfo != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(K + sigma^2 I) failed: leading minor %d is not positive definite", hinfo);
  for (int64_t i = 0; i < Ns; i++) var[i] = kss - hq[i] + 1e-18;   // + Stheno's default 1e-18 observation noise of post(x*)
  return GPAR_OK;
}

}  // extern "C"
