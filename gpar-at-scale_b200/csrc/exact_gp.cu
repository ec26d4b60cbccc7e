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

// ---- small problems (N <= EXACT_SMALL_MAX): ONE CTA per hyper-parameter candidate does everything -----------------
// The reference's exact-GP sizes are N = 30 (toy), 156 (EEG), ~200: there a log-pdf evaluation is launch latency, and the
// optimiser (optimized.jl:45,164) re-enters it hundreds of times.  The whole evaluation — kernel matrix (packed lower
// triangle in shared memory), right-looking Cholesky, forward substitution per output sequence, log-determinant and
// quadratic form — is one kernel, and a grid of such CTAs evaluates a BATCH of candidates (simplex vertices x restarts)
// on the shared resident data in one launch (gpar_exact_logpdf_batch).
constexpr int EXACT_SMALL_MAX = 200;       // packed N (N + 1) / 2 doubles + vectors within 227 KB of shared memory
constexpr int EXACT_SMALL_RHS = 8;         // output sequences handled by the fused kernels

struct ExactCand { ExactKernel ek; double noise, kss; };

__device__ __forceinline__ double exact_k(const ExactKernel& ek, const double* xa, const double* xb, int D) {
  if (ek.gpar) {
    const double dt = xa[0] - xb[0]; double d2 = 0.0;
    for (int d = 1; d < D; d++) { const double df = xa[d] - xb[d]; d2 = fma(df, df, d2); }
    return ek.time_s * eval_kind(ek.k_time, dt * dt * ek.time_il2) + ek.out_s * eval_kind(ek.k_out, d2 * ek.out_il2);
  }
  double d2 = 0.0;
  for (int d = 0; d < D; d++) { const double df = xa[d] - xb[d]; d2 = fma(df, df, d2); }
  return ek.time_s * eval_kind(ek.k_time, d2 * ek.time_il2);
}
#define PK(i, j) ((i) * ((i) + 1) / 2 + (j))

// K + noise I (packed lower) and its Cholesky factor in place; rd = 1 / diag(L).  Returns 0 or the failing minor.
__device__ int exact_small_factor(double* P, double* rd, const double* __restrict__ X, int n, int D, const ExactCand& cd, int* sbad) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < n; i += nw)
    for (int j = lane; j <= i; j += 32) P[PK(i, j)] = exact_k(cd.ek, X + (int64_t)i * D, X + (int64_t)j * D, D) + (i == j ? cd.noise : 0.0);
  __syncthreads();
  // right-looking with deferred scaling (one barrier per column): trailing(i, c) -= a_ij a_cj / d_j on the unscaled column
  int bad = 0;
  for (int j = 0; j < n; j++) {
    const double d = P[PK(j, j)];
    if (!(d > 0.0)) { bad = j + 1; break; }          // uniform: every thread reads the same pivot
    const double inv = 1.0 / d;
    // threads as a 16 x 16 grid over (row, column) residues of the trailing block; four columns per step so that the
    // shared-memory loads of a step are independent of its stores
    const int ti = tid & 15, tc = tid >> 4;
    for (int i = j + 1 + ti; i < n; i += 16) {
      const double lij = P[PK(i, j)] * inv;
      double* row = P + PK(i, 0);
      int c = j + 1 + tc;
      for (; c + 48 <= i; c += 64) {
        const double l0 = P[PK(c, j)], l1 = P[PK(c + 16, j)], l2 = P[PK(c + 32, j)], l3 = P[PK(c + 48, j)];
        const double p0 = row[c], p1 = row[c + 16], p2 = row[c + 32], p3 = row[c + 48];
        row[c] = fma(-lij, l0, p0); row[c + 16] = fma(-lij, l1, p1); row[c + 32] = fma(-lij, l2, p2); row[c + 48] = fma(-lij, l3, p3);
      }
      for (; c <= i; c += 16) row[c] = fma(-lij, P[PK(c, j)], row[c]);
    }
    __syncthreads();
  }
  (void)sbad;
  if (bad) return bad;
  for (int i = warp; i < n; i += nw)
    for (int c = lane; c < i; c += 32) P[PK(i, c)] *= rsqrt(P[PK(c, c)]);
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) { const double l = sqrt(P[PK(i, i)]); P[PK(i, i)] = l; rd[i] = 1.0 / l; }
  __syncthreads();
  return 0;
}
// w <- L^-1 w by ONE warp (w in shared memory)
__device__ __forceinline__ void exact_small_forward(const double* P, const double* rd, double* w, int n, int lane) {
  for (int j = 0; j < n; j++) {
    const double wj = w[j] * rd[j];
    __syncwarp();
    if (lane == 0) w[j] = wj;
    for (int i = j + 1 + lane; i < n; i += 32) w[i] = fma(-P[PK(i, j)], wj, w[i]);
    __syncwarp();
  }
}

// grid = candidates; lml[cand * batch + b]; info[cand] = 0 or the failing minor
__global__ void __launch_bounds__(256)
exact_logpdf_small_kernel(const double* __restrict__ X, const double* __restrict__ Y, int n, int D, int batch, const ExactCand* __restrict__ cands,
                          double* __restrict__ lml, int* __restrict__ info) {
  extern __shared__ double esm[];
  double* P = esm; double* rd = P + n * (n + 1) / 2; double* W = rd + n;       // W: batch x n
  __shared__ int sbad;
  __shared__ double part[8];
  const ExactCand cd = cands[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < batch * n; e += blockDim.x) W[e] = Y[e];
  const int bad = exact_small_factor(P, rd, X, n, D, cd, &sbad);
  if (tid == 0) info[blockIdx.x] = bad;
  if (bad) { for (int b = tid; b < batch; b += blockDim.x) lml[blockIdx.x * batch + b] = nan(""); return; }
  if (warp < batch) exact_small_forward(P, rd, W + warp * n, n, lane);          // batch <= 8 = number of warps
  // log det = 2 sum log L_ii (block reduction), quadratic forms per warp
  double ld = 0.0;
  for (int i = tid; i < n; i += blockDim.x) ld -= log(rd[i]);
  for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
  if (lane == 0) part[warp] = ld;
  __syncthreads();
  if (warp < batch) {
    double q = 0.0;
    for (int i = lane; i < n; i += 32) q = fma(W[warp * n + i], W[warp * n + i], q);
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) {
      double t = 0.0;
      for (int k = 0; k < 8; k++) t += part[k];
      lml[blockIdx.x * batch + warp] = -0.5 * ((double)n * 1.8378770664093454835606594728112 + 2.0 * t + q);
    }
  }
}

// posterior marginals: every CTA factors K again (cheap) and takes a chunk of the test points, one warp per point
__global__ void __launch_bounds__(256)
exact_posterior_small_kernel(const double* __restrict__ X, const double* __restrict__ Y, int n, int D, int batch, ExactCand cd,
                             const double* __restrict__ Xs, int Ns, double* __restrict__ mean, double* __restrict__ var, int* __restrict__ info) {
  extern __shared__ double esm[];
  double* P = esm; double* rd = P + n * (n + 1) / 2; double* W = rd + n; double* V = W + batch * n;     // V: 8 warps x n
  __shared__ int sbad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < batch * n; e += blockDim.x) W[e] = Y[e];
  const int bad = exact_small_factor(P, rd, X, n, D, cd, &sbad);
  if (tid == 0 && blockIdx.x == 0) info[0] = bad;
  if (bad) return;
  if (warp < batch) exact_small_forward(P, rd, W + warp * n, n, lane);
  __syncthreads();
  double* v = V + warp * n;
  for (int s = blockIdx.x * 8 + warp; s < Ns; s += gridDim.x * 8) {
    for (int i = lane; i < n; i += 32) v[i] = exact_k(cd.ek, X + (int64_t)i * D, Xs + (int64_t)s * D, D);
    __syncwarp();
    exact_small_forward(P, rd, v, n, lane);
    double q = 0.0;
    for (int i = lane; i < n; i += 32) q = fma(v[i], v[i], q);
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) var[s] = cd.kss - q + 1e-18;        // + Stheno's default 1e-18 observation noise of post(x*)
    for (int b = 0; b < batch; b++) {
      double m = 0.0;
      for (int i = lane; i < n; i += 32) m = fma(v[i], W[b * n + i], m);
      for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
      if (lane == 0) mean[(int64_t)b * Ns + s] = m;
    }
    __syncwarp();
  }
}
#undef PK

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

// `ncand` candidates (theta + ntheta * c) on the resident (X, y): lml[c * batch + b], codes[c] = 0 or the failing minor.
int exact_logpdf_small(gpar_ctx* ctx, int k_time, int k_out, const double* thetas, int ntheta, int ncand, double* lml, int32_t* codes) {
  const int n = (int)ctx->N, batch = ctx->ybatch, D = ctx->D;
  std::vector<ExactCand> hc(ncand);
  for (int c = 0; c < ncand; c++) CHK(setup_kernel(ctx, k_time, k_out, thetas + (size_t)ntheta * c, ntheta, &hc[c].ek, &hc[c].noise, &hc[c].kss));
  CU(ctx->dense.reserve((size_t)ncand * sizeof(ExactCand) + (size_t)ncand * batch * sizeof(double) + (size_t)ncand * sizeof(int) + 64));
  ExactCand* dc = ctx->dense.as<ExactCand>();
  double* dl = reinterpret_cast<double*>(dc + ncand); int* di = reinterpret_cast<int*>(dl + (size_t)ncand * batch);
  CU(cudaMemcpyAsync(dc, hc.data(), (size_t)ncand * sizeof(ExactCand), cudaMemcpyHostToDevice, ctx->stream));
  CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
  const size_t sm = ((size_t)n * (n + 1) / 2 + n + (size_t)batch * n) * sizeof(double);
  CU(cudaFuncSetAttribute(exact_logpdf_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  LAUNCH(ctx, exact_logpdf_small_kernel, ncand, 256, sm, ctx->X.as<double>(), ctx->y.as<double>(), n, D, batch, dc, dl, di);
  timer.stop();
  std::vector<int> hi(ncand);
  CU(cudaMemcpyAsync(lml, dl, (size_t)ncand * batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hi.data(), di, (size_t)ncand * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));        // hc / hi go out of scope below
  for (int c = 0; c < ncand; c++) codes[c] = hi[c];
  return GPAR_OK;
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
  if (ctx->N <= EXACT_SMALL_MAX && ctx->ybatch <= EXACT_SMALL_RHS) {      // one fused kernel (see exact_logpdf_small_kernel)
    int32_t code = 0;
    CHK(exact_logpdf_small(ctx, k_time, k_out, theta, ntheta, 1, lml, &code));
    if (code != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(K + sigma^2 I) failed: leading minor %d is not positive definite", (int)code);
    return GPAR_OK;
  }
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

// `ncand` hyper-parameter candidates (columns of thetas, ntheta x ncand) on the resident data in ONE launch.
int gpar_exact_logpdf_batch(gpar_ctx* ctx, int k_time, int k_out, const double* thetas, int32_t ntheta, int32_t ncand, double* lml, int32_t* codes) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!thetas || !lml || ncand < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "exact_logpdf_batch: NULL argument or ncand < 1");
  CHK(check_exact(ctx, "exact_logpdf_batch"));
  CU(cudaSetDevice(ctx->device));
  std::vector<int32_t> cd(ncand, 0);
  if (ctx->N <= EXACT_SMALL_MAX && ctx->ybatch <= EXACT_SMALL_RHS) {
    CHK(exact_logpdf_small(ctx, k_time, k_out, thetas, ntheta, ncand, lml, cd.data()));
  } else {        // large N: an evaluation fills the device by itself — one after the other
    double ms = 0.0; int64_t nl = 0;
    for (int c = 0; c < ncand; c++) {
      const int rc = gpar_exact_logpdf(ctx, k_time, k_out, thetas + (size_t)ntheta * c, ntheta, lml + (size_t)c * ctx->ybatch);
      if (rc == GPAR_ERR_NOT_POSDEF) { cd[c] = 1; for (int b = 0; b < ctx->ybatch; b++) lml[(size_t)c * ctx->ybatch + b] = NAN; }
      else if (rc != GPAR_OK) return rc;
      ms += ctx->last_ms; nl += ctx->last_launches;
    }
    ctx->last_ms = ms; ctx->last_launches = nl;
  }
  bool any = false;
  for (int c = 0; c < ncand; c++) { if (codes) codes[c] = cd[c]; any = any || cd[c] != 0; }
  if (any && !codes) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "exact_logpdf_batch: a candidate's cholesky(K + sigma^2 I) failed (pass `codes` to get them individually)");
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
  if (n <= EXACT_SMALL_MAX && batch <= EXACT_SMALL_RHS && ((size_t)n * (n + 1) / 2 + n + (size_t)(batch + 8) * n) * sizeof(double) <= 220 * 1024) {
    // fused small path: every CTA factors K (packed, shared memory) and takes a chunk of the test points
    CU(ctx->kal_b.reserve(((size_t)Ns * D + (size_t)Ns * batch + Ns) * sizeof(double) + 64));
    double* dXs = ctx->kal_b.as<double>(); double* dmean = dXs + (size_t)Ns * D; double* dvar = dmean + (size_t)Ns * batch;
    CU(ctx->info.reserve(4 * sizeof(int)));
    CU(cudaMemcpyAsync(dXs, Xs, (size_t)Ns * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
    ExactCand cd{ek, noise, kss};
    const size_t sm = ((size_t)n * (n + 1) / 2 + n + (size_t)(batch + 8) * n) * sizeof(double);
    CU(cudaFuncSetAttribute(exact_posterior_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((Ns + 15) / 16, 2 * ctx->num_sms));
    LAUNCH(ctx, exact_posterior_small_kernel, grid, 256, sm, ctx->X.as<double>(), ctx->y.as<double>(), n, D, batch, cd, dXs, (int)Ns, dmean, dvar, ctx->info.as<int>());
    timer.stop();
    int hinfo = 0;
    CU(cudaMemcpyAsync(mean, dmean, (size_t)Ns * batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(var, dvar, (size_t)Ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&hinfo, ctx->info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (hinfo != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(K + sigma^2 I) failed: leading minor %d is not positive definite", hinfo);
    return GPAR_OK;
  }
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
