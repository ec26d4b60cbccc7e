// smooth_shared.cu — RTS smoothing of MANY sequences that share one LGSSM (same times, same parameters).
//
// This is the reference's Monte-Carlo loop (src/gp/gpar_scaled_inference.jl:110-123: 100 x
// `smooth(time_lgssm_star, y_star - fx)` with ONE model) and `smooth` on several outputs of one grid.
// Gains, smoother gains and all covariances do not depend on the data, so they are computed ONCE from a
// single-sequence run of the general scan (kalman.cu), which leaves two per-step tables:
//   forward  (filter)   row k: Phi_k = (I - K_k H) A_k, K_k, ...        m_k   = Phi_k m_{k-1} + K_k y_k
//   backward (smoother) row k: B_k = I - G_k A_{k+1}, G_k                m^s_k = G_k m^s_{k+1} + B_k m_k
// Every sequence then needs only these two affine recursions (12 + 18 FMA per step instead of a full
// covariance smoother, ~700): chunked along time exactly like the whitening of scaled.cu — zero-start chunk
// responses, ordered chunk products of the shared matrices, a carry scan, final pass.  Data are TIME-MAJOR
// ([n][S], S padded to 128): a warp is 32 sequences at one time step, so every global access is a coalesced
// 256-byte run and the table row of a step is a shared-memory broadcast.  Bound: HBM (8 + 24 B/step forward,
// 24 + 8 B/step backward).
#include "lgssm_math.cuh"
#include <algorithm>
#include <cstdlib>

struct SharedSmoothExtra { const double* var1; const double* sum_logS; const double* a2part; int nch; };   // device pointers into ctx->shbuf

namespace {

constexpr int SH_SUB = 32;       // steps per staged table window
constexpr int SH_PF = 8;         // steps whose global operands are fetched ahead (SH_SUB is a multiple)
__device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

__device__ __forceinline__ void sh_cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <int ROW>
__device__ __forceinline__ void sh_stage(double* buf, const double* __restrict__ table, int64_t k_lo, int nrows) {
  const int total = nrows * ROW;
  const double* src = table + k_lo * ROW;
  for (int e = threadIdx.x; e < total; e += blockDim.x) sh_cp_async8(buf + e, src + e);
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Ordered product of the chunk's matrices, one warp per chunk (lane = contiguous run of steps, then a shuffle tree).
// forward: Psi_c = M_{k1-1} ... M_{k0} (later steps on the left);  reverse: Psi'_c = M_{k0} ... M_{k1-1}.
template <int D, bool REVERSE>
__global__ void __launch_bounds__(32)
sh_chunk_product_kernel(const double* __restrict__ table, int row, int off, int64_t N, int LC, double* __restrict__ psi) {
  const int c = blockIdx.x, lane = threadIdx.x;
  const int64_t k0 = (int64_t)c * LC, k1 = imin64(N, k0 + LC);
  const int per = LC / 32;
  double Pm[D * D];
#pragma unroll
  for (int i = 0; i < D * D; i++) Pm[i] = (i / D == i % D) ? 1.0 : 0.0;
  for (int q = 0; q < per; q++) {
    const int64_t k = k0 + (int64_t)lane * per + q;
    if (k < k1) {
      double F[D * D], R[D * D];
#pragma unroll
      for (int i = 0; i < D * D; i++) F[i] = __ldg(table + k * row + off + i);
      if (REVERSE) matmul<D>(Pm, F, R); else matmul<D>(F, Pm, R);
#pragma unroll
      for (int i = 0; i < D * D; i++) Pm[i] = R[i];
    }
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {       // after round d, lane l (l % 2d == 0) holds the product of lanes l .. l+2d-1
    double O[D * D], R[D * D];
#pragma unroll
    for (int i = 0; i < D * D; i++) O[i] = __shfl_down_sync(0xffffffffu, Pm[i], d);
    if (REVERSE) matmul<D>(Pm, O, R); else matmul<D>(O, Pm, R);     // the partner covers LATER steps
#pragma unroll
    for (int i = 0; i < D * D; i++) Pm[i] = R[i];
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < D * D; i++) psi[(int64_t)c * D * D + i] = Pm[i];
  }
}

// carry over the chunks, in place: resp[c] (zero-start response of chunk c) -> state entering chunk c.
// forward: in[0] = 0, in[c+1] = Psi_c in[c] + resp[c];  reverse: in[nch-1] = 0, in[c-1] = Psi'_c in[c] + resp[c].
// A warp per sequence: every lane composes the affine maps of its K consecutive chunks (in
// processing order), the lane maps go through a warp-shuffle scan, and the lane walks its chunks again from its exclusive
// prefix: 2K + 5 dependent steps instead of nch (105 chunks at 1024 x 10k: 27 us -> a few us per carry).
template <int D, bool REVERSE>
__global__ void __launch_bounds__(256)
sh_carry_warp_kernel(const double* __restrict__ psi, double* __restrict__ resp, int nch, int Sp) {
  const int lane = threadIdx.x & 31, s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= Sp) return;
  const int K = (nch + 31) / 32;
  const int q0 = lane * K, q1 = min(q0 + K, nch);
  double P[D * D], r[D];
#pragma unroll
  for (int i = 0; i < D * D; i++) P[i] = (i / D == i % D) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) r[i] = 0.0;
  for (int q = q0; q < q1; q++) {
    const int c = REVERSE ? nch - 1 - q : q;
    double F[D * D], R[D * D], nr[D];
#pragma unroll
    for (int i = 0; i < D * D; i++) F[i] = __ldg(psi + (int64_t)c * D * D + i);
#pragma unroll
    for (int i = 0; i < D; i++) { double a = resp[((int64_t)c * D + i) * Sp + s];
#pragma unroll
      for (int j = 0; j < D; j++) a = fma(F[i * D + j], r[j], a);
      nr[i] = a; }
    matmul<D>(F, P, R);
#pragma unroll
    for (int i = 0; i < D * D; i++) P[i] = R[i];
#pragma unroll
    for (int i = 0; i < D; i++) r[i] = nr[i];
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double Po[D * D], ro[D];
#pragma unroll
    for (int i = 0; i < D * D; i++) Po[i] = __shfl_up_sync(0xffffffffu, P[i], d);
#pragma unroll
    for (int i = 0; i < D; i++) ro[i] = __shfl_up_sync(0xffffffffu, r[i], d);
    if (lane >= d) {      // (earlier o this):  P <- P Po,  r <- P ro + r
      double R[D * D], u2[D];
      matmul<D>(P, Po, R);
      matvec<D>(P, ro, u2);
#pragma unroll
      for (int i = 0; i < D * D; i++) P[i] = R[i];
#pragma unroll
      for (int i = 0; i < D; i++) r[i] += u2[i];
    }
  }
  double x[D];
#pragma unroll
  for (int i = 0; i < D; i++) { const double v = __shfl_up_sync(0xffffffffu, r[i], 1); x[i] = lane == 0 ? 0.0 : v; }
  for (int q = q0; q < q1; q++) {
    const int c = REVERSE ? nch - 1 - q : q;
    double nx[D];
#pragma unroll
    for (int i = 0; i < D; i++) { double* p = resp + ((int64_t)c * D + i) * Sp + s; double a = *p; *p = x[i];
#pragma unroll
      for (int j = 0; j < D; j++) a = fma(__ldg(psi + (int64_t)c * D * D + i * D + j), x[j], a);
      nx[i] = a; }
#pragma unroll
    for (int i = 0; i < D; i++) x[i] = nx[i];
  }
}

// forward pass: m_k = Phi_k m_{k-1} + K_k y_k.  FINAL: from the scanned start states, filtered means stored; else zero-start response.
template <int D, bool FINAL>
__global__ void __launch_bounds__(128)
sh_forward_kernel(int64_t N, int LC, const double* __restrict__ table, const double* __restrict__ yt, int Sp,
                  double* __restrict__ state /* [nch][D][Sp]: start states (FINAL) or responses */, double* __restrict__ mst /* [N][D][Sp] */,
                  double* __restrict__ a2part /* nullable [nch][Sp]: sum alpha^2 of the chunk (FINAL) */,
                  double* __restrict__ alpha_t /* nullable [N][Sp]: whitened innovations (FINAL) */) {
  constexpr int TS = D * D + 2 * D + 1;
  __shared__ double tbl[2][SH_SUB * TS];
  const int s = blockIdx.x * 128 + threadIdx.x, c = blockIdx.y;
  const int64_t k0 = (int64_t)c * LC, k1 = imin64(N, k0 + LC);
  double x[D];
#pragma unroll
  for (int i = 0; i < D; i++) x[i] = FINAL ? state[((int64_t)c * D + i) * Sp + s] : 0.0;
  sh_stage<TS>(tbl[0], table, k0, (int)imin64(SH_SUB, k1 - k0));
  int cur = 0;
  double a2 = 0.0;
  double ycur[SH_PF], ynxt[SH_PF];
#pragma unroll
  for (int u = 0; u < SH_PF; u++) ycur[u] = k0 + u < k1 ? yt[(k0 + u) * Sp + s] : 0.0;
  for (int64_t kw = k0; kw < k1; kw += SH_SUB, cur ^= 1) {
    const int nw = (int)imin64(SH_SUB, k1 - kw);
    if (kw + SH_SUB < k1) { sh_stage<TS>(tbl[cur ^ 1], table, kw + SH_SUB, (int)imin64(SH_SUB, k1 - kw - SH_SUB)); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int q0 = 0; q0 < nw; q0 += SH_PF) {
      // every thread walks its own sequence, so a step's load sits on the critical path of its recurrence: the
      // values of the next SH_PF steps are fetched while this group is computed (1.9 TB/s without, latency bound)
#pragma unroll
      for (int u = 0; u < SH_PF; u++) { const int64_t k = kw + q0 + SH_PF + u; ynxt[u] = k < k1 ? yt[k * Sp + s] : 0.0; }
#pragma unroll
      for (int u = 0; u < SH_PF; u++) {
      const int q = q0 + u;
      if (q >= nw) break;
      const double* r = tbl[cur] + q * TS;
      const double yv = ycur[u];
      double nx[D];
      if (FINAL && a2part) {       // alpha_k = (y_k - HA m_{k-1}) / sqrt(S_k)
        double pred = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) pred = fma(r[D * D + D + j], x[j], pred);
        const double a = (yv - pred) * r[D * D + 2 * D];
        a2 = fma(a, a, a2);
        if (alpha_t) alpha_t[(kw + q) * Sp + s] = a;
      }
#pragma unroll
      for (int i = 0; i < D; i++) { double v = r[D * D + i] * yv;
#pragma unroll
        for (int j = 0; j < D; j++) v = fma(r[i * D + j], x[j], v);
        nx[i] = v; }
#pragma unroll
      for (int i = 0; i < D; i++) { x[i] = nx[i]; if (FINAL && mst) mst[((kw + q) * D + i) * Sp + s] = nx[i]; }
      }
#pragma unroll
      for (int u = 0; u < SH_PF; u++) ycur[u] = ynxt[u];
    }
    __syncthreads();
  }
  if (!FINAL) {
#pragma unroll
    for (int i = 0; i < D; i++) state[((int64_t)c * D + i) * Sp + s] = x[i];
  } else if (a2part) a2part[(int64_t)c * Sp + s] = a2;
}

// backward pass: m^s_k = G_k m^s_{k+1} + B_k m_k, k descending.  FINAL: emits m^s_k[1] (the mean of f(t_k)).
template <int D, bool FINAL>
__global__ void __launch_bounds__(128)
sh_backward_kernel(int64_t N, int LC, const double* __restrict__ table2, const double* __restrict__ mst, int Sp,
                   double* __restrict__ state, double* __restrict__ mean_t /* [N][Sp] */) {
  constexpr int TS = 2 * D * D;
  __shared__ double tbl[2][SH_SUB * TS];
  const int s = blockIdx.x * 128 + threadIdx.x, c = blockIdx.y;
  const int64_t k0 = (int64_t)c * LC, k1 = imin64(N, k0 + LC);
  double z[D];
#pragma unroll
  for (int i = 0; i < D; i++) z[i] = FINAL ? state[((int64_t)c * D + i) * Sp + s] : 0.0;
  // windows are walked from the chunk's end; window w covers [hi - nw, hi)
  int64_t hi = k1;
  int nw = (int)imin64(SH_SUB, hi - k0);
  sh_stage<TS>(tbl[0], table2, hi - nw, nw);
  int cur = 0;
  double mcur[SH_PF][D], mnxt[SH_PF][D];
#pragma unroll
  for (int u = 0; u < SH_PF; u++) {
    const int64_t kk = k1 - 1 - u;
#pragma unroll
    for (int i = 0; i < D; i++) mcur[u][i] = kk >= k0 ? mst[(kk * D + i) * Sp + s] : 0.0;
  }
  while (hi > k0) {
    const int64_t lo = hi - nw;
    const int nnext = (int)imin64(SH_SUB, lo - k0);
    if (nnext > 0) { sh_stage<TS>(tbl[cur ^ 1], table2, lo - nnext, nnext); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int q0 = nw - 1; q0 >= 0; q0 -= SH_PF) {
      const int cnt = q0 + 1 < SH_PF ? q0 + 1 : SH_PF;      // steps this group consumes (a chunk's first window may be ragged)
#pragma unroll
      for (int u = 0; u < SH_PF; u++) {           // the filtered means of the next (earlier) SH_PF steps
        const int64_t kk = lo + q0 - cnt - u;
#pragma unroll
        for (int i = 0; i < D; i++) mnxt[u][i] = kk >= k0 ? mst[(kk * D + i) * Sp + s] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < SH_PF; u++) {
      const int q = q0 - u;
      if (q < 0) break;
      const double* r = tbl[cur] + q * TS;
      const int64_t k = lo + q;
      double m[D], nz[D];
#pragma unroll
      for (int i = 0; i < D; i++) m[i] = mcur[u][i];
#pragma unroll
      for (int i = 0; i < D; i++) { double v = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) { v = fma(r[i * D + j], m[j], v); v = fma(r[D * D + i * D + j], z[j], v); }
        nz[i] = v; }
#pragma unroll
      for (int i = 0; i < D; i++) z[i] = nz[i];
      if (FINAL) mean_t[k * Sp + s] = z[0];
      }
#pragma unroll
      for (int u = 0; u < SH_PF; u++)
#pragma unroll
        for (int i = 0; i < D; i++) mcur[u][i] = mnxt[u][i];
    }
    __syncthreads();
    hi = lo; nw = nnext; cur ^= 1;
  }
  if (!FINAL) {
#pragma unroll
    for (int i = 0; i < D; i++) state[((int64_t)c * D + i) * Sp + s] = z[i];
  }
}

template <int D>
int smooth_shared_d(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, const double* t, const double* y1,
                    const double* rvec, const double* yt, int Sp, double* mean_t, SharedSmoothExtra* ex) {
  constexpr int TS = D * D + 2 * D + 1, TS2 = 2 * D * D;
  // chunk length: enough threads to fill the machine, multiple of 32
  int LC = (int)std::min<int64_t>(1024, std::max<int64_t>(64, ((int64_t)Sp * N / 150000 + 31) / 32 * 32));
  if (const char* e = getenv("GPAR_SH_LC")) { int v = atoi(e); if (v >= 32 && v <= 4096 && v % 32 == 0) LC = v; }      // tuning knob
  const int nch = (int)((N + LC - 1) / LC);
  const size_t state_doubles = (size_t)nch * D * Sp;
  // tables | psi | states | single-sequence scratch (alpha-free: lml, mean, var) | filtered means of all sequences
  CU(ctx->shbuf.reserve(((size_t)N * (TS + TS2) + (size_t)nch * D * D + state_doubles + 2 * (size_t)N + 16 + (size_t)N * D * Sp
                         + (ex ? (size_t)nch * Sp : 0)) * sizeof(double)));
  double* table = ctx->shbuf.as<double>(); double* table2 = table + (size_t)N * TS; double* psi = table2 + (size_t)N * TS2;
  double* state = psi + (size_t)nch * D * D; double* m1 = state + state_doubles; double* v1 = m1 + N; double* lml1 = v1 + N;
  double* mst = lml1 + 16;
  double* a2part = ex ? mst + (size_t)N * D * Sp : nullptr;
  // the data-independent part, once: forward table from a filter pass, backward table from a smoother pass (one sequence)
  // (one smoother run leaves both tables and the sums: lml1[2] = sum log S)
  CHK(lgssm_run(ctx, kind, &l, &s, &noise, 1, 1, N, t, y1, rvec, nullptr, lml1, m1, v1, table2, lml1 + 2, table));
  dim3 grid(Sp / 128, nch);
  LAUNCH(ctx, (sh_forward_kernel<D, false>), grid, 128, 0, N, LC, table, yt, Sp, state, mst, (double*)nullptr, (double*)nullptr);
  LAUNCH(ctx, (sh_chunk_product_kernel<D, false>), nch, 32, 0, table, TS, 0, N, LC, psi);
  LAUNCH(ctx, (sh_carry_warp_kernel<D, false>), (Sp + 7) / 8, 256, 0, psi, state, nch, Sp);
  LAUNCH(ctx, (sh_forward_kernel<D, true>), grid, 128, 0, N, LC, table, yt, Sp, state, mst, a2part, (double*)nullptr);
  LAUNCH(ctx, (sh_backward_kernel<D, false>), grid, 128, 0, N, LC, table2, mst, Sp, state, mean_t);
  LAUNCH(ctx, (sh_chunk_product_kernel<D, true>), nch, 32, 0, table2, TS2, D * D, N, LC, psi);
  LAUNCH(ctx, (sh_carry_warp_kernel<D, true>), (Sp + 7) / 8, 256, 0, psi, state, nch, Sp);
  LAUNCH(ctx, (sh_backward_kernel<D, true>), grid, 128, 0, N, LC, table2, mst, Sp, state, mean_t);
  if (ex) { ex->var1 = v1; ex->sum_logS = lml1 + 2; ex->a2part = a2part; ex->nch = nch; }
  return GPAR_OK;
}

// filter only (logpdf / decorrelate of many sequences sharing one model): forward passes, no filtered means stored
template <int D>
int filter_shared_d(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, const double* t, const double* y1,
                    const double* rvec, const double* yt, int Sp, double* alpha_t, SharedSmoothExtra* ex) {
  constexpr int TS = D * D + 2 * D + 1;
  int LC = (int)std::min<int64_t>(1024, std::max<int64_t>(64, ((int64_t)Sp * N / 150000 + 31) / 32 * 32));
  if (const char* e = getenv("GPAR_SH_LC")) { int v = atoi(e); if (v >= 32 && v <= 4096 && v % 32 == 0) LC = v; }      // tuning knob
  const int nch = (int)((N + LC - 1) / LC);
  const size_t state_doubles = (size_t)nch * D * Sp;
  CU(ctx->shbuf.reserve(((size_t)N * TS + (size_t)nch * D * D + state_doubles + 16 + (size_t)nch * Sp) * sizeof(double)));
  double* table = ctx->shbuf.as<double>(); double* psi = table + (size_t)N * TS; double* state = psi + (size_t)nch * D * D;
  double* lml1 = state + state_doubles; double* a2part = lml1 + 16;
  CHK(lgssm_run(ctx, kind, &l, &s, &noise, 1, 1, N, t, y1, rvec, nullptr, lml1, nullptr, nullptr, table, lml1 + 2));
  dim3 grid(Sp / 128, nch);
  LAUNCH(ctx, (sh_forward_kernel<D, false>), grid, 128, 0, N, LC, table, yt, Sp, state, (double*)nullptr, (double*)nullptr, (double*)nullptr);
  LAUNCH(ctx, (sh_chunk_product_kernel<D, false>), nch, 32, 0, table, TS, 0, N, LC, psi);
  LAUNCH(ctx, (sh_carry_warp_kernel<D, false>), (Sp + 7) / 8, 256, 0, psi, state, nch, Sp);
  LAUNCH(ctx, (sh_forward_kernel<D, true>), grid, 128, 0, N, LC, table, yt, Sp, state, (double*)nullptr, a2part, alpha_t);
  ex->var1 = nullptr; ex->sum_logS = lml1 + 2; ex->a2part = a2part; ex->nch = nch;
  return GPAR_OK;
}

}  // namespace

namespace {
// tiled transposes between sequence-major [b][N] and time-major [n][Sp]
__global__ void __launch_bounds__(256)
seq_to_time_major_kernel(const double* __restrict__ y, int64_t N, int batch, int Sp, double* __restrict__ yt) {
  __shared__ double tile[32][33];
  const int64_t n0 = (int64_t)blockIdx.x * 32; const int s0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) { const int sq = s0 + r; const int64_t n = n0 + tx; tile[r][tx] = (sq < batch && n < N) ? y[(int64_t)sq * N + n] : 0.0; }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) { const int64_t n = n0 + r; const int sq = s0 + tx; if (n < N && sq < Sp) yt[n * Sp + sq] = tile[tx][r]; }
}
__global__ void __launch_bounds__(256)
time_to_seq_major_kernel(const double* __restrict__ mt, int64_t N, int batch, int Sp, double* __restrict__ m) {
  __shared__ double tile[32][33];
  const int64_t n0 = (int64_t)blockIdx.x * 32; const int s0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) { const int64_t n = n0 + r; const int sq = s0 + tx; tile[r][tx] = (n < N && sq < Sp) ? mt[n * Sp + sq] : 0.0; }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) { const int sq = s0 + r; const int64_t n = n0 + tx; if (sq < batch && n < N) m[(int64_t)sq * N + n] = tile[tx][r]; }
}
// var[b][n] = var1[n] for every sequence;  lml[b] = -1/2 (N log 2pi + sum log S + sum_c a2part[c][b])
__global__ void shared_finish_kernel(const double* __restrict__ var1, const double* __restrict__ sum_logS, const double* __restrict__ a2part,
                                     int nch, int Sp, int64_t N, int batch, double* __restrict__ var, double* __restrict__ lml) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (var && i < (int64_t)batch * N) var[i] = var1[i % N];
  if (lml && i < batch) {
    double a = 0.0;
    for (int c = 0; c < nch; c++) a += a2part[(int64_t)c * Sp + i];
    lml[i] = -0.5 * ((double)N * 1.8378770664093454835606594728112 + sum_logS[0] + a);
  }
}
}  // namespace

// Smoothed means m^s_k[1] of Sp (multiple of 128) time-major sequences yt[n][Sp] sharing one model on the times t
// (noise vector rvec or scalar noise).  y1: any one sequence-major sequence of length N (only the data-independent
// gains are taken from it).  mean_t: [N][Sp].
static int smooth_shared_dispatch(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, const double* t, const double* y1,
                                  const double* rvec, const double* yt, int Sp, double* mean_t, SharedSmoothExtra* ex) {
  if (Sp % 128 != 0) return gpar_fail(ctx, GPAR_ERR_INVALID, "smooth_shared: the sequence count must be padded to a multiple of 128");
  switch (kind) {
    case GPAR_MATERN12: return smooth_shared_d<1>(ctx, kind, l, s, noise, N, t, y1, rvec, yt, Sp, mean_t, ex);
    case GPAR_MATERN32: return smooth_shared_d<2>(ctx, kind, l, s, noise, N, t, y1, rvec, yt, Sp, mean_t, ex);
    case GPAR_MATERN52: return smooth_shared_d<3>(ctx, kind, l, s, noise, N, t, y1, rvec, yt, Sp, mean_t, ex);
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "kernel code %d has no state-space form (use Matern12/32/52)", kind);
  }
}
int lgssm_smooth_shared(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, const double* t, const double* y1,
                        const double* rvec, const double* yt, int Sp, double* mean_t) {
  return smooth_shared_dispatch(ctx, kind, l, s, noise, N, t, y1, rvec, yt, Sp, mean_t, nullptr);
}

// The same for sequence-major data (gpar_lgssm_smooth with several sequences — they always share theta):
// y [batch][N] -> mean, var [batch][N], lml [batch] (nullable).  Transposes in and out; var is the one shared sequence.
int lgssm_smooth_shared_seqmajor(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, int batch, const double* t,
                                 const double* y, const double* rvec, double* d_mean, double* d_var, double* d_lml) {
  const int Sp = (batch + 127) / 128 * 128;
  CU(ctx->panelB.reserve((size_t)2 * N * Sp * sizeof(double)));      // (kal_b is the single-sequence smoother's own scratch)
  double* yt = ctx->panelB.as<double>(); double* mt = yt + (size_t)N * Sp;
  dim3 tg((unsigned)((N + 31) / 32), Sp / 32);
  LAUNCH(ctx, seq_to_time_major_kernel, tg, 256, 0, y, N, batch, Sp, yt);
  SharedSmoothExtra ex;
  CHK(smooth_shared_dispatch(ctx, kind, l, s, noise, N, t, y, rvec, yt, Sp, mt, &ex));
  LAUNCH(ctx, time_to_seq_major_kernel, tg, 256, 0, mt, N, batch, Sp, d_mean);
  const int64_t total = (int64_t)batch * N;
  LAUNCH(ctx, shared_finish_kernel, (unsigned)((total + 255) / 256), 256, 0, ex.var1, ex.sum_logS, ex.a2part, ex.nch, Sp, N, batch, d_var, d_lml);
  return GPAR_OK;
}

// logpdf / decorrelate of `batch` sequence-major sequences sharing one model: lml [batch], alpha [batch][N] (nullable).
int lgssm_filter_shared_seqmajor(gpar_ctx* ctx, int kind, double l, double s, double noise, int64_t N, int batch, const double* t,
                                 const double* y, const double* rvec, double* d_alpha, double* d_lml) {
  const int Sp = (batch + 127) / 128 * 128;
  CU(ctx->panelB.reserve((size_t)(d_alpha ? 2 : 1) * N * Sp * sizeof(double)));
  double* yt = ctx->panelB.as<double>(); double* at = d_alpha ? yt + (size_t)N * Sp : nullptr;
  dim3 tg((unsigned)((N + 31) / 32), Sp / 32);
  LAUNCH(ctx, seq_to_time_major_kernel, tg, 256, 0, y, N, batch, Sp, yt);
  SharedSmoothExtra ex;
  switch (kind) {
    case GPAR_MATERN12: CHK(filter_shared_d<1>(ctx, kind, l, s, noise, N, t, y, rvec, yt, Sp, at, &ex)); break;
    case GPAR_MATERN32: CHK(filter_shared_d<2>(ctx, kind, l, s, noise, N, t, y, rvec, yt, Sp, at, &ex)); break;
    case GPAR_MATERN52: CHK(filter_shared_d<3>(ctx, kind, l, s, noise, N, t, y, rvec, yt, Sp, at, &ex)); break;
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "kernel code %d has no state-space form (use Matern12/32/52)", kind);
  }
  if (d_alpha) LAUNCH(ctx, time_to_seq_major_kernel, tg, 256, 0, at, N, batch, Sp, d_alpha);
  LAUNCH(ctx, shared_finish_kernel, (unsigned)((batch + 255) / 256), 256, 0, (const double*)nullptr, ex.sum_logS, ex.a2part, ex.nch, Sp, N, batch,
         (double*)nullptr, d_lml);
  return GPAR_OK;
}
