// scaled.cu — the scaled-GPAR objective and q(u): pseudo-point DTC on f_x with the temporal GP
// f_t + eps as structured noise, whitened in O(N) by the Matern LGSSM.
//
// Replaces compute_gpar_dtc_objective (src/gp/dtc.jl:83-128) and compute_q_u
// (src/gp/gpar_scaled_inference.jl:141-196).  The reference runs M+1 sequential Kalman filters
// (one per column of Cfu, each re-doing the identical covariance recursion, dtc.jl:108-117),
// builds a dense N x N noise matrix for its log-determinant (dtc.jl:98-99,123) and forms A (M x N).
// Here:
//  1. ONE temporally-parallel filter pass over (t, y) (kalman.cu) yields alpha, sum alpha^2,
//     logdet Sigma_y = sum log S_k (the O(N) identity) and the shared per-step table
//     [Phi_k = (I - K_k H) A_k, K_k, H A_k, S_k^{-1/2}].
//  2. all M columns are whitened by an affine scan with those shared matrices, one thread per
//     column, chunked along N: pass 1 evaluates Kuf (once, into the DMMA panel layout) and the
//     chunk's zero-state response; a carry scan propagates the 3-vector filter means across chunks;
//     pass 2 whitens the panel in place and accumulates g = beta^T alpha.
//  3. G = beta^T beta by the DMMA panel SYRK (panel_syrk.cu); the M x M tail follows dtc.jl:119-125.
#include "lgssm_math.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int WH_GROUPS = 256;   // 4-step groups per whitening chunk (1024 steps)


// ---- shared-memory staging of the per-step table (Phi_k, K_k, HA_k, S_k^-1/2) -------------------
// Every thread of a whitening block walks the same steps, and the loads of a step's row sit on the
// critical path of the recurrence; rows are therefore streamed with cp.async into a double-buffered
// shared-memory window WH_SUB steps ahead of their use.
constexpr int WH_SUB = 64;      // steps per staged window (16 groups of 4)
__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NKEEP> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NKEEP) : "memory"); }
template <int TS>
__device__ __forceinline__ void stage_table(double* buf, const double* __restrict__ table, int64_t step0, int64_t N) {
  int64_t nrows = N - step0; if (nrows > WH_SUB) nrows = WH_SUB; if (nrows < 0) nrows = 0;
  const int total = (int)nrows * TS;
  const double* src = table + step0 * TS;
  for (int e = threadIdx.x; e < total; e += blockDim.x) cp_async8(buf + e, src + e);
  cp_async_commit();
}

// pass 1: K panel + zero-state chunk response b_c[m] (D doubles) -> resp[(c*D + i)*Mpad + m].
// Each thread carries CT columns (one per M-tile mt0..mt0+CT-1) so that a step's table row
// (Phi_k, K_k — shared by every column) is loaded once per CT kernel evaluations.
template <int KIND, int DX, int D, int CT>
__global__ void __launch_bounds__(GPAR_TILE)
whiten_pass1_kernel(const double* __restrict__ X, const double* __restrict__ Z, int64_t N, int M, int64_t NB4, double inv_l2, double s,
                    const double* __restrict__ table, double* __restrict__ panel, double* __restrict__ resp, int Mpad) {
  constexpr int TS = D * D + 2 * D + 1;
  const int mt0 = blockIdx.x * CT, mi = threadIdx.x;
  double z[CT][DX], ms[CT][D];
  bool mvalid[CT];
  double* out[CT];
  const int64_t g0 = (int64_t)blockIdx.y * WH_GROUPS;
  const int64_t g1 = (g0 + WH_GROUPS < NB4) ? g0 + WH_GROUPS : NB4;
#pragma unroll
  for (int c = 0; c < CT; c++) {
    const int m = (mt0 + c) * GPAR_TILE + mi;
    mvalid[c] = m < M;
#pragma unroll
    for (int d = 0; d < DX; d++) z[c][d] = mvalid[c] ? Z[(int64_t)m * DX + d] : 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) ms[c][i] = 0.0;
    out[c] = panel + (((int64_t)(mt0 + c) * NB4 + g0) * GPAR_TILE + mi) * 4;
  }
  __shared__ double tbl[2][WH_SUB * TS];
  stage_table<TS>(tbl[0], table, g0 * 4, N);
  int cur = 0;
  for (int64_t gs = g0; gs < g1; gs += WH_SUB / 4) {
    if (gs + WH_SUB / 4 < g1) { stage_table<TS>(tbl[cur ^ 1], table, (gs + WH_SUB / 4) * 4, N); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int64_t ge = (gs + WH_SUB / 4 < g1) ? gs + WH_SUB / 4 : g1;
    for (int64_t g = gs; g < ge; g++) {
      double kv[CT][4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int64_t n = g * 4 + j;
        const bool valid = n < N;
        const int64_t nn = valid ? n : N - 1;
        double x[DX], F[D * D], Kg[D];
#pragma unroll
        for (int d = 0; d < DX; d++) x[d] = __ldg(X + nn * DX + d);
        const double* row = tbl[cur] + (int)(n - gs * 4) * TS;
        if (valid) {
#pragma unroll
          for (int i = 0; i < D * D; i++) F[i] = row[i];
#pragma unroll
          for (int i = 0; i < D; i++) Kg[i] = row[D * D + i];
        }
#pragma unroll
        for (int c = 0; c < CT; c++) {
          double d2 = 0.0;
#pragma unroll
          for (int d = 0; d < DX; d++) { double df = x[d] - z[c][d]; d2 = fma(df, df, d2); }
          double dummy; double k = base_kernel_dev<KIND, false>(d2 * inv_l2, dummy);
          k = (valid && mvalid[c]) ? s * k : 0.0;
          kv[c][j] = k;
          if (valid) {   // m <- Phi m + K v
            double nm[D];
#pragma unroll
            for (int i = 0; i < D; i++) { double v = Kg[i] * k;
#pragma unroll
              for (int q = 0; q < D; q++) v = fma(F[i * D + q], ms[c][q], v);
              nm[i] = v; }
#pragma unroll
            for (int i = 0; i < D; i++) ms[c][i] = nm[i];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CT; c++) {
        reinterpret_cast<double2*>(out[c])[0] = make_double2(kv[c][0], kv[c][1]);
        reinterpret_cast<double2*>(out[c])[1] = make_double2(kv[c][2], kv[c][3]);
        out[c] += GPAR_TILE * 4;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
#pragma unroll
  for (int c = 0; c < CT; c++)
#pragma unroll
    for (int i = 0; i < D; i++) resp[((int64_t)blockIdx.y * D + i) * Mpad + (mt0 + c) * GPAR_TILE + mi] = ms[c][i];
}

// chunk transition Psi_c = Phi_{k1-1} ... Phi_{k0}: one warp per chunk; lane l multiplies its own
// contiguous run of steps (independent loads, pipelined), then an ordered shuffle tree multiplies
// the 32 partial products (later steps on the left).
template <int D>
__global__ void __launch_bounds__(32)
chunk_transition_kernel(const double* __restrict__ table, int64_t N, double* __restrict__ psi) {
  constexpr int TS = D * D + 2 * D + 1;
  const int c = blockIdx.x, lane = threadIdx.x;
  const int64_t k0 = (int64_t)c * WH_GROUPS * 4, k1 = (k0 + (int64_t)WH_GROUPS * 4 < N) ? k0 + (int64_t)WH_GROUPS * 4 : N;
  constexpr int PER = WH_GROUPS * 4 / 32;
  double Pm[D * D];
#pragma unroll
  for (int i = 0; i < D * D; i++) Pm[i] = (i / D == i % D) ? 1.0 : 0.0;
  const int64_t a0 = k0 + (int64_t)lane * PER;
#pragma unroll 4
  for (int q = 0; q < PER; q++) {
    const int64_t k = a0 + q;
    if (k < k1) {
      const double* row = table + k * TS;
      double F[D * D], R[D * D];
#pragma unroll
      for (int i = 0; i < D * D; i++) F[i] = __ldg(row + i);
      matmul<D>(F, Pm, R);
#pragma unroll
      for (int i = 0; i < D * D; i++) Pm[i] = R[i];
    }
  }
  // ordered reduction: after round d, lane l (l % 2d == 0) holds the product of lanes l .. l+2d-1
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double O[D * D], R[D * D];
#pragma unroll
    for (int i = 0; i < D * D; i++) O[i] = __shfl_down_sync(0xffffffffu, Pm[i], d);
    matmul<D>(O, Pm, R);     // the partner covers LATER steps: it multiplies from the left
#pragma unroll
    for (int i = 0; i < D * D; i++) Pm[i] = R[i];
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < D * D; i++) psi[(int64_t)c * D * D + i] = Pm[i];
  }
}

// carry scan over chunks, one thread per column: start[c] = state entering chunk c (in place over
// resp).  The chain is latency-bound (one dependent 3x3 mat-vec per chunk), so the loads of the next
// 8 chunks are issued ahead of the dependent arithmetic.
template <int D>
__global__ void __launch_bounds__(32)
carry_scan_kernel(const double* __restrict__ psi, double* __restrict__ resp, int nch, int Mpad) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  constexpr int PF = 8;
  double st[D];
#pragma unroll
  for (int i = 0; i < D; i++) st[i] = 0.0;
  for (int c0 = 0; c0 < nch; c0 += PF) {
    double b[PF][D], ps[PF][D * D];
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int c = c0 + u;
      if (c < nch) {
#pragma unroll
        for (int i = 0; i < D; i++) b[u][i] = resp[((int64_t)c * D + i) * Mpad + m];
#pragma unroll
        for (int i = 0; i < D * D; i++) ps[u][i] = __ldg(psi + (int64_t)c * D * D + i);
      }
    }
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int c = c0 + u;
      if (c < nch) {
        double nx[D];
#pragma unroll
        for (int i = 0; i < D; i++) { resp[((int64_t)c * D + i) * Mpad + m] = st[i]; double a = b[u][i];
#pragma unroll
          for (int q = 0; q < D; q++) a = fma(ps[u][i * D + q], st[q], a);
          nx[i] = a; }
#pragma unroll
        for (int i = 0; i < D; i++) st[i] = nx[i];
      }
    }
  }
}

// pass 2: in-place whitening of the panel: beta = (K - HA m) / sqrt(S); m <- Phi m + K_k K; g partials.
// CT columns per thread share each step's table row, as in pass 1.
template <int D, int CT>
__global__ void __launch_bounds__(GPAR_TILE)
whiten_pass2_kernel(int64_t N, int64_t NB4, const double* __restrict__ table, const double* __restrict__ alpha,
                    const double* pin, double* pout, const double* __restrict__ start, double* __restrict__ gpart, int Mpad) {
  constexpr int TS = D * D + 2 * D + 1;
  const int mt0 = blockIdx.x * CT, mi = threadIdx.x;
  const int64_t g0 = (int64_t)blockIdx.y * WH_GROUPS;
  const int64_t g1 = (g0 + WH_GROUPS < NB4) ? g0 + WH_GROUPS : NB4;
  const double* in[CT];
  double* io[CT];
  double ms[CT][D], gacc[CT];
#pragma unroll
  for (int c = 0; c < CT; c++) {
    in[c] = pin + (((int64_t)(mt0 + c) * NB4 + g0) * GPAR_TILE + mi) * 4;
    io[c] = pout + (((int64_t)(mt0 + c) * NB4 + g0) * GPAR_TILE + mi) * 4;
#pragma unroll
    for (int i = 0; i < D; i++) ms[c][i] = start[((int64_t)blockIdx.y * D + i) * Mpad + (mt0 + c) * GPAR_TILE + mi];
    gacc[c] = 0.0;
  }
  __shared__ double tbl[2][WH_SUB * TS];
  stage_table<TS>(tbl[0], table, g0 * 4, N);
  int cur = 0;
  for (int64_t gs = g0; gs < g1; gs += WH_SUB / 4) {
    if (gs + WH_SUB / 4 < g1) { stage_table<TS>(tbl[cur ^ 1], table, (gs + WH_SUB / 4) * 4, N); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int64_t ge = (gs + WH_SUB / 4 < g1) ? gs + WH_SUB / 4 : g1;
    for (int64_t g = gs; g < ge; g++) {
      double kv[CT][4];
#pragma unroll
      for (int c = 0; c < CT; c++) {
        double2 a01 = reinterpret_cast<const double2*>(in[c])[0], a23 = reinterpret_cast<const double2*>(in[c])[1];
        kv[c][0] = a01.x; kv[c][1] = a01.y; kv[c][2] = a23.x; kv[c][3] = a23.y;
        in[c] += GPAR_TILE * 4;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int64_t n = g * 4 + j;
        if (n < N) {
          const double* row = tbl[cur] + (int)(n - gs * 4) * TS;
          double F[D * D], Kg[D], ha[D];
#pragma unroll
          for (int i = 0; i < D * D; i++) F[i] = row[i];
#pragma unroll
          for (int i = 0; i < D; i++) { Kg[i] = row[D * D + i]; ha[i] = row[D * D + D + i]; }
          const double rs = row[D * D + 2 * D], an = __ldg(alpha + n);
#pragma unroll
          for (int c = 0; c < CT; c++) {
            double pred = 0.0;
#pragma unroll
            for (int q = 0; q < D; q++) pred = fma(ha[q], ms[c][q], pred);
            const double k = kv[c][j];
            const double beta = (k - pred) * rs;
            double nm[D];
#pragma unroll
            for (int i = 0; i < D; i++) { double v = Kg[i] * k;
#pragma unroll
              for (int q = 0; q < D; q++) v = fma(F[i * D + q], ms[c][q], v);
              nm[i] = v; }
#pragma unroll
            for (int i = 0; i < D; i++) ms[c][i] = nm[i];
            kv[c][j] = beta;
            gacc[c] = fma(beta, an, gacc[c]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CT; c++) {
        reinterpret_cast<double2*>(io[c])[0] = make_double2(kv[c][0], kv[c][1]);
        reinterpret_cast<double2*>(io[c])[1] = make_double2(kv[c][2], kv[c][3]);
        io[c] += GPAR_TILE * 4;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
#pragma unroll
  for (int c = 0; c < CT; c++) gpart[(int64_t)blockIdx.y * Mpad + (mt0 + c) * GPAR_TILE + mi] = gacc[c];
}

// bare Kuu + jitter I
template <int KIND>
__global__ void kuu_plain_kernel(const double* __restrict__ Z, int M, int D, double inv_l2, double s, double jitter, double* __restrict__ K) {
  int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= M) return;
  double d2 = 0.0;
  for (int d = 0; d < D; d++) { double df = Z[(int64_t)a * D + d] - Z[(int64_t)b * D + d]; d2 = fma(df, df, d2); }
  double dummy; double k = base_kernel_dev<KIND, false>(d2 * inv_l2, dummy);
  K[(int64_t)a + (int64_t)b * M] = s * k + (a == b ? jitter : 0.0);
}
__global__ void trace_add_identity2_kernel(double* B, int M) {
  for (int i = threadIdx.x; i < M; i += blockDim.x) B[(int64_t)i * M + i] += 1.0;
}
__global__ void logdet2_kernel(const double* L, int M, double* out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) acc += log(L[(int64_t)i * M + i]);
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = 2.0 * r;
}
// out (M x M col-major) = upper triangle of L^T (U_u = chol.U), zeros below
__global__ void lower_to_upper_kernel(const double* __restrict__ L, int M, double* __restrict__ U) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * M) return;
  int r = (int)(e % M), c = (int)(e / M);
  U[e] = (r <= c) ? L[(int64_t)c + (int64_t)r * M] : 0.0;
}
__global__ void mirror_lower_kernel(double* A, int M) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * M) return;
  int r = (int)(e % M), c = (int)(e / M);
  if (r < c) A[e] = A[(int64_t)c + (int64_t)r * M];
}
// Bt (M x N col-major) from the panel: Bt[m + n*M] = panel(m, n)
__global__ void panel_to_dense_t_kernel(const double* __restrict__ panel, int64_t N, int M, int64_t NB4, double* __restrict__ Bt) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= N * M) return;
  int m = (int)(e % M); int64_t n = e / M;
  Bt[e] = panel[((((int64_t)(m / GPAR_TILE)) * NB4 + n / 4) * GPAR_TILE + (m % GPAR_TILE)) * 4 + (n % 4)];
}

template <int KIND, int D, int CT>
int launch_pass1_dx(gpar_ctx* ctx, int DX, dim3 grid, const double* X, const double* Z, int64_t N, int M, int64_t NB4, double inv_l2, double s,
                    const double* table, double* panel, double* resp, int Mpad) {
#define CASE_DX(DD) case DD: LAUNCH(ctx, (whiten_pass1_kernel<KIND, DD, D, CT>), grid, GPAR_TILE, 0, X, Z, N, M, NB4, inv_l2, s, table, panel, resp, Mpad); break;
  switch (DX) {
    CASE_DX(1) CASE_DX(2) CASE_DX(3) CASE_DX(4) CASE_DX(5) CASE_DX(6) CASE_DX(7) CASE_DX(8)
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "input dimension D=%d not supported (1..8)", DX);
  }
#undef CASE_DX
  return GPAR_OK;
}
template <int D, int CT>
int launch_pass1(gpar_ctx* ctx, int k_out, dim3 grid, const double* X, const double* Z, int64_t N, int M, int64_t NB4, double inv_l2, double s,
                 const double* table, double* panel, double* resp, int Mpad) {
  switch (k_out) {
    case GPAR_EQ: return launch_pass1_dx<GPAR_EQ, D, CT>(ctx, ctx->D, grid, X, Z, N, M, NB4, inv_l2, s, table, panel, resp, Mpad);
    case GPAR_MATERN12: return launch_pass1_dx<GPAR_MATERN12, D, CT>(ctx, ctx->D, grid, X, Z, N, M, NB4, inv_l2, s, table, panel, resp, Mpad);
    case GPAR_MATERN32: return launch_pass1_dx<GPAR_MATERN32, D, CT>(ctx, ctx->D, grid, X, Z, N, M, NB4, inv_l2, s, table, panel, resp, Mpad);
    case GPAR_MATERN52: return launch_pass1_dx<GPAR_MATERN52, D, CT>(ctx, ctx->D, grid, X, Z, N, M, NB4, inv_l2, s, table, panel, resp, Mpad);
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "unknown output kernel code %d", k_out);
  }
}

struct ScaledStats { double* G; double* g; double sum_logS, sum_a2; int Mpad; int64_t Npad; double *table, *alpha, *beta; int nch; };

// Steps 1-3 of the header comment.  Leaves G (M x M) and g (M) on the device.
template <int D>
int scaled_stats_d(gpar_ctx* ctx, int k_out, double time_l, double time_s, double out_l, double out_s, double noise, ScaledStats* st, bool keep_k) {
  constexpr int TS = D * D + 2 * D + 1;
  const int64_t N = ctx->N; const int M = (int)ctx->M;
  const int Mpad = (M + GPAR_TILE - 1) / GPAR_TILE * GPAR_TILE;
  const int64_t npad_to = keep_k ? 128 : GPAR_KT;        // the gradient's panel GEMM works on 128-step blocks
  const int64_t Npad = (N + npad_to - 1) / npad_to * npad_to;
  const int64_t NB4 = Npad / 4;
  const int nch = (int)((NB4 + WH_GROUPS - 1) / WH_GROUPS);
  const int T = Mpad / GPAR_TILE;
  CU(ctx->panelK.reserve((size_t)Npad * Mpad * sizeof(double)));
  if (keep_k) CU(ctx->panelB.reserve((size_t)Npad * Mpad * sizeof(double)));
  CU(ctx->kal_e.reserve(((size_t)N * TS + (size_t)N + 16) * sizeof(double)));                 // table, alpha, sums
  CU(ctx->gpart.reserve(((size_t)nch * D * Mpad + (size_t)nch * D * D + (size_t)nch * Mpad + Mpad) * sizeof(double)));
  const size_t MM = (size_t)M * M;
  CU(ctx->kal_d.reserve((MM + Mpad) * sizeof(double)));
  double* table = ctx->kal_e.as<double>(); double* alpha = table + (size_t)N * TS; double* sums = alpha + N; double* lml = sums + 2;
  double* resp = ctx->gpart.as<double>(); double* psi = resp + (size_t)nch * D * Mpad; double* gp = psi + (size_t)nch * D * D;
  double* G = ctx->kal_d.as<double>(); double* g = G + MM;
  const int kind_time = D == 1 ? GPAR_MATERN12 : (D == 2 ? GPAR_MATERN32 : GPAR_MATERN52);
  CHK(lgssm_run(ctx, kind_time, &time_l, &time_s, &noise, 1, 1, N, ctx->t.as<double>(), ctx->y.as<double>(), nullptr,
                alpha, lml, nullptr, nullptr, table, sums));
  const double inv_l2 = 1.0 / (out_l * out_l);
  const double* X = ctx->X.as<double>(); const double* Z = ctx->Z.as<double>();
  double* panel = ctx->panelK.as<double>();
  double* beta = keep_k ? ctx->panelB.as<double>() : panel;     // in place unless K is needed again (gradient)
  // column tiles per thread: amortises the shared per-step table loads (registers limit it for large D)
  int CT = (T % 2 == 0) ? 2 : 1;   // measured best on B200 (CT = 1 / 2 / 4: 8.5 / 7.6 / 7.9 ms at N = 1M, M = 1024)
  if (const char* e = getenv("GPAR_WH_CT")) { int v = atoi(e); if ((v == 1 || v == 2 || v == 4) && T % v == 0) CT = v; }   // tuning knob
  dim3 grid(T / CT, nch);
  if (CT == 4) CHK((launch_pass1<D, 4>(ctx, k_out, grid, X, Z, N, M, NB4, inv_l2, out_s, table, panel, resp, Mpad)));
  else if (CT == 2) CHK((launch_pass1<D, 2>(ctx, k_out, grid, X, Z, N, M, NB4, inv_l2, out_s, table, panel, resp, Mpad)));
  else CHK((launch_pass1<D, 1>(ctx, k_out, grid, X, Z, N, M, NB4, inv_l2, out_s, table, panel, resp, Mpad)));
  LAUNCH(ctx, chunk_transition_kernel<D>, nch, 32, 0, table, N, psi);
  LAUNCH(ctx, carry_scan_kernel<D>, (Mpad + 31) / 32, 32, 0, psi, resp, nch, Mpad);
  if (CT == 4) LAUNCH(ctx, (whiten_pass2_kernel<D, 4>), grid, GPAR_TILE, 0, N, NB4, table, alpha, panel, beta, resp, gp, Mpad);
  else if (CT == 2) LAUNCH(ctx, (whiten_pass2_kernel<D, 2>), grid, GPAR_TILE, 0, N, NB4, table, alpha, panel, beta, resp, gp, Mpad);
  else LAUNCH(ctx, (whiten_pass2_kernel<D, 1>), grid, GPAR_TILE, 0, N, NB4, table, alpha, panel, beta, resp, gp, Mpad);
  CHK(launch_reduce_gh(ctx, gp, nch, Mpad, 1, g));
  cudaEventRecord(ctx->pev[0], ctx->stream);
  CHK(panel_syrk_run(ctx, beta, nullptr, Npad, Mpad, M, false, G, nullptr));
  ctx->phase_valid = true;
  double hs[2];
  CU(cudaMemcpyAsync(hs, sums, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  st->G = G; st->g = g; st->sum_logS = hs[0]; st->sum_a2 = hs[1]; st->Mpad = Mpad; st->Npad = Npad;
  st->table = table; st->alpha = alpha; st->beta = beta; st->nch = nch;
  return GPAR_OK;
}

int scaled_stats(gpar_ctx* ctx, int k_time, int k_out, double time_l, double time_s, double out_l, double out_s, double noise, ScaledStats* st,
                 bool keep_k = false) {
  switch (k_time) {
    case GPAR_MATERN12: return scaled_stats_d<1>(ctx, k_out, time_l, time_s, out_l, out_s, noise, st, keep_k);
    case GPAR_MATERN32: return scaled_stats_d<2>(ctx, k_out, time_l, time_s, out_l, out_s, noise, st, keep_k);
    case GPAR_MATERN52: return scaled_stats_d<3>(ctx, k_out, time_l, time_s, out_l, out_s, noise, st, keep_k);
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "time kernel code %d has no state-space form (use Matern12/32/52)", k_time);
  }
}

int check_scaled(gpar_ctx* ctx, const char* who) {
  if (ctx->N < 1 || ctx->M < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: inputs and pseudo-inputs must be set", who);
  if (ctx->D != ctx->Dz) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: X has D=%d but Z has D=%d", who, ctx->D, ctx->Dz);
  if (ctx->Nt != ctx->N) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: %lld time locations but %lld inputs", who, (long long)ctx->Nt, (long long)ctx->N);
  if (ctx->Ny != ctx->N || ctx->ybatch < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: outputs length %lld != N %lld", who, (long long)ctx->Ny, (long long)ctx->N);
  return GPAR_OK;
}

int launch_kuu_plain(gpar_ctx* ctx, int kind, double l, double s, double jitter, double* K) {
  const int M = (int)ctx->M, D = ctx->Dz;
  dim3 kgrid((M + 127) / 128, M);
  const double inv_l2 = 1.0 / (l * l);
  const double* Zd = ctx->Z.as<double>();
  switch (kind) {
    case GPAR_EQ: LAUNCH(ctx, kuu_plain_kernel<GPAR_EQ>, kgrid, 128, 0, Zd, M, D, inv_l2, s, jitter, K); break;
    case GPAR_MATERN12: LAUNCH(ctx, kuu_plain_kernel<GPAR_MATERN12>, kgrid, 128, 0, Zd, M, D, inv_l2, s, jitter, K); break;
    case GPAR_MATERN32: LAUNCH(ctx, kuu_plain_kernel<GPAR_MATERN32>, kgrid, 128, 0, Zd, M, D, inv_l2, s, jitter, K); break;
    default: LAUNCH(ctx, kuu_plain_kernel<GPAR_MATERN52>, kgrid, 128, 0, Zd, M, D, inv_l2, s, jitter, K); break;
  }
  return GPAR_OK;
}

}  // namespace

static const double LOG2PI_S = 1.8378770664093454835606594728112;

extern "C" {

int gpar_scaled_dtc(gpar_ctx* ctx, int k_time, int k_out, const double theta[5], double* dtc, double* A_or_null) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!theta || !dtc) return gpar_fail(ctx, GPAR_ERR_INVALID, "scaled_dtc: theta and dtc must not be NULL");
  CHK(check_scaled(ctx, "scaled_dtc"));
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx);
  // unpack_gpar (util.jl:45-55); variances squared, noise squared (dtc.jl:31-37)
  double pv[5];
  for (int i = 0; i < 5; i++) pv[i] = exp(theta[i]) + 1e-3;
  const double time_l = pv[0], time_s = pv[1] * pv[1], out_l = pv[2], out_s = pv[3] * pv[3], noise = pv[4] * pv[4];
  ScaledStats st;
  CHK(scaled_stats(ctx, k_time, k_out, time_l, time_s, out_l, out_s, noise, &st));
  const int M = (int)ctx->M; const int64_t N = ctx->N; const size_t MM = (size_t)M * M;
  cublasSetStream(ctx->blas, ctx->stream); cusolverDnSetStream(ctx->solver, ctx->stream);
  CU(ctx->dense.reserve((2 * MM + 2 * (size_t)M + 16) * sizeof(double)));
  double* Lu = ctx->dense.as<double>(); double* Bm = Lu + MM; double* cvec = Bm + MM; double* sc = cvec + 2 * M;
  int lwork = 0;
  CS(cusolverDnDpotrf_bufferSize(ctx->solver, CUBLAS_FILL_MODE_LOWER, M, Lu, M, &lwork));
  CU(ctx->tailws.reserve((size_t)lwork * sizeof(double)));
  CU(ctx->info.reserve(4 * sizeof(int)));
  int* dinfo = ctx->info.as<int>();
  CHK(launch_kuu_plain(ctx, k_out, out_l, out_s, noise, Lu));       // cov(u) = Kuu + noise_sigma^2 I (dtc.jl:35,119)
  CS(cusolverDnDpotrf(ctx->solver, CUBLAS_FILL_MODE_LOWER, M, Lu, M, ctx->tailws.as<double>(), lwork, dinfo));
  CU(cudaMemcpyAsync(Bm, st.G, MM * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  const double one = 1.0;
  CB(cublasDtrsm(ctx->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, M, &one, Lu, M, Bm, M));
  CB(cublasDtrsm(ctx->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, M, &one, Lu, M, Bm, M));
  LAUNCH(ctx, trace_add_identity2_kernel, 1, 256, 0, Bm, M);
  CS(cusolverDnDpotrf(ctx->solver, CUBLAS_FILL_MODE_LOWER, M, Bm, M, ctx->tailws.as<double>(), lwork, dinfo + 1));
  LAUNCH(ctx, logdet2_kernel, 1, 256, 0, Bm, M, sc);
  CU(cudaMemcpyAsync(cvec, st.g, M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CB(cublasDtrsv(ctx->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, Lu, M, cvec, 1));
  CB(cublasDtrsv(ctx->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, Bm, M, cvec, 1));
  CB(cublasSetPointerMode(ctx->blas, CUBLAS_POINTER_MODE_DEVICE));
  cublasStatus_t bst = cublasDdot(ctx->blas, M, cvec, 1, cvec, 1, sc + 1);
  cublasSetPointerMode(ctx->blas, CUBLAS_POINTER_MODE_HOST);
  CB(bst);
  if (A_or_null) {   // A = chol(cov(u)).U' \ beta'  (dtc.jl:119), M x N column-major — small problems only
    CU(ctx->kal_b.reserve((size_t)N * M * sizeof(double)));
    double* Bt = ctx->kal_b.as<double>();
    LAUNCH(ctx, panel_to_dense_t_kernel, (int)(((size_t)N * M + 255) / 256), 256, 0, ctx->panelK.as<double>(), N, M, st.Npad / 4, Bt);
    CB(cublasDtrsm(ctx->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, (int)N, &one, Lu, M, Bt, M));
  }
  timer.stop();
  double hs[2]; int hinfo[2];
  CU(cudaMemcpyAsync(hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hinfo, dinfo, sizeof(hinfo), cudaMemcpyDeviceToHost, ctx->stream));
  if (A_or_null) CU(cudaMemcpyAsync(A_or_null, ctx->kal_b.p, (size_t)N * M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (hinfo[0] != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(cov(u)) failed: leading minor %d is not positive definite", hinfo[0]);
  if (hinfo[1] != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(A*A' + I) failed: leading minor %d is not positive definite", hinfo[1]);
  // dtc.jl:122-125 with logdet(noise_matrix) = sum log S_k
  *dtc = -0.5 * ((double)N * LOG2PI_S + st.sum_logS + hs[0] + st.sum_a2 - hs[1]);
  return GPAR_OK;
}

int gpar_compute_q_u(gpar_ctx* ctx, int k_time, int k_out, const double params[5], double* m_e, double* Dinv, double* U_u) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!params || !m_e || !Dinv || !U_u) return gpar_fail(ctx, GPAR_ERR_INVALID, "compute_q_u: NULL argument");
  CHK(check_scaled(ctx, "compute_q_u"));
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx);
  const double time_l = params[0], time_s = params[1] * params[1], out_l = params[2], out_s = params[3] * params[3], noise = params[4] * params[4];
  ScaledStats st;
  CHK(scaled_stats(ctx, k_time, k_out, time_l, time_s, out_l, out_s, noise, &st));
  const int M = (int)ctx->M; const size_t MM = (size_t)M * M;
  cublasSetStream(ctx->blas, ctx->stream); cusolverDnSetStream(ctx->solver, ctx->stream);
  CU(ctx->dense.reserve((3 * MM + 2 * (size_t)M + 16) * sizeof(double)));
  double* Lu = ctx->dense.as<double>(); double* Dm = Lu + MM; double* Uu = Dm + MM; double* vec = Uu + MM;
  int lwork = 0, lwork2 = 0;
  CS(cusolverDnDpotrf_bufferSize(ctx->solver, CUBLAS_FILL_MODE_LOWER, M, Lu, M, &lwork));
  CS(cusolverDnDpotri_bufferSize(ctx->solver, CUBLAS_FILL_MODE_LOWER, M, Dm, M, &lwork2));
  lwork = std::max(lwork, lwork2);
  CU(ctx->tailws.reserve((size_t)lwork * sizeof(double)));
  CU(ctx->info.reserve(4 * sizeof(int)));
  int* dinfo = ctx->info.as<int>();
  CHK(launch_kuu_plain(ctx, k_out, out_l, out_s, 0.0, Lu));         // bare Cuu (gpar_scaled_inference.jl:157-159)
  CS(cusolverDnDpotrf(ctx->solver, CUBLAS_FILL_MODE_LOWER, M, Lu, M, ctx->tailws.as<double>(), lwork, dinfo));
  LAUNCH(ctx, lower_to_upper_kernel, (int)((MM + 255) / 256), 256, 0, Lu, M, Uu);
  CU(cudaMemcpyAsync(Dm, st.G, MM * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  const double one = 1.0;
  CB(cublasDtrsm(ctx->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, M, &one, Lu, M, Dm, M));
  CB(cublasDtrsm(ctx->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, M, &one, Lu, M, Dm, M));
  LAUNCH(ctx, trace_add_identity2_kernel, 1, 256, 0, Dm, M);          // D = B_ef B_ef' + I (:187)
  CS(cusolverDnDpotrf(ctx->solver, CUBLAS_FILL_MODE_LOWER, M, Dm, M, ctx->tailws.as<double>(), lwork, dinfo + 1));
  // m_e = chol_D \ (B_ef b_y) = L_D^{-T} L_D^{-1} L_u^{-1} g  (:189)
  CU(cudaMemcpyAsync(vec, st.g, M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CB(cublasDtrsv(ctx->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, Lu, M, vec, 1));
  CB(cublasDtrsv(ctx->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, M, Dm, M, vec, 1));
  CB(cublasDtrsv(ctx->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, M, Dm, M, vec, 1));
  // inv(D) (:192)
  CS(cusolverDnDpotri(ctx->solver, CUBLAS_FILL_MODE_LOWER, M, Dm, M, ctx->tailws.as<double>(), lwork, dinfo + 2));
  LAUNCH(ctx, mirror_lower_kernel, (int)((MM + 255) / 256), 256, 0, Dm, M);
  timer.stop();
  int hinfo[3];
  CU(cudaMemcpyAsync(hinfo, dinfo, sizeof(hinfo), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(m_e, vec, M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(Dinv, Dm, MM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(U_u, Uu, MM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (hinfo[0] != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(Cuu) failed: leading minor %d is not positive definite", hinfo[0]);
  if (hinfo[1] != 0 || hinfo[2] != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(D) failed: leading minor %d is not positive definite", hinfo[1]);
  return GPAR_OK;
}

}  // extern "C"
