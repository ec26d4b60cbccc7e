// scaled_small.cu — MANY hyper-parameter candidates of the scaled-GPAR objective at the reference's own problem sizes.
//
// The Nelder-Mead loop of the reference (src/gp/dtc.jl:58-61) evaluates compute_gpar_dtc_objective (dtc.jl:83-128)
// hundreds of times on data of N = 8 496 points and M = 50 ... 81 pseudo-inputs; one evaluation through the large-
// problem pipeline (scaled.cu) is ~30 short launches = 0.4 ms of latency on an otherwise idle B200.  Here `ncand`
// candidates (simplex vertices x restarts) share ONE launch sequence, the candidate being a grid dimension:
//   1. the LGSSM filter of the time kernel on (t, y) for all candidates at once — kalman.cu's chunked scan with one
//      parameter set per "sequence", every sequence reading the same y — leaves per candidate the step table
//      (Phi_k, K_k, H A_k, S_k^-1/2), alpha, sum log S_k and sum alpha_k^2                       (dtc.jl:106);
//   2. beta[:, m] = decorrelate(cov(f, u)[:, m]) for every pseudo-input (dtc.jl:108-117), chunked along time like the
//      whitening of scaled.cu: ss_walk_kernel<false> (zero-start chunk responses; a thread per column, so a step's table
//      row and input point are staged once per block and read as shared-memory broadcasts), ss_chunk_product_kernel +
//      ss_carry_kernel (start state of every chunk), ss_walk_kernel<true> (beta through a transposing tile, g = beta' alpha);
//   3. ss_syrk_kernel: G = beta' beta on the FP64 tensor cores (DMMA), 64 x 64 tiles of the lower triangle, split along N
//      (partial sums added in fixed order by the tail), operands staged in shared memory with register prefetch;
//   4. ss_tail_kernel: one CTA per candidate, one M x M matrix in shared memory at a time: cov(u) = Kuu + sigma^2 I
//      (dtc.jl:35,119) and W = cov(u) + G = L_u Lambda L_u' are factorised in turn — log det Lambda = log det W - log det cov(u),
//      c'c = |L_W^-1 g|^2 — and the value follows (dtc.jl:122-125).
// This is the collapsed form of the objective (G, g instead of A = L_u^-1 beta'), as in scaled.cu; a candidate whose
// cov(u) is too poorly conditioned for it — (max / min diag L_u)^2 > GPAR_ROBUST_COND, the same test — is handed back to
// the caller, which evaluates it through the whitened-panel path of gpar_scaled_dtc.
#include "lgssm_math.cuh"
#include "dmma_pipe.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <limits>
#include <vector>

namespace {

constexpr int SS_TILE = 32;                 // SYRK tile
constexpr double LOG2PI_SS = 1.8378770664093454835606594728112;

struct SmallCand { double inv_l2, out_s, noise; };      // per candidate: 1 / out_l^2, out_s = out_var^2, noise = sigma^2

// kappa(r), r = ||x - z|| / l, with the branch-free exponential of lgssm_math.cuh (argument <= 0): ~1 ulp from the
// library exp of base_kernel_dev, a third fewer FP64 issue slots
template <int KIND>
__device__ __forceinline__ double ss_kappa_r(double r) {
  if (KIND == GPAR_EQ) return exp_nonpos(-0.5 * r * r);
  if (KIND == GPAR_MATERN12) return exp_nonpos(-r);
  if (KIND == GPAR_MATERN32) { const double a = 1.7320508075688772935274463415059 * r; return (1.0 + a) * exp_nonpos(-a); }
  const double a = 2.2360679774997896964091736687313 * r;
  return (1.0 + a + a * a * (1.0 / 3.0)) * exp_nonpos(-a);
}

__device__ __forceinline__ void ss_cp8(double* smem, const double* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void ss_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NP> __device__ __forceinline__ void ss_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NP) : "memory"); }

// Ordered product of the step matrices of every whitening chunk, Psi_c = Phi_{k1-1} ... Phi_{k0}: a warp per (chunk,
// candidate), each lane a contiguous run of steps, then an ordered shuffle tree.
template <int D>
__global__ void __launch_bounds__(32)
ss_chunk_product_kernel(const double* __restrict__ table, int64_t N, int Lc, int nch, double* __restrict__ psi) {
  constexpr int TS = D * D + 2 * D + 1;
  const int c = blockIdx.x, cd = blockIdx.y, lane = threadIdx.x;
  const double* tab = table + (int64_t)cd * N * TS;
  const int per = Lc / 32;
  const int64_t k0 = (int64_t)c * Lc + (int64_t)lane * per;
  double P[D * D];
#pragma unroll
  for (int i = 0; i < D * D; i++) P[i] = (i / D == i % D) ? 1.0 : 0.0;
  for (int q = 0; q < per; q++) {
    const int64_t k = k0 + q;
    if (k < N) {
      double F[D * D], R[D * D];
#pragma unroll
      for (int i = 0; i < D * D; i++) F[i] = __ldg(tab + k * TS + i);
      matmul<D>(F, P, R);
#pragma unroll
      for (int i = 0; i < D * D; i++) P[i] = R[i];
    }
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {       // after round d, lane l (l % 2d == 0) holds the product of lanes l .. l+2d-1 (later lanes on the left)
    double O[D * D], R[D * D];
#pragma unroll
    for (int i = 0; i < D * D; i++) O[i] = __shfl_down_sync(0xffffffffu, P[i], d);
    matmul<D>(O, P, R);
#pragma unroll
    for (int i = 0; i < D * D; i++) P[i] = R[i];
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < D * D; i++) psi[((int64_t)cd * nch + c) * D * D + i] = P[i];
  }
}

// The whitening walk of one chunk: a thread per pseudo-input (column), so that a step's table row and input point are
// read ONCE per block (staged through shared memory in tiles of 32 steps, broadcast reads) instead of once per column.
// FINAL = false: zero-start response of the chunk -> state[cd][c][i][m]; the kernel values u = cov(f, u)[n, m] are parked
//                in beta[cd][m][n] through a transposing shared-memory tile (coalesced 256-byte runs).
// FINAL = true : from the carried start state (same array), reading u back through the tile (no second exp / sqrt per
//                element): beta = (u - HA x) / sqrt(S) overwrites it, g partial sums -> gpart[cd][c][m].
template <int KIND, int D, bool FINAL>
__global__ void __launch_bounds__(192)
ss_walk_kernel(const double* __restrict__ X, const double* __restrict__ Z, int DX, int64_t N, int M, int Mp, int64_t Ns, int Lc, int nch,
               const SmallCand* __restrict__ cand, const double* __restrict__ table, const double* __restrict__ alpha,
               double* __restrict__ state, double* __restrict__ beta, double* __restrict__ gpart) {
  constexpr int TS = D * D + 2 * D + 1;
  extern __shared__ double sm[];
  // [2][32][TS] table rows | [2][32][DX] inputs | [2][32] alpha | [32][Mp + 1] transposing tile (kernel values / beta)
  double* stab = sm; double* sx = stab + 2 * 32 * TS; double* sal = sx + 2 * 32 * DX; double* tile = sal + 2 * 32;
  const int m = threadIdx.x, c = blockIdx.x, cd = blockIdx.y;
  const int64_t k0 = (int64_t)c * Lc;
  const int ntile = Lc / 32;
  const double* tab = table + (int64_t)cd * N * TS;
  const double* al = alpha + (int64_t)cd * N;
  double* bc = beta + (int64_t)cd * Mp * Ns;
  const SmallCand cp = cand[cd];
  const double inv_l = sqrt(cp.inv_l2);
  const bool live = m < M;
  double z[8];
#pragma unroll
  for (int d = 0; d < 8; d++) z[d] = (!FINAL && live && d < DX) ? Z[(int64_t)m * DX + d] : 0.0;
  auto stage = [&](int tl, int buf) {
    const int64_t kb = k0 + (int64_t)tl * 32;
    const int64_t rows = N - kb < 32 ? (N - kb > 0 ? N - kb : 0) : 32;       // steps of this tile inside the sequence
    for (int e = threadIdx.x; e < rows * TS; e += blockDim.x) ss_cp8(stab + buf * 32 * TS + e, tab + kb * TS + e);
    if (!FINAL) for (int e = threadIdx.x; e < rows * DX; e += blockDim.x) ss_cp8(sx + buf * 32 * DX + e, X + kb * DX + e);
    if (FINAL) for (int e = threadIdx.x; e < rows; e += blockDim.x) ss_cp8(sal + buf * 32 + e, al + kb + e);
    ss_commit();
  };
  double x[D];
#pragma unroll
  for (int i = 0; i < D; i++) x[i] = FINAL ? state[(((int64_t)cd * nch + c) * D + i) * Mp + m] : 0.0;
  double gacc = 0.0;
  stage(0, 0);
  for (int tl = 0; tl < ntile; tl++) {
    const int buf = tl & 1;
    const int64_t kb = k0 + (int64_t)tl * 32;
    if (FINAL) {      // the kernel values of this tile, parked in the beta array by the first walk (coalesced runs -> transposed tile)
      for (int e = threadIdx.x; e < 32 * Mp; e += blockDim.x) {
        const int col = e >> 5, sidx = e & 31;
        if (kb + sidx < Ns) ss_cp8(tile + sidx * (Mp + 1) + col, bc + (int64_t)col * Ns + kb + sidx);
      }
      ss_commit();
    }
    if (tl + 1 < ntile) { stage(tl + 1, buf ^ 1); ss_wait<1>(); } else ss_wait<0>();
    __syncthreads();
    const int rows = (int)(N - kb < 32 ? (N - kb > 0 ? N - kb : 0) : 32);
    for (int sidx = 0; sidx < 32; sidx++) {
      double oval = 0.0;
      if (sidx < rows) {
        const double* row = stab + (buf * 32 + sidx) * TS;
        double u = 0.0;
        if (FINAL) u = tile[sidx * (Mp + 1) + m];
        else if (live) {
          double r;
          if (DX == 1) r = fabs(sx[buf * 32 + sidx] - z[0]) * inv_l;        // one input dimension: no square root
          else {
            double d2 = 0.0;
#pragma unroll
            for (int d = 0; d < 8; d++) if (d < DX) { const double df = sx[(buf * 32 + sidx) * DX + d] - z[d]; d2 = fma(df, df, d2); }
            r = sqrt(d2) * inv_l;
          }
          u = cp.out_s * ss_kappa_r<KIND>(r);
        }
        oval = u;
        if (FINAL) {
          double pred = 0.0;
#pragma unroll
          for (int j = 0; j < D; j++) pred = fma(row[D * D + D + j], x[j], pred);
          oval = (u - pred) * row[D * D + 2 * D];
          gacc = fma(oval, sal[buf * 32 + sidx], gacc);
        }
        double nx[D];
#pragma unroll
        for (int i = 0; i < D; i++) { double a = row[D * D + i] * u;
#pragma unroll
          for (int j = 0; j < D; j++) a = fma(row[i * D + j], x[j], a);
          nx[i] = a; }
#pragma unroll
        for (int i = 0; i < D; i++) x[i] = nx[i];
      }
      tile[sidx * (Mp + 1) + m] = oval;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * Mp; e += blockDim.x) {
      const int col = e >> 5, sidx = e & 31;
      if (kb + sidx < Ns) bc[(int64_t)col * Ns + kb + sidx] = tile[sidx * (Mp + 1) + col];
    }
    __syncthreads();
  }
  if (FINAL) gpart[((int64_t)cd * nch + c) * Mp + m] = gacc;
  else {
#pragma unroll
    for (int i = 0; i < D; i++) state[(((int64_t)cd * nch + c) * D + i) * Mp + m] = x[i];
  }
}

// start state of every chunk, in place of its zero-start response: in[0] = 0, in[c+1] = Psi_c in[c] + resp[c].
// A warp per column: every lane composes the affine maps of its K consecutive chunks, the lane maps go through a
// warp-shuffle scan, and the lane walks its chunks again from its exclusive prefix — 2K + 5 dependent steps instead of
// nch (a single candidate at N = 8 496 has 266 chunks: the sequential chain was 30 us of a 170 us evaluation).
template <int D>
__global__ void __launch_bounds__(256)
ss_carry_kernel(const double* __restrict__ psi, double* __restrict__ state, int nch, int Mp) {
  const int lane = threadIdx.x & 31, m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), cd = blockIdx.y;
  if (m >= Mp) return;
  const int K = (nch + 31) / 32;
  const int c0 = lane * K, c1 = min(c0 + K, nch);
  const double* ps = psi + (int64_t)cd * nch * D * D;
  double* st = state + (int64_t)cd * nch * D * Mp + m;
  double P[D * D], r[D];
#pragma unroll
  for (int i = 0; i < D * D; i++) P[i] = (i / D == i % D) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) r[i] = 0.0;
  for (int c = c0; c < c1; c++) {
    double F[D * D], R[D * D], nr[D];
#pragma unroll
    for (int i = 0; i < D * D; i++) F[i] = __ldg(ps + (int64_t)c * D * D + i);
#pragma unroll
    for (int i = 0; i < D; i++) { double a = st[((int64_t)c * D + i) * Mp];
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
  for (int c = c0; c < c1; c++) {
    double nx[D];
#pragma unroll
    for (int i = 0; i < D; i++) { double* p = st + ((int64_t)c * D + i) * Mp; double a = *p; *p = x[i];
#pragma unroll
      for (int j = 0; j < D; j++) a = fma(__ldg(ps + (int64_t)c * D * D + i * D + j), x[j], a);
      nx[i] = a; }
#pragma unroll
    for (int i = 0; i < D; i++) x[i] = nx[i];
  }
}

// the same carry, one thread per column walking the chunks in order: the better choice for a few chunks (many candidates)
template <int D>
__global__ void ss_carry_seq_kernel(const double* __restrict__ psi, double* __restrict__ state, int nch, int Mp) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x, cd = blockIdx.y;
  if (m >= Mp) return;
  constexpr int PF = 6;
  double st[D];
#pragma unroll
  for (int i = 0; i < D; i++) st[i] = 0.0;
  for (int c0 = 0; c0 < nch; c0 += PF) {
    double r[PF][D], ps[PF][D * D];
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int c = c0 + u;
      if (c < nch) {
#pragma unroll
        for (int i = 0; i < D; i++) r[u][i] = state[(((int64_t)cd * nch + c) * D + i) * Mp + m];
#pragma unroll
        for (int i = 0; i < D * D; i++) ps[u][i] = __ldg(psi + ((int64_t)cd * nch + c) * D * D + i);
      }
    }
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int c = c0 + u;
      if (c < nch) {
        double nx[D];
#pragma unroll
        for (int i = 0; i < D; i++) { state[(((int64_t)cd * nch + c) * D + i) * Mp + m] = st[i]; double a = r[u][i];
#pragma unroll
          for (int j = 0; j < D; j++) a = fma(ps[u][i * D + j], st[j], a);
          nx[i] = a; }
#pragma unroll
        for (int i = 0; i < D; i++) st[i] = nx[i];
      }
    }
  }
}

// Gp (per candidate and split: Mp x Mp, both triangles written) = beta' beta over the split's steps.  Block (pair, split,
// cand): 64 x 64 tile (ti >= tj) on the FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4): 8 warps, each a
// 32 x 16 sub-tile = 4 x 2 fragments (6 shared-memory loads per 8 DMMA; rows padded to 36 doubles: a fragment load is
// bank-conflict free per half-warp).  Operands through shared memory in slabs of 32 steps, the next slab fetched into
// registers while the current one is multiplied.
constexpr int SS_ST = 64;
constexpr int SS_LD = SS_TILE + 4;
__global__ void __launch_bounds__(256)
ss_syrk_kernel(const double* __restrict__ beta, int Mp, int64_t Ns, int slabs_per_split, int nsplit, double* __restrict__ Gp) {
  __shared__ double As[SS_ST][SS_LD], Bs[SS_ST][SS_LD];
  int p = blockIdx.x, ti = 0;
  while (p > ti) { p -= ti + 1; ti++; }
  const int tj = p, sp = blockIdx.y, cd = blockIdx.z;
  const double* bc = beta + (int64_t)cd * Mp * Ns;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wy = w >> 2, wx = w & 3;
  const int fr = lane >> 2, fc = lane & 3;                      // fragment row / k (A, B) — C: row fr, columns 2 fc + {0, 1}
  const int lr = threadIdx.x >> 5, lc = threadIdx.x & 31;       // loader: row group (8 rows per pass), step
  const int64_t nslab = Ns / SS_TILE;
  const int64_t s0 = (int64_t)sp * slabs_per_split, s1 = s0 + slabs_per_split < nslab ? s0 + slabs_per_split : nslab;
  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  double ra[8], rb[8];
  auto fetch = [&](int64_t sl) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int ra_row = ti * SS_ST + lr + 8 * q, rb_row = tj * SS_ST + lr + 8 * q;
      ra[q] = ra_row < Mp ? bc[(int64_t)ra_row * Ns + sl * SS_TILE + lc] : 0.0;
      rb[q] = rb_row < Mp ? bc[(int64_t)rb_row * Ns + sl * SS_TILE + lc] : 0.0;
    }
  };
  if (s0 < s1) fetch(s0);
  for (int64_t sl = s0; sl < s1; sl++) {
#pragma unroll
    for (int q = 0; q < 8; q++) { As[lr + 8 * q][lc] = ra[q]; Bs[lr + 8 * q][lc] = rb[q]; }
    __syncthreads();
    if (sl + 1 < s1) fetch(sl + 1);
#pragma unroll
    for (int k4 = 0; k4 < SS_TILE / 4; k4++) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[wy * 32 + i * 8 + fr][k4 * 4 + fc];
#pragma unroll
      for (int j = 0; j < 2; j++) b[j] = Bs[wx * 16 + j * 8 + fr][k4 * 4 + fc];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) dmma::dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  double* Gc = Gp + ((int64_t)cd * nsplit + sp) * Mp * Mp;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int gi = ti * SS_ST + wy * 32 + i * 8 + fr, gj = tj * SS_ST + wx * 16 + j * 8 + 2 * fc + h;
        if (gi < Mp && gj < Mp) {
          Gc[(int64_t)gi + (int64_t)gj * Mp] = acc[i][j][h];
          Gc[(int64_t)gj + (int64_t)gi * Mp] = acc[i][j][h];
        }
      }
}

// In-place Cholesky (lower, column-major, leading dimension ld) of the M x M matrix A in shared memory by the whole
// block; the 1-based index of the first non-positive pivot goes to *info (shared; 0: success).  Blocked by 8 columns:
// the 8 x 8 diagonal block is factorised by one warp (warp-level barriers only), the rows below are solved against it
// one thread per row, and the trailing block takes a rank-8 update — three block barriers per 8 columns instead of one
// or more per column (the factorisation is a chain of M dependent pivots; everything else is latency to hide).
__device__ void ss_chol_inplace(double* A, int M, int ld, int* info) {
  constexpr int NB = 8;
  const int lane = threadIdx.x & 31;
  for (int j0 = 0; j0 < M; j0 += NB) {
    const int nb = M - j0 < NB ? M - j0 : NB;
    __syncthreads();
    if (threadIdx.x < 32) {
      for (int j = 0; j < nb; j++) {
        const double d = A[(j0 + j) + (j0 + j) * ld];
        if (!(d > 0.0)) { if (lane == 0 && *info == 0) *info = j0 + j + 1; break; }
        const double inv = rsqrt(d);
        __syncwarp();
        if (lane > j && lane < nb) A[(j0 + lane) + (j0 + j) * ld] *= inv;
        if (lane == j) A[(j0 + j) + (j0 + j) * ld] = d * inv;
        __syncwarp();
        for (int e = lane; e < NB * NB; e += 32) {
          const int i = e % NB, k = e / NB;
          if (k > j && i >= k && i < nb && k < nb) A[(j0 + i) + (j0 + k) * ld] = fma(-A[(j0 + i) + (j0 + j) * ld], A[(j0 + k) + (j0 + j) * ld], A[(j0 + i) + (j0 + k) * ld]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (*info) return;
    for (int i = j0 + nb + threadIdx.x; i < M; i += blockDim.x) {        // panel: row i of L21 = A21 L11^-T
#pragma unroll
      for (int j = 0; j < NB; j++) {
        if (j < nb) {
          double v = A[i + (j0 + j) * ld];
#pragma unroll
          for (int k = 0; k < NB; k++) if (k < j) v = fma(-A[i + (j0 + k) * ld], A[(j0 + j) + (j0 + k) * ld], v);
          A[i + (j0 + j) * ld] = v / A[(j0 + j) + (j0 + j) * ld];
        }
      }
    }
    __syncthreads();
    for (int k = j0 + nb + (threadIdx.x >> 4); k < M; k += (int)(blockDim.x >> 4)) {        // trailing block: rank-nb update of the lower triangle
      double lk[NB];
#pragma unroll
      for (int j = 0; j < NB; j++) lk[j] = j < nb ? A[k + (j0 + j) * ld] : 0.0;
      for (int i = k + (threadIdx.x & 15); i < M; i += 16) {
        double a = A[i + k * ld];
#pragma unroll
        for (int j = 0; j < NB; j++) if (j < nb) a = fma(-A[i + (j0 + j) * ld], lk[j], a);
        A[i + k * ld] = a;
      }
    }
  }
  __syncthreads();
}
// v <- L^-1 v for a lower-triangular L in shared memory (column-oriented substitution, one barrier per column)
__device__ void ss_trsv_inplace(const double* L, int M, int ld, double* v) {
  for (int j = 0; j < M; j++) {
    __syncthreads();
    const double cj = v[j] / L[j + j * ld];
    __syncthreads();
    if (threadIdx.x == 0) v[j] = cj;
    for (int i = j + 1 + threadIdx.x; i < M; i += blockDim.x) v[i] = fma(-L[i + j * ld], cj, v[i]);
  }
  __syncthreads();
}

// out per candidate: [value, code]; code 0 ok, 1 / 2: Cholesky of cov(u) / of cov(u) + G failed, 3: cov(u) too poorly
// conditioned for the collapsed statistic (the caller re-evaluates the candidate through the whitened-panel path).
// With W = cov(u) + G = L_u Lambda L_u':  log det Lambda = log det W - log det cov(u)  and  c'c = g' W^-1 g = |L_W^-1 g|^2,
// so the value needs two Cholesky factorisations and one triangular solve — ONE M x M matrix in shared memory at a time
// (M up to ~160: the EEG shape M = N = 156 of examples/eeg.jl:212-232 fits).
template <int KIND>
__global__ void __launch_bounds__(256)
ss_tail_kernel(const double* __restrict__ Z, int DX, int M, int Mp, int64_t N, const SmallCand* __restrict__ cand,
               const double* __restrict__ Gp, int nsplit, const double* __restrict__ gpart, int nch, const double* __restrict__ sums,
               double cond_thr, double* __restrict__ out) {
  extern __shared__ double sm[];
  const int ld = M | 1;                                  // odd leading dimension: conflict-free columns and rows
  double* S = sm; double* v = S + (size_t)ld * M;
  __shared__ int info;
  __shared__ double red[32];
  const int cd = blockIdx.x;
  const SmallCand c = cand[cd];
  const double inv_l = sqrt(c.inv_l2);
  if (threadIdx.x == 0) info = 0;
  auto fill_cov_u = [&](bool add_G) {
    for (int e = threadIdx.x; e < M * M; e += blockDim.x) {
      const int i = e % M, j = e / M;
      if (i < j) continue;                               // lower triangle only
      double d2 = 0.0;
      for (int d = 0; d < DX; d++) { const double df = Z[(int64_t)i * DX + d] - Z[(int64_t)j * DX + d]; d2 = fma(df, df, d2); }
      double val = c.out_s * ss_kappa_r<KIND>(sqrt(d2) * inv_l) + (i == j ? c.noise : 0.0);
      if (add_G) {
        double g = 0.0;
        const double* gp = Gp + (int64_t)cd * nsplit * Mp * Mp + i + (int64_t)j * Mp;
        int q = 0;
        for (; q + 7 < nsplit; q += 8) {        // independent loads first, then the additions in fixed order
          double t8[8];
#pragma unroll
          for (int u = 0; u < 8; u++) t8[u] = __ldg(gp + (int64_t)(q + u) * Mp * Mp);
#pragma unroll
          for (int u = 0; u < 8; u++) g += t8[u];
        }
        for (; q < nsplit; q++) g += __ldg(gp + (int64_t)q * Mp * Mp);
        val += g;
      }
      S[i + j * ld] = val;
    }
  };
  fill_cov_u(false);
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    double g = 0.0;
    int c2 = 0;
    for (; c2 + 7 < nch; c2 += 8) {
      double t8[8];
#pragma unroll
      for (int u = 0; u < 8; u++) t8[u] = __ldg(gpart + ((int64_t)cd * nch + c2 + u) * Mp + i);
#pragma unroll
      for (int u = 0; u < 8; u++) g += t8[u];
    }
    for (; c2 < nch; c2++) g += __ldg(gpart + ((int64_t)cd * nch + c2) * Mp + i);
    v[i] = g;
  }
  __syncthreads();
  ss_chol_inplace(S, M, ld, &info);
  if (info) { if (threadIdx.x == 0) { out[2 * cd] = 0.0; out[2 * cd + 1] = 1.0; } return; }
  double ldu = 0.0;
  {   // log det cov(u) and its conditioning estimate: (max / min of diag L_u)^2
    double mn = 1e300, mx = 0.0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) { const double d = S[i + i * ld]; mn = fmin(mn, d); mx = fmax(mx, d); ldu += log(d); }
    for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mn; red[8 + (threadIdx.x >> 5)] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 1; q < (int)(blockDim.x >> 5); q++) { mn = fmin(mn, red[q]); mx = fmax(mx, red[8 + q]); }
      const double rr = mx / mn;
      if (rr * rr > cond_thr) info = -1;
    }
    __syncthreads();
    if (info) { if (threadIdx.x == 0) { out[2 * cd] = 0.0; out[2 * cd + 1] = 3.0; } return; }
    ldu = block_sum(ldu, red);
  }
  __syncthreads();
  fill_cov_u(true);                                       // W = cov(u) + G
  __syncthreads();
  ss_chol_inplace(S, M, ld, &info);
  if (info) { if (threadIdx.x == 0) { out[2 * cd] = 0.0; out[2 * cd + 1] = 2.0; } return; }
  ss_trsv_inplace(S, M, ld, v);                           // L_W^-1 g
  double ldw = 0.0, cc = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) { ldw += log(S[i + i * ld]); cc = fma(v[i], v[i], cc); }
  ldw = block_sum(ldw, red); cc = block_sum(cc, red);
  if (threadIdx.x == 0) {
    const double sum_logS = sums[2 * cd], sum_a2 = sums[2 * cd + 1];
    out[2 * cd] = -0.5 * ((double)N * LOG2PI_SS + sum_logS + 2.0 * (ldw - ldu) + sum_a2 - cc);      // dtc.jl:122-125
    out[2 * cd + 1] = 0.0;
  }
}

struct SmallPlan { int D, DX, M, Mp, Lc, nch, nsplit, slabs_per_split, ncand; int64_t N, Ns; double cond_thr; };
struct SmallBufs { const SmallCand* cand; const double *table, *alpha, *sums; double *beta, *Gp, *gpart, *state, *psi, *out; };

template <int KIND, int D>
int ss_run_kd(gpar_ctx* ctx, const SmallPlan& p, const SmallBufs& b) {
  constexpr int TS = D * D + 2 * D + 1;
  const double* X = ctx->X.as<double>(); const double* Z = ctx->Z.as<double>();
  const dim3 gwalk(p.nch, p.ncand);
  const size_t sm1 = ((size_t)2 * 32 * TS + (size_t)2 * 32 * p.DX + 64 + (size_t)32 * (p.Mp + 1)) * sizeof(double);
  const size_t sm0 = sm1;
  {   // the chunk products only need the table: on the side stream, underneath the first walk
    cudaStream_t main_stream = ctx->stream;
    CU(cudaEventRecord(ctx->ev_fork, main_stream));
    CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    ctx->stream = ctx->stream2;
    const int rc = [&]() -> int { LAUNCH(ctx, ss_chunk_product_kernel<D>, gwalk, 32, 0, b.table, p.N, p.Lc, p.nch, b.psi); return GPAR_OK; }();
    cudaEventRecord(ctx->ev_side, ctx->stream2);
    ctx->stream = main_stream;
    CHK(rc);
  }
  CU(cudaFuncSetAttribute((ss_walk_kernel<KIND, D, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0));
  CU(cudaFuncSetAttribute((ss_walk_kernel<KIND, D, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  LAUNCH(ctx, (ss_walk_kernel<KIND, D, false>), gwalk, p.Mp, sm0, X, Z, p.DX, p.N, p.M, p.Mp, p.Ns, p.Lc, p.nch, b.cand, b.table, b.alpha,
         b.state, b.beta, b.gpart);
  CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
  if (p.nch <= 32) LAUNCH(ctx, ss_carry_seq_kernel<D>, dim3((p.Mp + 127) / 128, p.ncand), 128, 0, b.psi, b.state, p.nch, p.Mp);
  else LAUNCH(ctx, ss_carry_kernel<D>, dim3((p.Mp + 7) / 8, p.ncand), 256, 0, b.psi, b.state, p.nch, p.Mp);
  LAUNCH(ctx, (ss_walk_kernel<KIND, D, true>), gwalk, p.Mp, sm1, X, Z, p.DX, p.N, p.M, p.Mp, p.Ns, p.Lc, p.nch, b.cand, b.table, b.alpha,
         b.state, b.beta, b.gpart);
  const int T = (p.Mp + SS_ST - 1) / SS_ST;
  LAUNCH(ctx, ss_syrk_kernel, dim3(T * (T + 1) / 2, p.nsplit, p.ncand), 256, 0, b.beta, p.Mp, p.Ns, p.slabs_per_split, p.nsplit, b.Gp);
  const int ld = p.M | 1;
  const size_t smem = ((size_t)ld * p.M + p.M + 8) * sizeof(double);
  CU(cudaFuncSetAttribute(ss_tail_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  LAUNCH(ctx, ss_tail_kernel<KIND>, p.ncand, 256, smem, Z, p.DX, p.M, p.Mp, p.N, b.cand, b.Gp, p.nsplit, b.gpart, p.nch, b.sums, p.cond_thr, b.out);
  return GPAR_OK;
}
template <int KIND>
int ss_run_kind(gpar_ctx* ctx, const SmallPlan& p, const SmallBufs& b) {
  switch (p.D) {
    case 1: return ss_run_kd<KIND, 1>(ctx, p, b);
    case 2: return ss_run_kd<KIND, 2>(ctx, p, b);
    default: return ss_run_kd<KIND, 3>(ctx, p, b);
  }
}

}  // namespace

// Is the fused small-problem path applicable to the resident problem?  (M x M matrices of the tail in shared memory.)
bool scaled_small_applicable(const gpar_ctx* ctx) {
  const int64_t M = ctx->M, N = ctx->N;
  const size_t smem = ((size_t)(M | 1) * M + M + 8) * sizeof(double);
  return M >= 1 && M <= 192 && smem <= (size_t)210 * 1024 && N >= 64 && N <= ((int64_t)1 << 18) && ctx->D >= 1 && ctx->D <= 8;
}

// vals[c], codes[c] (0 ok; GPAR_ERR_NOT_POSDEF; -1: hand this candidate to the whitened-panel path) for the candidates
// thetas (5 x ncand, raw parameters).  Everything on ctx->stream; one host synchronisation at the end.
int scaled_small_batch(gpar_ctx* ctx, int k_time, int k_out, const double* thetas, int ncand, double* vals, int* codes) {
  const int D = k_time == GPAR_MATERN12 ? 1 : (k_time == GPAR_MATERN32 ? 2 : (k_time == GPAR_MATERN52 ? 3 : 0));
  if (D == 0) return gpar_fail(ctx, GPAR_ERR_INVALID, "time kernel code %d has no state-space form (use Matern12/32/52)", k_time);
  if (k_out < GPAR_EQ || k_out > GPAR_MATERN52) return gpar_fail(ctx, GPAR_ERR_INVALID, "unknown output kernel code %d", k_out);
  const int64_t N = ctx->N; const int M = (int)ctx->M, DX = ctx->D;
  const int Mp = (M + SS_TILE - 1) / SS_TILE * SS_TILE;
  const int64_t Ns = (N + SS_TILE - 1) / SS_TILE * SS_TILE;
  const int TS = D * D + 2 * D + 1;
  double cond_thr = 1e5;
  if (const char* e = getenv("GPAR_ROBUST_COND")) cond_thr = atof(e);
  // whitening chunks: a multiple of 32 steps, enough blocks (of Mp threads) to fill the device
  const int want_chunks = std::max(1, (int)((int64_t)ctx->num_sms * 16 * 32 / ((int64_t)Mp * std::max(1, std::min(ncand, 64)))));
  int Lc = (int)(((N + want_chunks - 1) / want_chunks + 31) / 32 * 32);
  Lc = std::max(32, std::min(Lc, 2048));
  if (const char* e = getenv("GPAR_SS_LC")) { int v = atoi(e); if (v >= 32 && v <= 4096 && v % 32 == 0) Lc = v; }      // tuning knob
  const int nch = (int)((N + Lc - 1) / Lc);
  const int T = (Mp + 63) / 64, pairs = T * (T + 1) / 2;
  const int64_t nslab = Ns / SS_TILE;
  // SYRK splits along N: whole waves of the 2 resident CTAs per SM (a partly filled last wave costs a full one), few
  // enough that the tail's fixed-order sum of the partial G stays cheap
  int nsplit = 1;
  {
    const int64_t resident = (int64_t)ctx->num_sms * 2, base_blocks = (int64_t)pairs * std::min(ncand, 64);
    double best = 1e300;
    for (int ns = 1; ns <= 32 && ns <= nslab; ns++) {
      const int64_t waves = (base_blocks * ns + resident - 1) / resident;
      const double cost = (double)waves / ns + 0.004 * ns;
      if (cost < best - 1e-12) { best = cost; nsplit = ns; }
    }
  }
  const int slabs_per_split = (int)((nslab + nsplit - 1) / nsplit);
  nsplit = (int)((nslab + slabs_per_split - 1) / slabs_per_split);
  // candidates per pass: bounded scratch (table + alpha + beta + partial G + chunk states per candidate)
  const size_t per_cand = ((size_t)N * (TS + 1) + (size_t)Mp * Ns + (size_t)nsplit * Mp * Mp + (size_t)nch * Mp * (D + 1) + (size_t)nch * D * D + 8) * sizeof(double)
                          + sizeof(SmallCand);
  const int group = (int)std::max<size_t>(1, std::min<size_t>(std::min(ncand, 64), ((size_t)1024 << 20) / per_cand));
  CU(ctx->panelK.reserve(per_cand * group + 256));
  char* base = ctx->panelK.as<char>();
  double* table = reinterpret_cast<double*>(base);
  double* alpha = table + (size_t)group * N * TS;
  double* beta = alpha + (size_t)group * N;
  double* Gp = beta + (size_t)group * Mp * Ns;
  double* gpart = Gp + (size_t)group * nsplit * Mp * Mp;
  double* state = gpart + (size_t)group * nch * Mp;
  double* psi = state + (size_t)group * nch * D * Mp;
  double* sums = psi + (size_t)group * nch * D * D;
  double* out = sums + (size_t)2 * group;
  SmallCand* cand = reinterpret_cast<SmallCand*>(out + (size_t)2 * group);
  // pinned staging at fixed addresses: [l | s | noise] (3 x group), candidates, results — the launch sequence of a pass is
  // captured into a CUDA graph the second time a (shape, pointers) key is seen and replayed afterwards: one graph launch
  // instead of ~14 kernel launches and copies on the host's latency path (GPAR_SMALL_GRAPH=0: plain launches)
  const size_t pin_need = ((size_t)3 * group + 2 * (size_t)group) * sizeof(double) + (size_t)group * sizeof(SmallCand) + 64;
  if (ctx->small_pin_cap < pin_need) {
    if (ctx->small_pin) cudaFreeHost(ctx->small_pin);
    ctx->small_pin = nullptr; ctx->small_pin_cap = 0;
    CU(cudaMallocHost(&ctx->small_pin, pin_need));
    ctx->small_pin_cap = pin_need;
    ctx->sgraph.have_key = false;
  }
  double* pin_par = reinterpret_cast<double*>(ctx->small_pin);
  double* pin_out = pin_par + (size_t)3 * group;
  SmallCand* pin_cand = reinterpret_cast<SmallCand*>(pin_out + (size_t)2 * group);
  bool use_graph = true;
  if (const char* e = getenv("GPAR_SMALL_GRAPH")) use_graph = atoi(e) != 0;
  std::vector<double> hl(group), hs(group), hn(group);
  for (int c0 = 0; c0 < ncand; c0 += group) {
    const int nb = std::min(group, ncand - c0);
    for (int c = 0; c < nb; c++) {
      double pv[5];
      for (int i = 0; i < 5; i++) pv[i] = exp(thetas[5 * (size_t)(c0 + c) + i]) + 1e-3;         // unpack_gpar (util.jl:45-55)
      hl[c] = pv[0]; hs[c] = pv[1] * pv[1]; hn[c] = pv[4] * pv[4];
      pin_cand[c] = SmallCand{1.0 / (pv[2] * pv[2]), pv[3] * pv[3], pv[4] * pv[4]};
    }
    SmallPlan pl{D, DX, M, Mp, Lc, nch, nsplit, slabs_per_split, nb, N, Ns, cond_thr};
    SmallBufs bf{cand, table, alpha, sums, beta, Gp, gpart, state, psi, out};
    auto issue = [&]() -> int {       // everything of one pass, on ctx->stream (and the side stream, forked and joined)
      CU(cudaMemcpyAsync(cand, pin_cand, nb * sizeof(SmallCand), cudaMemcpyHostToDevice, ctx->stream));
      ctx->y_broadcast = true; ctx->param_staging = pin_par;
      const int rc = lgssm_run(ctx, k_time, hl.data(), hs.data(), hn.data(), nb, nb, N, ctx->t.as<double>(), ctx->y.as<double>(), nullptr,
                               alpha, nullptr, nullptr, nullptr, table, sums);
      ctx->y_broadcast = false; ctx->param_staging = nullptr;
      CHK(rc);
      switch (k_out) {
        case GPAR_EQ: CHK(ss_run_kind<GPAR_EQ>(ctx, pl, bf)); break;
        case GPAR_MATERN12: CHK(ss_run_kind<GPAR_MATERN12>(ctx, pl, bf)); break;
        case GPAR_MATERN32: CHK(ss_run_kind<GPAR_MATERN32>(ctx, pl, bf)); break;
        default: CHK(ss_run_kind<GPAR_MATERN52>(ctx, pl, bf)); break;
      }
      CU(cudaMemcpyAsync(pin_out, out, 2 * (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
      return GPAR_OK;
    };
    unsigned long long key[20] = {(unsigned long long)k_time, (unsigned long long)k_out, (unsigned long long)N, (unsigned long long)M,
                                  (unsigned long long)DX, (unsigned long long)nb, (unsigned long long)Lc, (unsigned long long)nsplit,
                                  (unsigned long long)(uintptr_t)ctx->X.p, (unsigned long long)(uintptr_t)ctx->Z.p, (unsigned long long)(uintptr_t)ctx->t.p,
                                  (unsigned long long)(uintptr_t)ctx->y.p, (unsigned long long)(uintptr_t)ctx->panelK.p, (unsigned long long)(uintptr_t)ctx->kal_a.p,
                                  (unsigned long long)(uintptr_t)ctx->kal_c.p, (unsigned long long)(uintptr_t)ctx->small_pin, 0ull, 0ull, 0ull, 0ull};
    { double rd = ctx->t_reg_dt; memcpy(&key[16], &rd, sizeof(double)); memcpy(&key[17], &cond_thr, sizeof(double)); key[18] = (unsigned long long)group; }
    gpar_ctx::SmallGraph& sg = ctx->sgraph;
    const bool same = sg.have_key && memcmp(sg.key, key, sizeof(key)) == 0;
    bool done = false;
    if (use_graph && same && sg.exec) {          // replay: the staging has been rewritten above; (l, s, noise) are packed here
      for (int i = 0; i < nb; i++) { pin_par[i] = hl[i]; pin_par[nb + i] = hs[i]; pin_par[2 * (size_t)nb + i] = hn[i]; }
      CU(cudaGraphLaunch(sg.exec, ctx->stream));
      ctx->launches += 1;
      done = true;
    } else if (use_graph && same && !sg.failed && sg.warm >= 1) {          // second sight of this key: capture
      cudaGraph_t graph = nullptr;
      bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        const int rc = issue();
        const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &graph);
        ok = rc == GPAR_OK && ee == cudaSuccess && graph != nullptr;
      }
      if (ok) {
        if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
        ok = cudaGraphInstantiate(&sg.exec, graph, 0) == cudaSuccess;
      }
      if (graph) cudaGraphDestroy(graph);
      if (ok) { CU(cudaGraphLaunch(sg.exec, ctx->stream)); ctx->launches += 1; done = true; }
      else { sg.failed = true; sg.exec = nullptr; (void)cudaGetLastError(); }
    }
    if (!done) {
      if (!same) {
        if (sg.exec) { cudaGraphExecDestroy(sg.exec); sg.exec = nullptr; }
        memcpy(sg.key, key, sizeof(key)); sg.have_key = true; sg.warm = 0; sg.failed = false;
      }
      CHK(issue());
      sg.warm++;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < nb; c++) {
      const int code = (int)pin_out[2 * c + 1];
      vals[c0 + c] = code == 0 ? pin_out[2 * c] : std::numeric_limits<double>::quiet_NaN();
      codes[c0 + c] = code == 0 ? 0 : (code == 3 ? -1 : GPAR_ERR_NOT_POSDEF);
    }
  }
  return GPAR_OK;
}
