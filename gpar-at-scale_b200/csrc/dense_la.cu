// dense_la.cu — the M x M (and exact-GP N x N) dense linear algebra of the path, hand-written for sm_100a:
// Cholesky factorisation, triangular inverse / solves and small GEMMs on column-major FP64 matrices.
//
// Replaces the LAPACK / BLAS calls the reference makes through Julia's LinearAlgebra on these small matrices —
// `cholesky(Symmetric(cov(u)))`, `cholesky(Symmetric(A * A' + I))`, the `\` triangular solves, `logdet`, `inv(D)`
// (src/gp/dtc.jl:119-125; src/gp/gpar_scaled_inference.jl:159,179,187-192) and Stheno's dense `logpdf` / posterior
// (src/gp/optimized.jl:34,152,94,236) — which round 1 delegated to cuSOLVER / cuBLAS.  They are < 0.3 % of the
// flops of a BASELINE-sized evaluation but all of the exact-GP path and the latency floor of the reference-sized
// problems (M = 50 ... 156), so they are latency-oriented:
//   * dla_gemm: C = alpha op(A) op(B) + beta C on 64 x 64 tiles, mma.sync.m8n8k4.f64 (DMMA), operands staged in the
//     fragment-native shared-memory layout; triangular operands restrict the k range and are masked at load time
//     (so the untouched triangle of an in-place Cholesky factor may hold anything); batched through blockIdx.z.
//   * dla_potrf: right-looking blocked Cholesky, 64 x 64 diagonal blocks factorised (and inverted) by one CTA in
//     shared memory, panel = product with the inverted diagonal block, trailing update = lower-tile GEMM.
//   * dla_trtri: full inverse of a lower-triangular factor by recursive doubling over the inverted diagonal blocks
//     (two batched GEMMs per level); the M x M "solves" of the tails are then single GEMMs.
//   * dla_trsm_left: blocked substitution (inverted 64 x 64 diagonal blocks + GEMM updates) for many right-hand sides
//     where backward stability matters (exact GP); dla_trsv: one-CTA blocked substitution for single vectors (c, w).
// Everything is enqueued on the context's stream; failure of a factorisation is reported through a device flag
// (the 1-based index of the first non-positive pivot, as LAPACK's info / Julia's PosDefException.info).
#include "common.cuh"
#include <algorithm>
#include "dmma_pipe.cuh"

namespace {

using dmma::dmma884;

constexpr int DT = 64;        // output tile
constexpr int DK = 16;        // k-block per shared-memory stage

struct GemmP {
  const double* A; const double* B; double* C;
  int m, n, k, lda, ldb, ldc;
  double alpha, beta;
  long long sA, sB, sC;       // batch strides (doubles)
  int flags;
};

// op(A) is m x k, op(B) is k x n.  Shared layout [k/4][row or col][k%4]: a DMMA fragment is 32 consecutive doubles.
// Operands move global -> shared with 8-byte cp.async (zero-filled where masked: outside the matrix or in the ignored
// triangle of a triangular operand) through a 3-stage ring, two k-blocks ahead of the DMMAs; one barrier per k-block.
constexpr int DST = 3;
__device__ __forceinline__ void cp_async8_zfill(double* dst, const double* src, bool valid) {
  const unsigned sz = valid ? 8u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(sz) : "memory");
}
template <bool TA, bool TB>
__global__ void __launch_bounds__(128)
dla_gemm_kernel(const GemmP p) {
  __shared__ double As[DST][DK / 4 * DT * 4];
  __shared__ double Bs[DST][DK / 4 * DT * 4];
  const int i0 = blockIdx.x * DT, j0 = blockIdx.y * DT;
  if ((p.flags & DLA_LOWER_TILES) && j0 > i0) return;
  const double* A = p.A + (long long)blockIdx.z * p.sA;
  const double* B = p.B + (long long)blockIdx.z * p.sB;
  double* C = p.C + (long long)blockIdx.z * p.sC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp >> 1, wc = warp & 1;
  // k range that can contribute to this tile
  int kbeg = 0, kend = p.k;
  if (p.flags & DLA_A_LOWER) kend = min(kend, i0 + DT);          // op(A)[i, q] = 0 for q > i
  if (p.flags & DLA_B_UPPER) kend = min(kend, j0 + DT);          // op(B)[q, j] = 0 for q > j
  if (p.flags & DLA_A_UPPER) kbeg = max(kbeg, i0);               // op(A)[i, q] = 0 for q < i
  if (p.flags & DLA_B_LOWER) kbeg = max(kbeg, j0);               // op(B)[q, j] = 0 for q < j
  kbeg = kbeg / DK * DK;
  const int nst = kend > kbeg ? (kend - kbeg + DK - 1) / DK : 0;
  // this thread's 8 + 8 elements of a k-block: position inside the tile, shared-memory slot, global pointer at k = 0
  int agi[8], aq[8], bgj[8], bq[8], aslot[8], bslot[8];
  const double* ap[8]; const double* bp[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int e = tid + 128 * r;
    int ai, bj;
    if (TA) { aq[r] = e & 15; ai = e >> 4; } else { ai = e & 63; aq[r] = e >> 6; }
    if (TB) { bj = e & 63; bq[r] = e >> 6; } else { bq[r] = e & 15; bj = e >> 4; }
    agi[r] = i0 + ai; bgj[r] = j0 + bj;
    aslot[r] = (aq[r] >> 2) * (DT * 4) + ai * 4 + (aq[r] & 3);
    bslot[r] = (bq[r] >> 2) * (DT * 4) + bj * 4 + (bq[r] & 3);
    ap[r] = TA ? A + (long long)agi[r] * p.lda + aq[r] : A + (long long)agi[r] + (long long)aq[r] * p.lda;
    bp[r] = TB ? B + (long long)bgj[r] + (long long)bq[r] * p.ldb : B + (long long)bgj[r] * p.ldb + bq[r];
  }
  const long long astep = TA ? 1 : p.lda, bstep = TB ? p.ldb : 1;      // pointer advance per unit of k
  auto issue = [&](int st) {
    const int q0 = kbeg + st * DK, buf = st % DST;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int gq = q0 + aq[r];
      bool keep = agi[r] < p.m && gq < p.k;
      if ((p.flags & DLA_A_LOWER) && gq > agi[r]) keep = false;
      if ((p.flags & DLA_A_UPPER) && gq < agi[r]) keep = false;
      cp_async8_zfill(&As[buf][aslot[r]], keep ? ap[r] + (long long)q0 * astep : A, keep);
      const int hq = q0 + bq[r];
      bool keepb = bgj[r] < p.n && hq < p.k;
      if ((p.flags & DLA_B_UPPER) && hq > bgj[r]) keepb = false;
      if ((p.flags & DLA_B_LOWER) && hq < bgj[r]) keepb = false;
      cp_async8_zfill(&Bs[buf][bslot[r]], keepb ? bp[r] + (long long)q0 * bstep : B, keepb);
    }
  };
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll
  for (int st = 0; st < DST - 1; st++) { if (st < nst) issue(st); asm volatile("cp.async.commit_group;" ::: "memory"); }
  for (int st = 0; st < nst; st++) {
    asm volatile("cp.async.wait_group %0;" ::"n"(DST - 2) : "memory");
    __syncthreads();                 // k-block st has landed for every thread; buffer (st - 1) % DST is free again
    if (st + DST - 1 < nst) issue(st + DST - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* as = As[st % DST]; const double* bs = Bs[st % DST];
#pragma unroll
    for (int k4 = 0; k4 < DK / 4; k4++) {
      double af[4], bf[4];
#pragma unroll
      for (int a = 0; a < 4; a++) af[a] = as[k4 * (DT * 4) + (wr * 32 + a * 8) * 4 + lane];
#pragma unroll
      for (int b = 0; b < 4; b++) bf[b] = bs[k4 * (DT * 4) + (wc * 32 + b * 8) * 4 + lane];
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int gi = i0 + wr * 32 + a * 8 + (lane >> 2), gj = j0 + wc * 32 + b * 8 + (lane & 3) * 2 + h;
        if (gi < p.m && gj < p.n) {
          double* c = C + (long long)gi + (long long)gj * p.ldc;
          const double v = p.alpha * acc[a][b][h];
          *c = (p.beta == 0.0) ? v : fma(p.beta, *c, v);
        }
      }
}

// Y = L^-1 for the nb x nb lower-triangular block in S (row-major, pitch 65) into Ys: thread c solves L y = e_c by
// forward substitution with four partial sums (the dependent chain is the latency that matters here).
__device__ __forceinline__ void tri_inverse_column(const double (*S)[DT + 1], double (*Ys)[DT + 1], int nb, int c) {
  for (int i = 0; i < c; i++) Ys[i][c] = 0.0;
  for (int i = c; i < nb; i++) {
    double v0 = (i == c) ? 1.0 : 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    int q = c;
    for (; q + 3 < i; q += 4) {
      v0 = fma(-S[i][q], Ys[q][c], v0); v1 = fma(-S[i][q + 1], Ys[q + 1][c], v1);
      v2 = fma(-S[i][q + 2], Ys[q + 2][c], v2); v3 = fma(-S[i][q + 3], Ys[q + 3][c], v3);
    }
    for (; q < i; q++) v0 = fma(-S[i][q], Ys[q][c], v0);
    Ys[i][c] = ((v0 + v1) + (v2 + v3)) / S[i][i];
  }
}

// In-place Cholesky of the nb x nb (nb <= 64) diagonal block at A (lower triangle; the upper one is not touched) by one
// CTA of 256 threads in shared memory, and Y = L^-1 (64 x 64 column-major scratch, zeros above the diagonal) for the panel
// product.  Both loops are right-looking with DEFERRED scaling, so a column costs one barrier and one sweep of
// independent updates: trailing(i, c) -= a_ij a_cj / d_j with the unscaled column j (l_ij = a_ij / sqrt(d_j) is applied
// once at the end); the inverse is the Gauss-Jordan elimination of [L | I] with the row scalings deferred likewise.
// info: set to j0 + j + 1 at the first non-positive pivot (if still 0); the block is then filled with NaN.
__global__ void __launch_bounds__(256)
dla_chol_diag_kernel(double* Ab, int lda, int nb, int j0, double* Yb, int* infob, long long sA, long long sY) {
  extern __shared__ double dla_sm[];            // 2 x 64 x 65 doubles (opt-in dynamic shared memory)
  double (*S)[DT + 1] = reinterpret_cast<double (*)[DT + 1]>(dla_sm);
  double (*Ys)[DT + 1] = reinterpret_cast<double (*)[DT + 1]>(dla_sm + DT * (DT + 1));
  double* A = Ab + (long long)blockIdx.x * sA;
  double* Y = Yb ? Yb + (long long)blockIdx.x * sY : nullptr;
  int* info = infob + blockIdx.x;
  const int tid = threadIdx.x;
  for (int e = tid; e < DT * DT; e += 256) {
    const int i = e & 63, c = e >> 6;
    S[i][c] = (i < nb && c < nb && i >= c) ? A[(long long)i + (long long)c * lda] : 0.0;
    Ys[i][c] = (i == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  // threads as a 16 x 16 grid over (row, column) residues of the trailing block: no integer division in the column loop,
  // the column's entries a thread needs are read once into registers
  const int ti = tid & 15, tc = tid >> 4;
  int bad = 0;
  for (int j = 0; j < nb; j++) {
    const double d = S[j][j];
    if (!(d > 0.0)) { bad = j + 1; break; }           // uniform: every thread reads the same pivot
    const double inv = 1.0 / d;
    double li[4], lc[4];
#pragma unroll
    for (int a = 0; a < 4; a++) { const int i = j + 1 + ti + 16 * a; li[a] = (i < nb) ? S[i][j] * inv : 0.0; }
#pragma unroll
    for (int b = 0; b < 4; b++) { const int c = j + 1 + tc + 16 * b; lc[b] = (c < nb) ? S[c][j] : 0.0; }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int i = j + 1 + ti + 16 * a, c = j + 1 + tc + 16 * b;
        if (i < nb && c <= i) S[i][c] = fma(-li[a], lc[b], S[i][c]);
      }
    __syncthreads();
  }
  if (bad) {
    if (tid == 0 && *info == 0) *info = j0 + bad;
    for (int e = tid; e < nb * nb; e += 256) { const int i = e % nb, c = e / nb; if (i >= c) A[(long long)i + (long long)c * lda] = nan(""); }
    if (Y) for (int e = tid; e < DT * DT; e += 256) Y[e] = nan("");
    return;
  }
  // l_ij = a_ij / sqrt(d_j)
  for (int e = tid; e < nb * nb; e += 256) {
    const int i = e % nb, c = e / nb;
    if (i >= c) { const double v = (i == c) ? sqrt(S[c][c]) : S[i][c] * rsqrt(S[c][c]); A[(long long)i + (long long)c * lda] = v; }
  }
  if (!Y) return;
  __syncthreads();
  for (int e = tid; e < nb * nb; e += 256) { const int i = e % nb, c = e / nb; if (i > c) S[i][c] *= rsqrt(S[c][c]); }
  __syncthreads();
  for (int e = tid; e < nb; e += 256) S[e][e] = sqrt(S[e][e]);
  __syncthreads();
  // Y = L^-1: for column j of L, rows i > j: Y[i][c] -= (L[i][j] / L[j][j]) Y~[j][c] for c <= j, with Y~[j] row j before its scaling
  for (int j = 0; j < nb; j++) {
    const double inv = 1.0 / S[j][j];
    double li[4], yc[4];
#pragma unroll
    for (int a = 0; a < 4; a++) { const int i = j + 1 + ti + 16 * a; li[a] = (i < nb) ? S[i][j] * inv : 0.0; }
#pragma unroll
    for (int b = 0; b < 4; b++) { const int c = tc + 16 * b; yc[b] = (c <= j) ? Ys[j][c] : 0.0; }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int i = j + 1 + ti + 16 * a, c = tc + 16 * b;
        if (i < nb && c <= j) Ys[i][c] = fma(-li[a], yc[b], Ys[i][c]);
      }
    __syncthreads();
  }
  for (int e = tid; e < DT * DT; e += 256) {
    const int i = e & 63, c = e >> 6;
    Y[e] = (i < nb && c < nb) ? ((i >= c) ? Ys[i][c] / S[i][i] : 0.0) : ((i == c) ? 1.0 : 0.0);
  }
}

// Inverses of the 64 x 64 diagonal blocks of the lower-triangular L (n x n): block blockIdx.x -> out + blockIdx.x * sblk
// (leading dimension ldo): sblk = 64 (1 + ldo) writes them in place on the diagonal of an n x n matrix (dla_trtri),
// sblk = 64 * 64 with ldo = 64 into a strip (dla_trsm_left, dla_trsv).  blockIdx.y: batch member.
__global__ void __launch_bounds__(DT)
dla_tri_inv_diag_kernel(const double* __restrict__ Lb, int ldl, int n, double* __restrict__ Ob, int ldo, long long sblk, long long sL, long long sO,
                        int pad_identity) {
  extern __shared__ double dla_sm[];
  double (*S)[DT + 1] = reinterpret_cast<double (*)[DT + 1]>(dla_sm);
  double (*Ys)[DT + 1] = reinterpret_cast<double (*)[DT + 1]>(dla_sm + DT * (DT + 1));
  const double* L = Lb + (long long)blockIdx.y * sL;
  double* O = Ob + (long long)blockIdx.y * sO + (long long)blockIdx.x * sblk;
  const int k0 = blockIdx.x * DT, nb = min(DT, n - k0), i = threadIdx.x;
  for (int c = 0; c < nb; c++) S[i][c] = (i < nb && i >= c) ? L[(long long)(k0 + i) + (long long)(k0 + c) * ldl] : 0.0;
  __syncthreads();
  if (i < nb) tri_inverse_column(S, Ys, nb, i);
  __syncthreads();
  if (pad_identity) { for (int c = 0; c < DT; c++) O[(long long)i + (long long)c * ldo] = (i < nb && c < nb) ? Ys[i][c] : ((i == c) ? 1.0 : 0.0); }
  else { for (int c = 0; c < nb; c++) if (i < nb) O[(long long)i + (long long)c * ldo] = Ys[i][c]; }
}

// x <- L^-1 (scale x) (TRANS = false) or L^-T (scale x) (TRANS = true), L lower triangular, as a dataflow pipeline over the
// 64-row blocks: CTA p owns the p-th block in solution order, accumulates (its rows) x (every solved block) as those
// blocks are published through `flags`, then solves its diagonal block by substitution (one warp, two rows per lane,
// shuffle broadcast of each solved entry: backward stable also for a poorly conditioned factor, unlike a product with
// the inverted block) and publishes its own.  The matrix entries of the next block are fetched before its flag is
// awaited, so the critical path is the chain of 64-step substitutions.  All CTAs must be co-resident:
// n <= 64 * (number of SMs).
template <bool TRANS>
__global__ void __launch_bounds__(256)
dla_trsv_pipe_kernel(const double* __restrict__ L, int ldl, int n, double* x, int* flags, double scale) {
  __shared__ double xs[DT];
  __shared__ double red[4][DT];
  __shared__ double Ds[DT][DT + 1];
  __shared__ double rdg[DT];
  const int nblk = (n + DT - 1) / DT;
  const int p = blockIdx.x, b = TRANS ? nblk - 1 - p : p;
  const int k0 = b * DT, nb = min(DT, n - k0);
  const int tid = threadIdx.x, r = tid & 63, g = tid >> 6;
  for (int e = tid; e < nb * nb; e += 256) { const int i = e % nb, c = e / nb; Ds[i][c] = (i >= c) ? L[(long long)(k0 + i) + (long long)(k0 + c) * ldl] : 0.0; }
  if (tid < nb) rdg[tid] = 1.0 / L[(long long)(k0 + tid) + (long long)(k0 + tid) * ldl];
  double acc = 0.0;
  for (int pj = 0; pj < p; pj++) {
    const int j = TRANS ? nblk - 1 - pj : pj, j0 = j * DT;
    double lv[16];
#pragma unroll
    for (int c = 0; c < 16; c++) {
      const int col = j0 + g * 16 + c;
      // element T(k0 + r, col) of the triangular matrix: L(k0 + r, col) or, transposed, L(col, k0 + r)
      lv[c] = (r < nb && col < n) ? (TRANS ? L[(long long)col + (long long)(k0 + r) * ldl] : L[(long long)(k0 + r) + (long long)col * ldl]) : 0.0;
    }
    if (tid == 0) { while (*reinterpret_cast<volatile int*>(flags + j) == 0) { } __threadfence(); }
    __syncthreads();
    if (tid < DT) xs[tid] = (j0 + tid < n) ? __ldcg(x + j0 + tid) : 0.0;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 16; c++) acc = fma(lv[c], xs[g * 16 + c], acc);
    __syncthreads();
  }
  red[g][r] = acc;
  __syncthreads();
  if (tid < 32) {
    const int lane = tid, ra = lane, rb = lane + 32;
    double va = (ra < nb) ? scale * x[k0 + ra] - ((red[0][ra] + red[1][ra]) + (red[2][ra] + red[3][ra])) : 0.0;
    double vb = (rb < nb) ? scale * x[k0 + rb] - ((red[0][rb] + red[1][rb]) + (red[2][rb] + red[3][rb])) : 0.0;
    for (int s = 0; s < nb; s++) {
      const int j = TRANS ? nb - 1 - s : s;
      double xj = (j < 32) ? va : vb;
      xj = __shfl_sync(0xffffffffu, xj, j & 31) * rdg[j];
      if (ra == j) va = xj;
      if (rb == j) vb = xj;
      // remaining rows: below j for the lower solve, above j for the transposed (upper) one; T(row, j) = L(row, j) or L(j, row)
      const bool pa = TRANS ? (ra < j) : (ra > j && ra < nb), pb = TRANS ? (rb < j) : (rb > j && rb < nb);
      if (pa) va = fma(-(TRANS ? Ds[j][ra] : Ds[ra][j]), xj, va);
      if (pb) vb = fma(-(TRANS ? Ds[j][rb] : Ds[rb][j]), xj, vb);
    }
    if (ra < nb) x[k0 + ra] = va;
    if (rb < nb) x[k0 + rb] = vb;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) { __threadfence(); *reinterpret_cast<volatile int*>(flags + b) = 1; }
}

__global__ void dla_dot_kernel(const double* __restrict__ x, const double* __restrict__ y, int n, double* __restrict__ out) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a = fma(x[i], y[i], a);
  const double r = block_sum(a, sh);
  if (threadIdx.x == 0) out[0] = r;
}
__global__ void dla_scal_kernel(double* __restrict__ x, long long n, double a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= a;
}
// A (n x n): mirror the lower triangle into the upper one
__global__ void dla_symmetrize_kernel(double* A, int lda, int n) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n * n) return;
  const int i = (int)(e % n), j = (int)(e / n);
  if (i < j) A[(long long)i + (long long)j * lda] = A[(long long)j + (long long)i * lda];
}

constexpr size_t DLA_DIAG_SMEM = (size_t)2 * DT * (DT + 1) * sizeof(double);
int dla_diag_smem_optin(gpar_ctx* ctx) {
  if (ctx->dla_optin) return GPAR_OK;            // once per context (= per device): the attribute is sticky
  ctx->dla_optin = true;
  CU(cudaFuncSetAttribute(dla_chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DLA_DIAG_SMEM));
  CU(cudaFuncSetAttribute(dla_tri_inv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DLA_DIAG_SMEM));
  return GPAR_OK;
}

}  // namespace

int dla_gemm_batched(gpar_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, long long sA,
                     const double* B, int ldb, long long sB, double beta, double* C, int ldc, long long sC, int batch, int flags) {
  if (m <= 0 || n <= 0 || batch <= 0) return GPAR_OK;
  GemmP p{A, B, C, m, n, k, lda, ldb, ldc, alpha, beta, sA, sB, sC, flags};
  dim3 grid((m + DT - 1) / DT, (n + DT - 1) / DT, batch);
  if (!ta && !tb) LAUNCH(ctx, (dla_gemm_kernel<false, false>), grid, 128, 0, p);
  else if (!ta && tb) LAUNCH(ctx, (dla_gemm_kernel<false, true>), grid, 128, 0, p);
  else if (ta && !tb) LAUNCH(ctx, (dla_gemm_kernel<true, false>), grid, 128, 0, p);
  else LAUNCH(ctx, (dla_gemm_kernel<true, true>), grid, 128, 0, p);
  return GPAR_OK;
}
int dla_gemm(gpar_ctx* ctx, bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double beta, double* C, int ldc, int flags) {
  return dla_gemm_batched(ctx, ta, tb, m, n, k, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1, flags);
}

// scratch of the factorisations: 64 x 64 inverted diagonal block per batch member
static int dla_scratch(gpar_ctx* ctx, size_t doubles, double** out) {
  DevBuf& ws = (ctx->stream == ctx->stream2) ? ctx->dla_ws_side : ctx->dla_ws;      // the side stream factors cov(u) underneath the main stream
  CU(ws.reserve(doubles * sizeof(double)));
  *out = ws.as<double>();
  return GPAR_OK;
}

// A (n x n, lower triangle referenced and overwritten by L; the strictly upper triangle is left as it was), `batch`
// matrices `sA` doubles apart; dinfo[b] = 0 or the 1-based index of the first non-positive pivot.
int dla_potrf_batched(gpar_ctx* ctx, int n, double* A, int lda, long long sA, int batch, int* dinfo) {
  double* Y;
  CHK(dla_scratch(ctx, (size_t)batch * DT * DT, &Y));
  CU(cudaMemsetAsync(dinfo, 0, (size_t)batch * sizeof(int), ctx->stream));
  CHK(dla_diag_smem_optin(ctx));
  for (int k0 = 0; k0 < n; k0 += DT) {
    const int nb = std::min(DT, n - k0), rem = n - k0 - nb;
    double* Akk = A + (size_t)k0 + (size_t)k0 * lda;
    LAUNCH(ctx, dla_chol_diag_kernel, batch, 256, DLA_DIAG_SMEM, Akk, lda, nb, k0, rem > 0 ? Y : (double*)nullptr, dinfo, sA, (long long)DT * DT);
    if (rem > 0) {
      double* A21 = Akk + nb;
      // L21 = A21 L_kk^-T = A21 Y^T: one tile column, so every CTA reads exactly the rows it overwrites (in place)
      CHK(dla_gemm_batched(ctx, false, true, rem, nb, nb, 1.0, A21, lda, sA, Y, DT, (long long)DT * DT, 0.0, A21, lda, sA, batch, DLA_B_UPPER));
      // A22 -= L21 L21^T on the lower tiles
      double* A22 = Akk + nb + (size_t)nb * lda;
      CHK(dla_gemm_batched(ctx, false, true, rem, rem, nb, -1.0, A21, lda, sA, A21, lda, sA, 1.0, A22, lda, sA, batch, DLA_LOWER_TILES));
    }
  }
  return GPAR_OK;
}
int dla_potrf(gpar_ctx* ctx, int n, double* A, int lda, int* dinfo) { return dla_potrf_batched(ctx, n, A, lda, 0, 1, dinfo); }

// V (n x n, ldv) = L^-1 for the lower-triangular L (its upper triangle is ignored); V's upper triangle is zeroed.
// Recursive doubling: with the inverses of the 2^s * 64 diagonal blocks, inv([L11 0; L21 L22]) = [V11 0; -V22 L21 V11, V22].
int dla_trtri(gpar_ctx* ctx, int n, const double* L, int ldl, double* V, int ldv) {
  CU(cudaMemset2DAsync(V, (size_t)ldv * sizeof(double), 0, (size_t)n * sizeof(double), n, ctx->stream));
  CHK(dla_diag_smem_optin(ctx));
  LAUNCH(ctx, dla_tri_inv_diag_kernel, dim3((n + DT - 1) / DT, 1), DT, DLA_DIAG_SMEM, L, ldl, n, V, ldv, (long long)DT * (1 + (long long)ldv), 0LL, 0LL, 0);
  double* T;
  CHK(dla_scratch(ctx, (size_t)n * n / 2 + (size_t)DT * DT, &T));      // per level: pairs * b * b <= n^2 / 4 ... n^2 / 2 with the ragged pair
  for (int b = DT; b < n; b *= 2) {
    const int full = n / (2 * b);                  // pairs whose second block is complete
    const long long sL = (long long)2 * b * (1 + (long long)ldl), sV = (long long)2 * b * (1 + (long long)ldv), sT = (long long)b * b;
    if (full > 0) {
      CHK(dla_gemm_batched(ctx, false, false, b, b, b, 1.0, L + b, ldl, sL, V, ldv, sV, 0.0, T, b, sT, full, DLA_B_LOWER));             // T = L21 V11
      CHK(dla_gemm_batched(ctx, false, false, b, b, b, -1.0, V + b + (size_t)b * ldv, ldv, sV, T, b, sT, 0.0, V + b, ldv, sV, full, DLA_A_LOWER));   // V21 = -V22 T
    }
    const int r0 = full * 2 * b, r1 = r0 + b;      // ragged last pair: second block shorter than b (or absent)
    if (r1 < n) {
      const int b2 = n - r1;
      double* Tr = T + (size_t)full * b * b;
      CHK(dla_gemm(ctx, false, false, b2, b, b, 1.0, L + r1 + (size_t)r0 * ldl, ldl, V + r0 + (size_t)r0 * ldv, ldv, 0.0, Tr, b2, DLA_B_LOWER));
      CHK(dla_gemm(ctx, false, false, b2, b, b2, -1.0, V + r1 + (size_t)r1 * ldv, ldv, Tr, b2, 0.0, V + r1 + (size_t)r0 * ldv, ldv, DLA_A_LOWER));
    }
  }
  return GPAR_OK;
}

// B (n x nrhs) <- L^-1 B (trans = false) or L^-T B (trans = true), L lower triangular: blocked substitution with the
// inverted 64 x 64 diagonal blocks (in Yd: n x 64 scratch strip) and GEMM updates.
int dla_trsm_left(gpar_ctx* ctx, bool trans, int n, int nrhs, const double* L, int ldl, double* B, int ldb) {
  if (n <= 0 || nrhs <= 0) return GPAR_OK;
  const int nblk = (n + DT - 1) / DT;
  double* ws;
  CHK(dla_scratch(ctx, (size_t)nblk * DT * DT, &ws));
  double* Yd = ws;
  // inverted diagonal blocks, block k at Yd + k * 64 * 64 (ld 64, identity-padded)
  CHK(dla_diag_smem_optin(ctx));
  LAUNCH(ctx, dla_tri_inv_diag_kernel, dim3(nblk, 1), DT, DLA_DIAG_SMEM, L, ldl, n, Yd, DT, (long long)DT * DT, 0LL, 0LL, 1);
  for (int q = 0; q < nblk; q++) {
    const int k = trans ? nblk - 1 - q : q;
    const int k0 = k * DT, nb = std::min(DT, n - k0);
    double* Bk = B + k0;
    // Bk <- Y_kk Bk (or Y_kk^T Bk), in place: one tile row, so a CTA reads exactly the columns it overwrites, all of them
    // before its epilogue
    CHK(dla_gemm(ctx, trans, false, nb, nrhs, nb, 1.0, Yd + (size_t)k * DT * DT, DT, Bk, ldb, 0.0, Bk, ldb, trans ? DLA_A_UPPER : DLA_A_LOWER));
    if (!trans) {
      const int rem = n - k0 - nb;
      if (rem > 0) CHK(dla_gemm(ctx, false, false, rem, nrhs, nb, -1.0, L + (size_t)(k0 + nb) + (size_t)k0 * ldl, ldl, Bk, ldb, 1.0, B + k0 + nb, ldb, 0));
    } else if (k0 > 0) {
      // B[0:k0] -= L[k0:k0+nb, 0:k0]^T Bk
      CHK(dla_gemm(ctx, true, false, k0, nrhs, nb, -1.0, L + (size_t)k0, ldl, Bk, ldb, 1.0, B, ldb, 0));
    }
  }
  return GPAR_OK;
}

// x (n) <- L^-1 (scale x) (trans = false) or L^-T (scale x) (trans = true): the dataflow pipeline kernel (one CTA per
// 64-row block, all co-resident).
int dla_trsv(gpar_ctx* ctx, bool trans, int n, const double* L, int ldl, double* x, double scale) {
  if (n <= 0) return GPAR_OK;
  const int nblk = (n + DT - 1) / DT;
  if (nblk > ctx->num_sms) {      // the pipeline needs every CTA resident: fall back to the blocked multi-launch solve
    CHK(dla_trsm_left(ctx, trans, n, 1, L, ldl, x, n));
    return scale == 1.0 ? GPAR_OK : dla_scal(ctx, n, scale, x);
  }
  double* ws;
  CHK(dla_scratch(ctx, (size_t)nblk, &ws));
  int* flags = reinterpret_cast<int*>(ws);
  CU(cudaMemsetAsync(flags, 0, (size_t)nblk * sizeof(int), ctx->stream));
  if (!trans) LAUNCH(ctx, dla_trsv_pipe_kernel<false>, nblk, 256, 0, L, ldl, n, x, flags, scale);
  else LAUNCH(ctx, dla_trsv_pipe_kernel<true>, nblk, 256, 0, L, ldl, n, x, flags, scale);
  return GPAR_OK;
}

int dla_dot(gpar_ctx* ctx, int n, const double* x, const double* y, double* out_dev) {
  LAUNCH(ctx, dla_dot_kernel, 1, 256, 0, x, y, n, out_dev);
  return GPAR_OK;
}
int dla_scal(gpar_ctx* ctx, long long n, double a, double* x) {
  if (n > 0) LAUNCH(ctx, dla_scal_kernel, (unsigned)((n + 255) / 256), 256, 0, x, n, a);
  return GPAR_OK;
}
int dla_symmetrize(gpar_ctx* ctx, int n, double* A, int lda) {
  LAUNCH(ctx, dla_symmetrize_kernel, (unsigned)(((long long)n * n + 255) / 256), 256, 0, A, lda, n);
  return GPAR_OK;
}

namespace {
__global__ void dla_fill_spd_kernel(double* A, int n, unsigned seed) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n * n) return;
  const int i = (int)(e % n), j = (int)(e / n);
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  unsigned h = (unsigned)(lo * 2654435761u) ^ (unsigned)(hi * 40503u) ^ seed; h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
  A[e] = (double)(h & 0xffff) / 65536.0 * 0.01 + (i == j ? 1.0 + 0.01 * n : 0.0);      // symmetric, diagonally dominant
}
}  // namespace

// Diagnostics: device time (ms, CUDA events, best of 3) of the dense routines at order n on a synthetic SPD matrix:
// out[0] potrf, [1] trtri, [2] gemm V G (triangular A), [3] gemm (V G) V' (lower tiles), [4] full gemm, [5] trsv (forward),
// [6] trsv (transposed).  Measurement support (profiles/), not part of the reference's interface.
extern "C" int gpar_dense_bench(gpar_ctx* ctx, int32_t n, double* out7) {
  if (!ctx || !out7 || n < 1) return GPAR_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  const size_t nn = (size_t)n * n;
  CU(ctx->dense.reserve((5 * nn + 2 * (size_t)n) * sizeof(double) + 64));
  CU(ctx->info.reserve(4 * sizeof(int)));
  double* A = ctx->dense.as<double>(); double* Lm = A + nn; double* V = Lm + nn; double* T1 = V + nn; double* T2 = T1 + nn; double* x = T2 + nn;
  LAUNCH(ctx, dla_fill_spd_kernel, (unsigned)((nn + 255) / 256), 256, 0, A, (int)n, 12345u);
  cudaEvent_t e0 = ctx->pev[0], e1 = ctx->pev[3];
  auto timeit = [&](int slot, auto fn) -> int {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      CU(cudaMemcpyAsync(Lm, A, nn * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
      if (slot > 0) CHK(dla_potrf(ctx, n, Lm, n, ctx->info.as<int>()));
      if (slot > 1) CHK(dla_trtri(ctx, n, Lm, n, V, n));
      CU(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), ctx->stream));
      CU(cudaEventRecord(e0, ctx->stream));
      CHK(fn());
      CU(cudaEventRecord(e1, ctx->stream));
      CU(cudaEventSynchronize(e1));
      float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    out7[slot] = best;
    return GPAR_OK;
  };
  CHK(timeit(0, [&]() { return dla_potrf(ctx, n, Lm, n, ctx->info.as<int>()); }));
  CHK(timeit(1, [&]() { return dla_trtri(ctx, n, Lm, n, V, n); }));
  CHK(timeit(2, [&]() { return dla_gemm(ctx, false, false, n, n, n, 1.0, V, n, A, n, 0.0, T1, n, DLA_A_LOWER); }));
  CHK(timeit(3, [&]() { return dla_gemm(ctx, false, true, n, n, n, 1.0, T1, n, V, n, 0.0, T2, n, DLA_B_UPPER | DLA_LOWER_TILES); }));
  CHK(timeit(4, [&]() { return dla_gemm(ctx, false, false, n, n, n, 1.0, A, n, A, n, 0.0, T1, n, 0); }));
  CHK(timeit(5, [&]() { return dla_trsv(ctx, false, n, Lm, n, x, 1.0); }));
  CHK(timeit(6, [&]() { return dla_trsv(ctx, true, n, Lm, n, x, 1.0); }));
  ctx->phase_valid = false;
  return GPAR_OK;
}
