// panel_gemm.cu — FP64 tensor-core (DMMA) product of an operand panel with a small dense matrix:
//
//     out[n, m'] = sum_m  C[m', m] * in[n, m]          (in, out: N x M panels; C: M x M)
//
// This is the N x M x M contraction of the gradients and of the panel whitening — the pieces that were library calls
// (cuBLAS DGEMM / DTRSM on transposed slabs) in round 1:
//   * S = beta P for the scaled-GPAR gradient (R = -beta P + e w'; the gradient of the objective dtc.jl:83-128 that the
//     optimiser loop dtc.jl:58-61 drives) and S = Kuf' A for the pseudo-input gradient (zgrad.cu);
//   * the panel whitening A = L_u^-1 beta' (dtc.jl:119; gpar_scaled_inference.jl:179) as a BLOCKED triangular solve:
//     with the inverses Y_ii of the 128 x 128 diagonal blocks, tile row i is  X_i = Y_ii beta_i - sum_{j<i} (Y_ii L_ij) X_j,
//     i.e. the same product with a block-lower-triangular C (C_ii = Y_ii, C_ij = -Y_ii L_ij), contraction over the tiles
//     j <= i only, in place, one launch per tile row (its inputs X_j, j < i, are the outputs of the earlier launches).
//
// Both operands contract over m, the panel's SLOW index inside a 4-step group ([mt][n/4][m%128][n%4], kuf_panel.cu), so
// the roles of the SYRK are swapped: the small matrix is the A operand (rows m', pre-arranged once into the fragment-
// native layout [m't][m/4][m'%128][m%4] — one 32 KB bulk copy per stage) and the panel is the B operand: a k-block of
// 32 pseudo-points x 128 steps is 32 pieces of 1 KB (one per 4-step group), fetched by the 32 lanes of the producer warp
// with one cp.async.bulk each onto the same mbarrier.  A B fragment (k = 4 pseudo-points x 8 steps) is then two runs of
// 16 consecutive doubles — one per half-warp, bank-conflict free for LDS.64.  The accumulator fragment (8 m' x 8 steps)
// maps onto two 256-byte runs of the output panel, so the epilogue writes the panel layout directly: the consumer of S
// (whiten_tangent_kernel) reads it exactly like beta, and no transposed slab ever exists.
// Main loop per warp and k4-step: 12 LDS.64 + 32 DMMA.8x8x4, as in panel_syrk.cu; persistent CTAs, one per SM,
// 3-stage ring; tiles ordered m'-fastest so that the CTAs of a wave share the panel rows in L2 and C stays L2-resident.
#include "common.cuh"
#include <algorithm>
#include "dmma_pipe.cuh"

namespace {

using namespace dmma;

constexpr int STAGES = 3;
constexpr int STAGE_DOUBLES = GPAR_KT * GPAR_TILE;        // 4096 doubles = 32 KB per operand and stage
constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
constexpr int NCONSUMER_WARPS = 8;
constexpr int NTHREADS = (NCONSUMER_WARPS + 1) * 32;
constexpr int TILE_GROUPS = GPAR_TILE / 4;               // 4-step groups per output tile (128 steps)
constexpr int PIECE_DOUBLES = GPAR_KT * 4;                // one group's share of a k-block: 32 pseudo-points x 4 steps
constexpr size_t SMEM_BYTES = (size_t)2 * STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;

struct GemmArgs {
  const double* Aop;          // the small matrix in operand layout, tile stride a_stride doubles
  const double* in;           // input panel, m-tile stride in_stride doubles
  const double* in_diag;      // triangular solves out of place (nullable): the DIAGONAL tile row (j == m') is read from this panel
                              // (the un-solved right-hand side), the tile rows j < m' from `in` (= out: already solved)
  double* out;                // output panel, m'-tile stride out_stride doubles
  int64_t a_stride, in_stride, out_stride;
  int64_t g_lo, ng;           // 4-step groups [g_lo, g_lo + ng) of the input panel
  int64_t out_g0;             // group g of the input lands at group g - out_g0 of the output
  int mt_lo, n_mt;            // output m'-tiles [mt_lo, mt_lo + n_mt)
  int kb_full;                // k-blocks (of 32 pseudo-points) of a full contraction
  int triangular;             // 1: m'-tile i contracts over the k-blocks of the tiles j <= i only
};

__global__ void __launch_bounds__(NTHREADS, 1)
panel_gemm_kernel(const GemmArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sB = sA + STAGES * STAGE_DOUBLES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * STAGE_DOUBLES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCONSUMER_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();
  const int64_t ntn = (a.ng + TILE_GROUPS - 1) / TILE_GROUPS;
  const int64_t ntiles = ntn * a.n_mt;
  int stage = 0; uint32_t phase = 0;
  if (warp == NCONSUMER_WARPS) {
    // producer warp: lane 0 moves the A stage, every lane one 1 KB piece of the B stage
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t nt = tile / a.n_mt; const int mt = a.mt_lo + (int)(tile % a.n_mt);
      const int64_t g0 = a.g_lo + nt * TILE_GROUPS;
      const int gvalid = (int)((a.g_lo + a.ng - g0 < TILE_GROUPS) ? a.g_lo + a.ng - g0 : TILE_GROUPS);
      const int kb1 = a.triangular ? (mt + 1) * (GPAR_TILE / GPAR_KT) : a.kb_full;
      const double* arow = a.Aop + (int64_t)mt * a.a_stride;
      const int64_t boffs = ((g0 + lane) * GPAR_TILE) * 4;
      for (int kb = 0; kb < kb1; kb++) {
        const double* brow = ((a.in_diag && (kb >> 2) == mt) ? a.in_diag : a.in) + boffs;
        mbar_wait(&empty[stage], phase ^ 1u);
        if (lane == 0) {
          mbar_expect_tx(&full[stage], (uint32_t)(STAGE_BYTES + gvalid * PIECE_DOUBLES * 8));
          bulk_g2s(sA + stage * STAGE_DOUBLES, arow + (int64_t)kb * STAGE_DOUBLES, STAGE_BYTES, &full[stage]);
        }
        if (lane < gvalid)
          bulk_g2s(sB + stage * STAGE_DOUBLES + lane * PIECE_DOUBLES,
                   brow + (int64_t)(kb >> 2) * a.in_stride + (kb & 3) * PIECE_DOUBLES, PIECE_DOUBLES * 8, &full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }
  const int wr = warp >> 1, wc = warp & 1;     // 4 x 2 warps over the 128 (m') x 128 (steps) tile
  // B fragment: lane holds in[step = lane/4][k = lane%4] of an 8-step x 4-pseudo-point block = piece (lane/16) of the
  // two groups, offset (k * 4 + step % 4) inside the k4-run of 16 doubles
  const int boff = (lane >> 4) * PIECE_DOUBLES + (lane & 3) * 4 + ((lane >> 2) & 3);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t nt = tile / a.n_mt; const int mt = a.mt_lo + (int)(tile % a.n_mt);
    const int kb1 = a.triangular ? (mt + 1) * (GPAR_TILE / GPAR_KT) : a.kb_full;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kb = 0; kb < kb1; kb++) {
      mbar_wait(&full[stage], phase);
      const double* A = sA + stage * STAGE_DOUBLES + wr * 32 * 4 + lane;
      const double* B = sB + stage * STAGE_DOUBLES + wc * 16 * PIECE_DOUBLES + boff;
#pragma unroll
      for (int k4 = 0; k4 < GPAR_KT / 4; k4++) {
        double af[4], bf[8];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = A[k4 * (GPAR_TILE * 4) + i * 32];
#pragma unroll
        for (int j = 0; j < 8; j++) bf[j] = B[k4 * 16 + j * 2 * PIECE_DOUBLES];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    // epilogue: accumulator (m' = lane/4, steps 2 (lane%4) + {0,1}) -> the output panel, 16-byte stores; a warp's store
    // covers two runs of 256 contiguous bytes
    const int64_t gbase = a.g_lo + nt * TILE_GROUPS + wc * 16 + ((lane & 3) >> 1);
    double* obase = a.out + (int64_t)mt * a.out_stride + (wr * 32 + (lane >> 2)) * 4 + (lane & 1) * 2;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int64_t g = gbase + 2 * j;
      if (g < a.g_lo + a.ng) {
        double* o = obase + (g - a.out_g0) * (GPAR_TILE * 4);
#pragma unroll
        for (int i = 0; i < 4; i++) *reinterpret_cast<double2*>(o + i * 32) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

// Aop[m't][k/4][m'%128][k%4] <- Q[m' + k M] (col-major M x M), zero padded to Mpad
__global__ void __launch_bounds__(256)
dense_to_operand_kernel(const double* __restrict__ Q, int M, int Mpad, double* __restrict__ Aop) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)Mpad * Mpad) return;
  const int kq = (int)(e & 3), mi = (int)((e >> 2) % GPAR_TILE);
  const int64_t r = (e >> 2) / GPAR_TILE;
  const int k4 = (int)(r % (Mpad / 4)), mt = (int)(r / (Mpad / 4));
  const int mp = mt * GPAR_TILE + mi, k = k4 * 4 + kq;
  Aop[e] = (mp < M && k < M) ? Q[(int64_t)mp + (int64_t)k * M] : 0.0;
}

// Y_ii = L_ii^-1 of every 128 x 128 diagonal block of the lower-triangular L (identity on the padding), row by row from
// Y L = I — the substitution whose LEFT residual Y L - I is small, which is the one a product Y_ii (.) needs.
// Yd: T blocks of 128 x 128, column-major.  One thread per row; the block of L sits in shared memory.
__global__ void __launch_bounds__(GPAR_TILE)
tri_diag_inverse_kernel(const double* __restrict__ L, int M, double* __restrict__ Yd) {
  extern __shared__ double sL[];                 // 128 x 128 column-major
  const int bi = blockIdx.x, r = threadIdx.x, base = bi * GPAR_TILE;
  for (int c = 0; c < GPAR_TILE; c++) {
    const int gr = base + r, gc = base + c;
    sL[r + c * GPAR_TILE] = (gr < M && gc < M) ? (r >= c ? L[(int64_t)gr + (int64_t)gc * M] : 0.0) : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  double* Y = Yd + (int64_t)bi * GPAR_TILE * GPAR_TILE;
  for (int c = GPAR_TILE - 1; c >= 0; c--) {
    if (c > r) { Y[r + c * GPAR_TILE] = 0.0; continue; }
    double acc = (c == r) ? 1.0 : 0.0;
    double a0 = 0.0, a1 = 0.0;                   // two chains: the loads are this thread's own earlier stores
    int k = c + 1;
    for (; k + 1 <= r; k += 2) {
      a0 = fma(Y[r + k * GPAR_TILE], sL[k + c * GPAR_TILE], a0);
      a1 = fma(Y[r + (k + 1) * GPAR_TILE], sL[k + 1 + c * GPAR_TILE], a1);
    }
    if (k <= r) a0 = fma(Y[r + k * GPAR_TILE], sL[k + c * GPAR_TILE], a0);
    acc -= a0 + a1;
    Y[r + c * GPAR_TILE] = acc / sL[c + c * GPAR_TILE];
  }
}

// The block-lower-triangular operand of the blocked solve, in operand layout: C_ii = Y_ii, C_ij = -Y_ii L_ij (j < i).
// grid (T*T, 8), 128 threads: thread = row of tile i, block.y = 16 columns of tile j.
__global__ void __launch_bounds__(GPAR_TILE)
tri_operand_kernel(const double* __restrict__ L, const double* __restrict__ Yd, int M, int Mpad, double* __restrict__ Aop) {
  const int T = Mpad / GPAR_TILE;
  const int i = blockIdx.x / T, j = blockIdx.x % T, r = threadIdx.x;
  if (j > i) return;
  const double* Y = Yd + (int64_t)i * GPAR_TILE * GPAR_TILE;
  for (int cc = 0; cc < 16; cc++) {
    const int c = blockIdx.y * 16 + cc;
    double v;
    if (i == j) v = Y[r + c * GPAR_TILE];
    else {
      const int gc = j * GPAR_TILE + c;        // < M: j < i <= T - 1
      const double* Lc = L + (int64_t)i * GPAR_TILE + (int64_t)gc * M;
      const int kmax = (i * GPAR_TILE + r < M) ? r : -1;      // padded rows of Y are identity rows; L has no such rows
      double a0 = 0.0, a1 = 0.0;
      int k = 0;
      for (; k + 1 <= kmax; k += 2) {
        a0 = fma(Y[r + k * GPAR_TILE], __ldg(Lc + k), a0);
        a1 = fma(Y[r + (k + 1) * GPAR_TILE], __ldg(Lc + k + 1), a1);
      }
      if (k <= kmax) a0 = fma(Y[r + k * GPAR_TILE], __ldg(Lc + k), a0);
      v = -(a0 + a1);
    }
    const int k = j * GPAR_TILE + c;
    Aop[(((int64_t)i * (Mpad / 4) + k / 4) * GPAR_TILE + r) * 4 + (k & 3)] = v;
  }
}

}  // namespace

int launch_dense_to_operand(gpar_ctx* ctx, const double* Q, int M, int Mpad, double* Aop) {
  const int64_t total = (int64_t)Mpad * Mpad;
  LAUNCH(ctx, dense_to_operand_kernel, (int)((total + 255) / 256), 256, 0, Q, M, Mpad, Aop);
  return GPAR_OK;
}

// out groups [g_lo - out_g0, ...) <- in groups [g_lo, g_lo + ng) times the operand matrix, m'-tiles [mt_lo, mt_lo + n_mt).
// in_groups / out_groups: 4-step groups per m-tile of the two panels (their tile strides are groups * 512 doubles).
int panel_gemm_run(gpar_ctx* ctx, const double* Aop, int Mpad, const double* in, int64_t in_groups, double* out, int64_t out_groups,
                   int64_t g_lo, int64_t ng, int64_t out_g0, int mt_lo, int n_mt, bool triangular, const double* in_diag) {
  if (ng <= 0 || n_mt <= 0) return GPAR_OK;
  GemmArgs a;
  a.Aop = Aop; a.in = in; a.out = out; a.in_diag = in_diag;
  a.a_stride = (int64_t)(Mpad / 4) * GPAR_TILE * 4; a.in_stride = in_groups * GPAR_TILE * 4; a.out_stride = out_groups * GPAR_TILE * 4;
  a.g_lo = g_lo; a.ng = ng; a.out_g0 = out_g0; a.mt_lo = mt_lo; a.n_mt = n_mt; a.kb_full = Mpad / GPAR_KT; a.triangular = triangular ? 1 : 0;
  const int64_t ntiles = ((ng + TILE_GROUPS - 1) / TILE_GROUPS) * n_mt;
  const int grid = (int)std::min<int64_t>(ntiles, ctx->num_sms);
  CU(cudaFuncSetAttribute(panel_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  LAUNCH(ctx, panel_gemm_kernel, grid, NTHREADS, SMEM_BYTES, a);
  return GPAR_OK;
}

// The operand of the blocked triangular solve with the Cholesky factor L (M x M lower, column-major): see the header.
// Yd: scratch of T * 128 * 128 doubles; Aop: Mpad * Mpad doubles.
int launch_tri_operand(gpar_ctx* ctx, const double* L, int M, int Mpad, double* Yd, double* Aop) {
  const int T = Mpad / GPAR_TILE;
  const size_t sm = (size_t)GPAR_TILE * GPAR_TILE * sizeof(double);
  CU(cudaFuncSetAttribute(tri_diag_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  LAUNCH(ctx, tri_diag_inverse_kernel, T, GPAR_TILE, sm, L, M, Yd);
  CU(cudaMemsetAsync(Aop, 0, (size_t)Mpad * Mpad * sizeof(double), ctx->stream));
  LAUNCH(ctx, tri_operand_kernel, dim3(T * T, GPAR_TILE / 16), GPAR_TILE, 0, L, Yd, M, Mpad, Aop);
  return GPAR_OK;
}

// panel rows [4 g_lo, 4 (g_lo + ng)) <- (L^-1 row')' in place: one launch per tile row, each contracting over the
// already solved tile rows above it.
// src (nullable): solve OUT OF PLACE — the right-hand sides are read from `src` (same shape), the solution goes to `panel`.
int panel_tri_solve_run(gpar_ctx* ctx, const double* Aop, int Mpad, double* panel, int64_t groups, int64_t g_lo, int64_t ng, const double* src) {
  const int T = Mpad / GPAR_TILE;
  for (int i = 0; i < T; i++) CHK(panel_gemm_run(ctx, Aop, Mpad, panel, groups, panel, groups, g_lo, ng, 0, i, 1, true, src));
  return GPAR_OK;
}
