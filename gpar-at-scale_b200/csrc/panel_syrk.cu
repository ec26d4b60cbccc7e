// panel_syrk.cu — FP64 tensor-core (DMMA) symmetric rank-N update over operand panels.
//
// Replaces `A * A'` (src/gp/dtc.jl:120), `B_ef * B_ef'` (gpar_scaled_inference.jl:187) — the one
// genuine dense contraction of the path:  G = P_K^T P_K  (M x M, contraction over N; only tiles
// i >= j are computed, N M (M+1) flops) and, for the forward-mode length-scale gradient,
// H = P_K^T P_D (full, 2 N M^2 flops).
//
// sm_100a has no tcgen05 .kind::f64: FP64 tensor math is mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4,
// and it shares the FP64 pipe with DFMA (profiles/peaks_r01.json: 37.0 TF each, 36.9 TF mixed), so
// the design goal is a main loop that issues nothing but DMMA and fragment loads:
//  * operands come pre-evaluated in the fragment-native panel layout (kuf_panel.cu) and are
//    staged with ONE cp.async.bulk (TMA engine, SASS UBLKCP) of 32 KB per operand per stage into
//    a 3-stage, mbarrier-synchronised shared-memory ring by a dedicated producer warp;
//  * 8 consumer warps each own a 32 x 64 accumulator tile (32 DMMA per k4-step, 12 LDS.64); the
//    diagonal tiles of G use a triangular 16x16-block mapping (20 DMMA per warp and k4-step);
//  * persistent stream-K schedule: the (tile-job x k-block) space is cut into one contiguous
//    range per CTA (grid = #SMs), so all 148 SMs are busy for any M; a CTA's partial tile goes to a
//    workspace slot and a second kernel sums the slots of each job in fixed order (deterministic,
//    no atomics) and writes G (mirrored) / H column-major.
#include "common.cuh"
#include <algorithm>
#include "syrk_plan.h"
#include "dmma_pipe.cuh"

namespace {

constexpr int STAGES = 3;
constexpr int STAGE_DOUBLES = GPAR_KT * GPAR_TILE;        // 4096 doubles = 32 KB per operand
constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
constexpr int NCONSUMER_WARPS = 8;
constexpr int NTHREADS = (NCONSUMER_WARPS + 1) * 32;
constexpr size_t SMEM_BYTES = (size_t)2 * STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;

typedef SyrkSeg Seg;
typedef SyrkJob Job;

using namespace dmma;

// lower-triangular 16x16 blocks (row, col, valid) of a 128 x 128 diagonal tile, 5 per consumer warp
__constant__ unsigned char c_diag_blocks[8][5][4] = {
  {{7, 0, 1, 0}, {7, 1, 1, 0}, {7, 2, 1, 0}, {7, 3, 1, 0}, {7, 4, 1, 0}},
  {{7, 5, 1, 0}, {7, 6, 1, 0}, {7, 7, 1, 0}, {1, 0, 1, 0}, {1, 1, 1, 0}},
  {{6, 0, 1, 0}, {6, 1, 1, 0}, {6, 2, 1, 0}, {6, 3, 1, 0}, {6, 4, 1, 0}},
  {{6, 5, 1, 0}, {6, 6, 1, 0}, {2, 0, 1, 0}, {2, 1, 1, 0}, {2, 2, 1, 0}},
  {{5, 0, 1, 0}, {5, 1, 1, 0}, {5, 2, 1, 0}, {5, 3, 1, 0}, {5, 4, 1, 0}},
  {{5, 5, 1, 0}, {3, 0, 1, 0}, {3, 1, 1, 0}, {3, 2, 1, 0}, {3, 3, 1, 0}},
  {{4, 0, 1, 0}, {4, 1, 1, 0}, {4, 2, 1, 0}, {4, 3, 1, 0}, {4, 4, 1, 0}},
  {{0, 0, 1, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}}};

__global__ void __launch_bounds__(NTHREADS, 1)
panel_syrk_kernel(const double* __restrict__ pK, const double* __restrict__ pD, int64_t tile_stride /*doubles per M-tile panel*/,
                  const Seg* __restrict__ segs, const int* __restrict__ cta_seg, double* __restrict__ partial) {
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
  const int s0 = cta_seg[blockIdx.x], s1 = cta_seg[blockIdx.x + 1];
  int stage = 0; uint32_t phase = 0;
  if (warp == NCONSUMER_WARPS) {
    if (lane == 0) {
      for (int si = s0; si < s1; si++) {
        const Seg sg = segs[si];
        const double* a = pK + (int64_t)sg.a_tile * tile_stride;
        const double* b = (sg.b_panel ? pD : pK) + (int64_t)sg.b_tile * tile_stride;
        const bool same = (a == b);
        for (int kb = sg.kb0; kb < sg.kb1; kb++) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], same ? STAGE_BYTES : 2 * STAGE_BYTES);
          bulk_g2s(sA + stage * STAGE_DOUBLES, a + (int64_t)kb * STAGE_DOUBLES, STAGE_BYTES, &full[stage]);
          if (!same) bulk_g2s(sB + stage * STAGE_DOUBLES, b + (int64_t)kb * STAGE_DOUBLES, STAGE_BYTES, &full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    return;
  }
  const int wr = warp >> 1, wc = warp & 1;     // 4 x 2 warps over the 128 x 128 tile
  for (int si = s0; si < s1; si++) {
    const Seg sg = segs[si];
    const bool diag = (sg.b_panel == 0 && sg.a_tile == sg.b_tile);
    double* out = partial + (int64_t)sg.slot * (GPAR_TILE * GPAR_TILE);
    if (!diag) {
      double acc[4][8][2];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
      for (int kb = sg.kb0; kb < sg.kb1; kb++) {
        mbar_wait(&full[stage], phase);
        const double* A = sA + stage * STAGE_DOUBLES + wr * 32 * 4 + lane;
        const double* B = sB + stage * STAGE_DOUBLES + wc * 64 * 4 + lane;
#pragma unroll
        for (int k4 = 0; k4 < GPAR_KT / 4; k4++) {
          double af[4], bf[8];
#pragma unroll
          for (int i = 0; i < 4; i++) af[i] = A[k4 * (GPAR_TILE * 4) + i * 32];
#pragma unroll
          for (int j = 0; j < 8; j++) bf[j] = B[k4 * (GPAR_TILE * 4) + j * 32];
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) {
          int r = wr * 32 + i * 8 + (lane >> 2), c = wc * 64 + j * 8 + (lane & 3) * 2;
          *reinterpret_cast<double2*>(out + r * GPAR_TILE + c) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    } else {
      // Diagonal tile of G: only the lower triangle is needed.  In 16x16 blocks that is 36 of 64;
      // a static table gives every warp 5 blocks (4 padding blocks on warp 7), i.e. 20 DMMA per warp
      // and k4-step instead of 32.  Offsets are loop-invariant registers, accumulators are static.
      double acc[5][2][2][2];
#pragma unroll
      for (int i = 0; i < 5; i++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[i][q >> 1][q & 1][0] = acc[i][q >> 1][q & 1][1] = 0.0;
      int offA[5], offB[5];
#pragma unroll
      for (int i = 0; i < 5; i++) { offA[i] = c_diag_blocks[warp][i][0] * 64; offB[i] = c_diag_blocks[warp][i][1] * 64; }
      for (int kb = sg.kb0; kb < sg.kb1; kb++) {
        mbar_wait(&full[stage], phase);
        const double* P = sA + stage * STAGE_DOUBLES + lane;
#pragma unroll
        for (int k4 = 0; k4 < GPAR_KT / 4; k4++) {
          const double* Pk = P + k4 * (GPAR_TILE * 4);
          double a[5][2], b[5][2];
#pragma unroll
          for (int i = 0; i < 5; i++) { a[i][0] = Pk[offA[i]]; a[i][1] = Pk[offA[i] + 32]; b[i][0] = Pk[offB[i]]; b[i][1] = Pk[offB[i] + 32]; }
#pragma unroll
          for (int i = 0; i < 5; i++)
#pragma unroll
            for (int q = 0; q < 4; q++) dmma884(acc[i][q >> 1][q & 1][0], acc[i][q >> 1][q & 1][1], a[i][q >> 1], b[i][q & 1]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
#pragma unroll
      for (int i = 0; i < 5; i++) {
        if (c_diag_blocks[warp][i][2]) {
#pragma unroll
          for (int q = 0; q < 4; q++) {
            int r = (offA[i] >> 2) + (q >> 1) * 8 + (lane >> 2), c = (offB[i] >> 2) + (q & 1) * 8 + (lane & 3) * 2;
            *reinterpret_cast<double2*>(out + r * GPAR_TILE + c) = make_double2(acc[i][q >> 1][q & 1][0], acc[i][q >> 1][q & 1][1]);
          }
        }
      }
    }
  }
}

// One block per (job, 8-row strip): fixed-order sum of the job's partial tiles, then the
// column-major M x M write (G: lower tile + mirror; H: plain).
__global__ void __launch_bounds__(256)
syrk_reduce_kernel(const Job* __restrict__ jobs, const double* __restrict__ partial, int M, double* __restrict__ G, double* __restrict__ H) {
  const Job jb = jobs[blockIdx.x];
  const int r0 = blockIdx.y * 16;
  for (int e = threadIdx.x; e < 16 * GPAR_TILE; e += blockDim.x) {
    int r = r0 + e / GPAR_TILE, c = e % GPAR_TILE;
    // fixed order (deterministic), four independent chains so that the loads of a long slot list overlap
    double v4[4] = {0.0, 0.0, 0.0, 0.0};
    const double* p0 = partial + (int64_t)jb.slot0 * (GPAR_TILE * GPAR_TILE) + r * GPAR_TILE + c;
    int s = 0;
    for (; s + 4 <= jb.nslots; s += 4) {
#pragma unroll
      for (int q = 0; q < 4; q++) v4[q] += p0[(int64_t)(s + q) * (GPAR_TILE * GPAR_TILE)];
    }
    for (; s < jb.nslots; s++) v4[0] += p0[(int64_t)s * (GPAR_TILE * GPAR_TILE)];
    const double v = (v4[0] + v4[1]) + (v4[2] + v4[3]);
    int gr = jb.a_tile * GPAR_TILE + r, gc = jb.b_tile * GPAR_TILE + c;
    if (gr >= M || gc >= M) continue;
    if (jb.b_panel == 0) {
      if (jb.a_tile == jb.b_tile && c > r) continue;          // diagonal tile: keep lower, mirror
      G[(int64_t)gr + (int64_t)gc * M] = v;
      G[(int64_t)gc + (int64_t)gr * M] = v;
    } else {
      H[(int64_t)gr + (int64_t)gc * M] = v;
    }
  }
}

}  // namespace

int panel_syrk_run(gpar_ctx* ctx, const double* panelK, const double* panelD, int64_t Npad, int Mpad, int M,
                   bool with_h, double* G, double* H) {
  const int T = Mpad / GPAR_TILE;
  const int64_t NBK = Npad / GPAR_KT;
  // the optimiser re-evaluates on fixed shapes: plan once, keep the segment tables on the device
  if (ctx->plan_T != T || ctx->plan_NBK != NBK || ctx->plan_h != (int)with_h) {
    // small problems: a CTA should own at least 16 k-blocks (512 steps) of work, otherwise the 148-way stream-K split
    // produces 148 partial tiles of a few k-blocks each and the fixed-order reduction dominates (290 us at N = 8 496)
    const int64_t njobs = (int64_t)T * (T + 1) / 2 + (with_h ? (int64_t)T * T : 0);
    const int ctas = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->num_sms, njobs * NBK / 16));
    SyrkPlan pl = plan_syrk(T, NBK, with_h, ctas);
    const size_t nseg = pl.segs.size();
    const int Jn = (int)pl.jobs.size();
    CU(ctx->partial.reserve(nseg * GPAR_TILE * GPAR_TILE * sizeof(double)));
    CU(ctx->segs.reserve(nseg * sizeof(Seg) + (pl.C + 1) * sizeof(int) + 64));
    CU(ctx->jobs.reserve(Jn * sizeof(Job)));
    Seg* ds = ctx->segs.as<Seg>();
    int* dc = reinterpret_cast<int*>(reinterpret_cast<char*>(ds) + nseg * sizeof(Seg));
    CU(cudaMemcpyAsync(ds, pl.segs.data(), nseg * sizeof(Seg), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dc, pl.cta_seg.data(), (pl.C + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->jobs.p, pl.jobs.data(), Jn * sizeof(Job), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope below
    ctx->plan_T = T; ctx->plan_NBK = NBK; ctx->plan_h = (int)with_h; ctx->plan_C = pl.C; ctx->plan_J = Jn; ctx->plan_nseg = nseg;
  }
  const int J = ctx->plan_J, C = ctx->plan_C;
  Seg* dsegs = ctx->segs.as<Seg>();
  int* dcta = reinterpret_cast<int*>(reinterpret_cast<char*>(dsegs) + ctx->plan_nseg * sizeof(Seg));
  CU(cudaFuncSetAttribute(panel_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  const int64_t tile_stride = (Npad / 4) * (GPAR_TILE * 4);
  cudaEventRecord(ctx->pev[1], ctx->stream);
  LAUNCH(ctx, panel_syrk_kernel, C, NTHREADS, SMEM_BYTES, panelK, panelD, tile_stride, dsegs, dcta, ctx->partial.as<double>());
  cudaEventRecord(ctx->pev[2], ctx->stream);
  LAUNCH(ctx, syrk_reduce_kernel, dim3(J, GPAR_TILE / 16), 256, 0, ctx->jobs.as<Job>(), ctx->partial.as<double>(), M, G, H);
  return GPAR_OK;
}
