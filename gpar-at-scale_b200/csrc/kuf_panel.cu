// kuf_panel.cu — cross-covariance panel producers.
//
// Replaces `Cfu = cov(f, u)` / `pairwise(k, X, Z)` (src/gp/dtc.jl:104;
// gpar_scaled_inference.jl:89,156) — the reference materialises Cfu as an N x M column-major array.
// Here every Kuf element is evaluated exactly ONCE (FP64 exp/sqrt share the FP64 pipe with DMMA on
// B200 — profiles/peaks_r01.json — so re-evaluating a tile per consumer CTA would tax the SYRK)
// and written in the operand layout the DMMA kernel (panel_syrk.cu) streams with one bulk copy per
// stage:
//     P[mt][n/4][m%128][n%4]      mt = m/128,  N padded to a multiple of GPAR_KT with zeros,
// i.e. per 128-wide M-tile a contiguous run over n; a warp's 8x4 DMMA fragment is 32 consecutive
// doubles.  With GRAD the l*dK/dl panel is produced from the same exp().  The kernel also
// accumulates g = K^T y and h = (l dK/dl)^T y (per n-split partials, reduced in fixed order).
// Bound: HBM write (8 or 16 B per element); X/y reads are warp-broadcast.
#include "common.cuh"

template <int KIND, bool GRAD, int D>
__global__ void __launch_bounds__(GPAR_TILE)
kuf_panel_kernel(const double* __restrict__ X, const double* __restrict__ Z, const double* __restrict__ y,
                 int64_t N, int M, int64_t NB4, int64_t groups_per_split, double inv_l2, double s,
                 double* __restrict__ panelK, double* __restrict__ panelD, double* __restrict__ gpart, int Mpad) {
  const int mt = blockIdx.x, mi = threadIdx.x, m = mt * GPAR_TILE + mi;
  const bool mvalid = m < M;
  double z[D];
#pragma unroll
  for (int d = 0; d < D; d++) z[d] = mvalid ? Z[(int64_t)m * D + d] : 0.0;
  int64_t g0 = (int64_t)blockIdx.y * groups_per_split;
  int64_t g1 = g0 + groups_per_split; if (g1 > NB4) g1 = NB4;
  double gacc = 0.0, hacc = 0.0;
  const double inv_l = sqrt(inv_l2);
  double* outK = panelK + (((int64_t)mt * NB4 + g0) * GPAR_TILE + mi) * 4;
  double* outD = GRAD ? panelD + (((int64_t)mt * NB4 + g0) * GPAR_TILE + mi) * 4 : nullptr;
  for (int64_t g = g0; g < g1; g++) {
    double kv[4], dv[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int64_t n = g * 4 + j;
      bool valid = mvalid && n < N;
      int64_t nn = n < N ? n : N - 1;
      double ld, k;
      if (D == 1) k = base_kernel_from_r<KIND, GRAD>(fabs(__ldg(X + nn) - z[0]) * inv_l, ld);      // no square root in one dimension
      else {
        double d2 = 0.0;
#pragma unroll
        for (int d = 0; d < D; d++) { double df = __ldg(X + nn * D + d) - z[d]; d2 = fma(df, df, d2); }
        k = base_kernel_dev<KIND, GRAD>(d2 * inv_l2, ld);
      }
      double yn = __ldg(y + nn);
      k = valid ? s * k : 0.0;
      kv[j] = k; gacc = fma(k, yn, gacc);
      if (GRAD) { ld = valid ? s * ld : 0.0; dv[j] = ld; hacc = fma(ld, yn, hacc); }
    }
    reinterpret_cast<double2*>(outK)[0] = make_double2(kv[0], kv[1]);
    reinterpret_cast<double2*>(outK)[1] = make_double2(kv[2], kv[3]);
    outK += GPAR_TILE * 4;
    if (GRAD) {
      reinterpret_cast<double2*>(outD)[0] = make_double2(dv[0], dv[1]);
      reinterpret_cast<double2*>(outD)[1] = make_double2(dv[2], dv[3]);
      outD += GPAR_TILE * 4;
    }
  }
  gpart[((int64_t)blockIdx.y * 2 + 0) * Mpad + m] = gacc;
  gpart[((int64_t)blockIdx.y * 2 + 1) * Mpad + m] = GRAD ? hacc : 0.0;
}

// out[v*Mpad + m] = sum_s gpart[(s*nvec + v)*Mpad + m]  (fixed order => deterministic)
__global__ void reduce_gh_kernel(const double* __restrict__ gpart, int nsplit, int Mpad, int nvec, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mpad * nvec) return;
  int v = i / Mpad, m = i % Mpad;
  double acc = 0.0;
  for (int s = 0; s < nsplit; s++) acc += gpart[((int64_t)s * nvec + v) * Mpad + m];
  out[i] = acc;
}

template <int KIND, bool GRAD>
static int launch_d(gpar_ctx* ctx, int D, dim3 grid, const double* X, const double* Z, const double* y, int64_t N, int M,
                    int64_t NB4, int64_t gps, double inv_l2, double s, double* pK, double* pD, double* gpart, int Mpad) {
#define CASE_D(DD) case DD: LAUNCH(ctx, (kuf_panel_kernel<KIND, GRAD, DD>), grid, GPAR_TILE, 0, X, Z, y, N, M, NB4, gps, inv_l2, s, pK, pD, gpart, Mpad); break;
  switch (D) {
    CASE_D(1) CASE_D(2) CASE_D(3) CASE_D(4) CASE_D(5) CASE_D(6) CASE_D(7) CASE_D(8)
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "input dimension D=%d not supported (1..8)", D);
  }
#undef CASE_D
  return GPAR_OK;
}

int launch_kuf_panels(gpar_ctx* ctx, int kind, bool grad, double l, double s, double* panelK, double* panelD,
                      double* gpart, int nsplit, int64_t Npad, int Mpad) {
  int64_t NB4 = Npad / 4;
  int64_t gps = (NB4 + nsplit - 1) / nsplit;
  dim3 grid(Mpad / GPAR_TILE, nsplit);
  const double* X = ctx->X.as<double>(); const double* Z = ctx->Z.as<double>(); const double* y = ctx->y.as<double>();
  double inv_l2 = 1.0 / (l * l);
  int D = ctx->D; int64_t N = ctx->N; int M = (int)ctx->M;
#define CASE_K(KK)                                                                                         \
  case KK:                                                                                                 \
    if (grad) return launch_d<KK, true>(ctx, D, grid, X, Z, y, N, M, NB4, gps, inv_l2, s, panelK, panelD, gpart, Mpad); \
    else return launch_d<KK, false>(ctx, D, grid, X, Z, y, N, M, NB4, gps, inv_l2, s, panelK, panelD, gpart, Mpad);
  switch (kind) {
    CASE_K(GPAR_EQ) CASE_K(GPAR_MATERN12) CASE_K(GPAR_MATERN32) CASE_K(GPAR_MATERN52)
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "unknown kernel code %d", kind);
  }
#undef CASE_K
}

int launch_reduce_gh(gpar_ctx* ctx, const double* gpart, int nsplit, int Mpad, int nvec, double* out) {
  int n = Mpad * nvec;
  LAUNCH(ctx, reduce_gh_kernel, (n + 255) / 256, 256, 0, gpart, nsplit, Mpad, nvec, out);
  return GPAR_OK;
}
