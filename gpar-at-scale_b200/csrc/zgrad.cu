// zgrad.cu — gradient of the DTC / VFE objective with respect to the pseudo-inputs Z.
//
// The reference keeps Z fixed (examples/dtc_example.jl:82-90; GPAR_scaled_examples.jl:115,145-148);
// pseudo-input optimisation is what the Titsias bound is for, so the library supplies dF/dZ next to
// dF/dtheta.  With the tail's P = (cov(u) + G/sigma^2)^-1, w = P g, T = P + w w'/sigma^4 (dense_tail.cu):
//   dF/dKuf[m, n] = (A Kuf)[m, n] + w_m q_n,    A = -P/sigma^2 (+ cov(u)^-1/sigma^2 for VFE),
//                                               q = (y - Kfu w/sigma^2)/sigma^4
//   dF/dcov(u)    = -1/2 (T - cov(u)^-1)  (- 1/2 cov(u)^-1 G cov(u)^-1/sigma^2 for VFE)
//   dk(x, z)/dz   = s f'(d2) 2 (z - x)/l^2,  d2 = |x - z|^2/l^2
// S = A Kuf is one M x M x N contraction — the DMMA panel-GEMM of panel_gemm.cu on slabs of the K panel, as in the
// scaled gradient; the contraction with dk/dz is a kernel of this file (one thread per pseudo-input).
#include "common.cuh"
#include <algorithm>

namespace {

// d kappa / d d2 for kappa as a function of d2 = r^2
template <int KIND>
__device__ __forceinline__ double base_kernel_dd2(double d2) {
  if (KIND == GPAR_EQ) return -0.5 * exp(-0.5 * d2);
  if (KIND == GPAR_MATERN12) { const double r = sqrt(d2); return r > 0.0 ? -0.5 * exp(-r) / r : 0.0; }   // not differentiable at 0: subgradient 0
  if (KIND == GPAR_MATERN32) { const double a = sqrt(3.0 * d2); return -1.5 * exp(-a); }
  const double a = sqrt(5.0 * d2); return -(5.0 / 6.0) * (1.0 + a) * exp(-a);
}

// A = -P/sigma^2 (+ Kinv/sigma^2);  Cb = -1/2 (P + w w'/sigma^4 - Kinv) (- 1/2 Cm/sigma^2)
__global__ void zgrad_matrices_kernel(int M, const double* __restrict__ P, const double* __restrict__ Kinv, const double* __restrict__ Cm,
                                      const double* __restrict__ w, double ip, double* __restrict__ A, double* __restrict__ Cb) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * M) return;
  const int a = (int)(e % M), b = (int)(e / M);
  const double p = P[e], ki = Kinv[e];
  A[e] = -p * ip + (Cm ? ki * ip : 0.0);
  Cb[e] = -0.5 * (p + w[a] * w[b] * ip * ip - ki) - (Cm ? 0.5 * Cm[e] * ip : 0.0);
}

// partial[(split * D + d) * Mpad + m] = sum over the split's steps of (S'[m, n] + w_m q_n) s f'(d2) 2 (z_md - x_nd)/l^2
template <int KIND, int D>
__global__ void __launch_bounds__(GPAR_TILE)
zgrad_cross_kernel(const double* __restrict__ X, const double* __restrict__ Z, const double* __restrict__ St, int64_t s_groups, const double* __restrict__ q,
                   const double* __restrict__ w, int64_t n_lo, int64_t n_hi, int64_t per_split, int M, int Mpad, double inv_l2, double s,
                   double* __restrict__ partial) {
  const int m = blockIdx.x * GPAR_TILE + threadIdx.x;
  const bool mvalid = m < M;
  double z[D], acc[D];
#pragma unroll
  for (int d = 0; d < D; d++) { z[d] = mvalid ? Z[(int64_t)m * D + d] : 0.0; acc[d] = 0.0; }
  const double wm = mvalid ? w[m] : 0.0;
  const int64_t a = n_lo + (int64_t)blockIdx.y * per_split, b = a + per_split < n_hi ? a + per_split : n_hi;
  const double* Sp = St + ((int64_t)blockIdx.x * s_groups * GPAR_TILE + threadIdx.x) * 4;      // slab panel of S: [mt][group][m%128][n%4]
  for (int64_t n = a; n < b; n++) {
    double df[D], d2 = 0.0;
#pragma unroll
    for (int d = 0; d < D; d++) { df[d] = z[d] - __ldg(X + n * D + d); d2 = fma(df[d], df[d], d2); }
    const double r = mvalid ? fma(wm, __ldg(q + n), __ldg(Sp + ((n - n_lo) >> 2) * (GPAR_TILE * 4) + ((n - n_lo) & 3))) : 0.0;
    const double c = r * s * base_kernel_dd2<KIND>(d2 * inv_l2) * 2.0 * inv_l2;
#pragma unroll
    for (int d = 0; d < D; d++) acc[d] = fma(c, df[d], acc[d]);
  }
#pragma unroll
  for (int d = 0; d < D; d++) partial[((int64_t)blockIdx.y * D + d) * Mpad + m] = acc[d];
}

// gz[m*D + d] = sum of the cross partials + 2 sum_m' Cb[m, m'] dKuu[m, m']/dz_m
template <int KIND>
__global__ void zgrad_finish_kernel(const double* __restrict__ Z, const double* __restrict__ Cb, const double* __restrict__ partial, int nparts,
                                    int M, int Mpad, int D, double inv_l2, double s, double* __restrict__ gz) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  for (int d = 0; d < D; d++) {
    double acc = 0.0;
    for (int sp = 0; sp < nparts; sp++) acc += partial[((int64_t)sp * D + d) * Mpad + m];
    gz[(int64_t)m * D + d] = acc;
  }
  for (int mp = 0; mp < M; mp++) {
    if (mp == m) continue;
    double d2 = 0.0;
    for (int d = 0; d < D; d++) { const double df = Z[(int64_t)m * D + d] - Z[(int64_t)mp * D + d]; d2 = fma(df, df, d2); }
    const double c = 2.0 * Cb[(int64_t)m + (int64_t)mp * M] * s * base_kernel_dd2<KIND>(d2 * inv_l2) * 2.0 * inv_l2;
    for (int d = 0; d < D; d++) gz[(int64_t)m * D + d] = fma(c, Z[(int64_t)m * D + d] - Z[(int64_t)mp * D + d], gz[(int64_t)m * D + d]);
  }
}

template <int KIND>
int zgrad_run(gpar_ctx* ctx, const GpParams& p, int vfe, double* grad_Z) {
  const int64_t N = ctx->N; const int M = (int)ctx->M, D = ctx->D;
  const int Mpad = (M + GPAR_TILE - 1) / GPAR_TILE * GPAR_TILE, T = Mpad / GPAR_TILE;
  const int64_t Npad = (N + GPAR_KT - 1) / GPAR_KT * GPAR_KT, NB4 = Npad / 4;
  const size_t MM = (size_t)M * M;
  const double ip = 1.0 / p.noise;
  TailBufs tb;
  CHK(tail_layout(ctx, true, vfe, &tb));
  // A | Cb | w/sigma^2 | q | gz after the tail's own matrices would need a re-layout: they live in kal_b / kal_d
  CU(ctx->kal_b.reserve((2 * MM + (size_t)Mpad) * sizeof(double)));
  double* A = ctx->kal_b.as<double>(); double* Cb = A + MM; double* ws = Cb + MM;
  CU(ctx->kal_d.reserve(((size_t)N + (size_t)M * D + 16) * sizeof(double)));
  double* q = ctx->kal_d.as<double>(); double* gz = q + N;
  LAUNCH(ctx, zgrad_matrices_kernel, (int)((MM + 255) / 256), 256, 0, M, tb.Pm, tb.Kinv, vfe ? tb.Cm : nullptr, tb.wvec, ip, A, Cb);
  // q = (y - K w/sigma^2)/sigma^4: residual with the scaled weights, then a scale
  CU(cudaMemcpyAsync(ws, tb.wvec, (size_t)M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CHK(dla_scal(ctx, M, ip, ws));
  CHK(launch_panel_residual(ctx, ctx->panelK.as<double>(), ws, ctx->y.as<double>(), N, NB4, T, M, q));
  const double ip2 = ip * ip;
  CHK(dla_scal(ctx, (long long)N, ip2, q));
  // slabs of the K panel -> S = K A (DMMA panel-GEMM, A symmetric) -> contraction with dk/dz
  const int64_t slab_steps = std::min<int64_t>(Npad, 131072);
  const int nslab = (int)((Npad + slab_steps - 1) / slab_steps);
  int nsplit = std::max(1, (ctx->num_sms * 8) / T);
  nsplit = (int)std::min<int64_t>(nsplit, std::max<int64_t>(1, slab_steps / 64));
  CU(ctx->panelB.reserve((size_t)Mpad * Mpad * sizeof(double)));
  CU(ctx->kal_f.reserve((size_t)slab_steps * Mpad * sizeof(double)));
  CU(ctx->gpart.reserve((size_t)nslab * nsplit * D * Mpad * sizeof(double)));
  double* Aop = ctx->panelB.as<double>(); double* St = ctx->kal_f.as<double>(); double* partial = ctx->gpart.as<double>();
  const double inv_l2 = 1.0 / (p.l * p.l);
  CHK(launch_dense_to_operand(ctx, A, M, Mpad, Aop));
  const double* X = ctx->X.as<double>(); const double* Z = ctx->Z.as<double>();
  for (int sl = 0; sl < nslab; sl++) {
    const int64_t n_lo = (int64_t)sl * slab_steps, g_lo = n_lo / 4, ng = std::min<int64_t>(slab_steps / 4, NB4 - g_lo);
    const int64_t n_hi = std::min<int64_t>(N, n_lo + ng * 4);
    if (n_hi <= n_lo) { CU(cudaMemsetAsync(partial + (size_t)sl * nsplit * D * Mpad, 0, (size_t)nsplit * D * Mpad * sizeof(double), ctx->stream)); continue; }
    CHK(panel_gemm_run(ctx, Aop, Mpad, ctx->panelK.as<double>(), NB4, St, slab_steps / 4, g_lo, ng, g_lo, 0, T, false));
    const int64_t per = (n_hi - n_lo + nsplit - 1) / nsplit;
    dim3 grid(T, nsplit);
    double* part = partial + (size_t)sl * nsplit * D * Mpad;
#define ZCASE(DD) case DD: LAUNCH(ctx, (zgrad_cross_kernel<KIND, DD>), grid, GPAR_TILE, 0, X, Z, St, slab_steps / 4, q, tb.wvec, n_lo, n_hi, per, M, Mpad, inv_l2, p.s, part); break;
    switch (D) {
      ZCASE(1) ZCASE(2) ZCASE(3) ZCASE(4) ZCASE(5) ZCASE(6) ZCASE(7) ZCASE(8)
      default: return gpar_fail(ctx, GPAR_ERR_INVALID, "input dimension D=%d not supported (1..8)", D);
    }
#undef ZCASE
  }
  LAUNCH(ctx, zgrad_finish_kernel<KIND>, (M + 127) / 128, 128, 0, Z, Cb, partial, nslab * nsplit, M, Mpad, D, inv_l2, p.s, gz);
  CU(cudaMemcpyAsync(grad_Z, gz, (size_t)M * D * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GPAR_OK;
}

}  // namespace

extern "C" int gpar_dtc_logpdf_zgrad(gpar_ctx* ctx, int kernel, const double theta[3], int vfe, double jitter, double* val,
                                     double* grad_theta, double* grad_Z) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!grad_theta || !grad_Z) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc_logpdf_zgrad: grad_theta and grad_Z must not be NULL");
  // value, d/dtheta and the tail's matrices (P, w, cov(u)^-1, ...) — they stay resident in the context
  CHK(gpar_dtc_logpdf(ctx, kernel, theta, vfe, jitter, val, grad_theta));
  const GpParams p = unpack_gp3(theta);
  const double t_prev = ctx->last_ms; const int64_t l_prev = ctx->last_launches;
  int rc;
  {
    CallTimer timer(ctx); gpar_drop_result(ctx);
    switch (kernel) {
      case GPAR_EQ: rc = zgrad_run<GPAR_EQ>(ctx, p, vfe, grad_Z); break;
      case GPAR_MATERN12: rc = zgrad_run<GPAR_MATERN12>(ctx, p, vfe, grad_Z); break;
      case GPAR_MATERN32: rc = zgrad_run<GPAR_MATERN32>(ctx, p, vfe, grad_Z); break;
      default: rc = zgrad_run<GPAR_MATERN52>(ctx, p, vfe, grad_Z); break;
    }
  }
  ctx->last_ms += t_prev; ctx->last_launches += l_prev;
  return rc;
}
