// predict.cu — the Monte-Carlo prediction of one scaled-GPAR output, batched on the device.
//
// Replaces the loop of get_gpar_scaled_predictions (src/gp/gpar_scaled_inference.jl:89-135):
//   Cfu_star = pairwise(out_kernel, X*_sorted, Z)                       (:89)
//   100 x { fx = Cfu_star * (U_u \ rand(q_u));  _, y_smooth, _ = smooth(time_lgssm_star, y* - fx);
//           f* = fx + [g.m[1] for g in y_smooth] }                      (:110-123)
//   mean(acc), std(acc)                                                 (:125)
// The draws stay with the caller (Julia's global RNG, :94): it passes W[:, j] = U_u \ eps_j (M x S).
// Device: (1) fx = K(X*, Z) W evaluated tile-wise without materialising Cfu_star, fused with
// y* - fx; (2) ONE batched, temporally-parallel smoother over the S residual sequences sharing the
// LGSSM (kalman.cu); (3) sample mean / corrected std over the S draws per location.
#include "common.cuh"
#include <algorithm>

namespace {

// fx[s][n] = sum_m k(x*_n, z_m) W[m][s], ys = y* - fx.  One thread per location n; a block-y covers SC samples so
// that every kernel value (one exp, ~50 FP64 slots) feeds SC FMAs — SC is chosen per call to minimise
// ceil(S / SC) (50 + SC) (S = 100 -> SC = 50: k is evaluated twice per location, not thirteen times as with 8).
template <int KIND, int SC>
__global__ void __launch_bounds__(128)
fx_kernel(const double* __restrict__ X, const double* __restrict__ Z, int64_t N, int M, int DX, double inv_l2, double s_out,
          const double* __restrict__ W, int S, const double* __restrict__ y, double* __restrict__ fx, double* __restrict__ ys, int Sp) {
  extern __shared__ __align__(16) double sm[];
  double* zs = sm;                    // 128 x DX
  double* ws = sm + 128 * DX;         // 128 x SC (SC even): 16-byte aligned for the double2 reads
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int s0 = blockIdx.y * SC;
  double acc[SC];
#pragma unroll
  for (int j = 0; j < SC; j++) acc[j] = 0.0;
  double x[8];
  for (int d = 0; d < DX; d++) x[d] = n < N ? X[n * DX + d] : 0.0;
  for (int m0 = 0; m0 < M; m0 += 128) {
    const int mc = min(128, M - m0);
    __syncthreads();
    for (int e = threadIdx.x; e < mc * DX; e += 128) zs[e] = Z[(int64_t)m0 * DX + e];
    for (int e = threadIdx.x; e < mc * SC; e += 128) {
      int mm = e / SC, j = e % SC;
      ws[e] = (s0 + j < S) ? W[(int64_t)(s0 + j) * M + m0 + mm] : 0.0;
    }
    __syncthreads();
    for (int mm = 0; mm < mc; mm++) {
      double d2 = 0.0;
      for (int d = 0; d < DX; d++) { double df = x[d] - zs[mm * DX + d]; d2 = fma(df, df, d2); }
      double dummy; const double k = s_out * base_kernel_dev<KIND, false>(d2 * inv_l2, dummy);
      const double2* w2 = reinterpret_cast<const double2*>(ws + mm * SC);
#pragma unroll
      for (int j = 0; j < SC / 2; j++) { const double2 w = w2[j]; acc[2 * j] = fma(k, w.x, acc[2 * j]); acc[2 * j + 1] = fma(k, w.y, acc[2 * j + 1]); }
    }
  }
  if (n >= N) return;
  const double yn = y[n];
#pragma unroll
  for (int j = 0; j < SC; j++)
    if (s0 + j < S) { fx[n * Sp + s0 + j] = acc[j]; ys[n * Sp + s0 + j] = yn - acc[j]; }      // time-major [n][Sp]
}

template <int KIND>
int launch_fx(gpar_ctx* ctx, const double* X, const double* Z, int64_t N, int M, int DX, double inv_l2, double s_out,
              const double* W, int S, const double* y, double* fx, double* ys, int Sp) {
  const int cands[4] = {8, 16, 32, 50};
  int SC = 8; long best = -1;
  for (int c : cands) { const long cost = (long)((S + c - 1) / c) * (50 + c); if (best < 0 || cost < best) { best = cost; SC = c; } }
  dim3 grid((unsigned)((N + 127) / 128), (S + SC - 1) / SC);
  const size_t smem = (size_t)(128 * DX + 128 * SC) * sizeof(double);
#define FXCASE(C) case C: \
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute((fx_kernel<KIND, C>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    LAUNCH(ctx, (fx_kernel<KIND, C>), grid, 128, smem, X, Z, N, M, DX, inv_l2, s_out, W, S, y, fx, ys, Sp); break;
  switch (SC) { FXCASE(8) FXCASE(16) FXCASE(32) FXCASE(50) }
#undef FXCASE
  return GPAR_OK;
}

// mean_n = mean_s(fx + sm), std_n = corrected sample std (Julia `std`), 0 for S == 1; fx, smean time-major [n][Sp].
// One warp per location: coalesced row reads, fixed-order shuffle sums.
__global__ void __launch_bounds__(256)
mc_reduce_kernel(const double* __restrict__ fx, const double* __restrict__ smean, int64_t N, int S, int Sp,
                 double* __restrict__ mean, double* __restrict__ sd) {
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const double* a = fx + n * Sp; const double* b = smean + n * Sp;
  double m = 0.0;
  for (int s = lane; s < S; s += 32) m += a[s] + b[s];
  for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
  m /= S;
  double v = 0.0;
  for (int s = lane; s < S; s += 32) { const double d = a[s] + b[s] - m; v = fma(d, d, v); }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) { mean[n] = m; sd[n] = S > 1 ? sqrt(v / (S - 1)) : 0.0; }
}
}  // namespace

extern "C" int gpar_scaled_predict(gpar_ctx* ctx, int k_time, int k_out, const double params[5], const double* W, int32_t S,
                                   double* mean, double* sd) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!params || ((mean == nullptr) != (sd == nullptr)) || S < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "scaled_predict: NULL params, S < 1, or only one of mean / std given");
  if (!W && (ctx->qW_S != S || ctx->qW_M != ctx->M))
    return gpar_fail(ctx, GPAR_ERR_INVALID, "scaled_predict: W == NULL needs a preceding gpar_sample_q_u with the same M and S");
  if (ctx->N < 1 || ctx->M < 1 || ctx->D != ctx->Dz) return gpar_fail(ctx, GPAR_ERR_INVALID, "scaled_predict: merged inputs and pseudo-inputs must be set with equal D");
  if (ctx->Nt != ctx->N || ctx->Ny != ctx->N) return gpar_fail(ctx, GPAR_ERR_INVALID, "scaled_predict: times/outputs must have the merged length N+N* = %lld", (long long)ctx->N);
  if (ctx->has_rvec && ctx->Nr != ctx->N) return gpar_fail(ctx, GPAR_ERR_INVALID, "scaled_predict: noise vector length mismatch");
  if (ctx->D > 8) return gpar_fail(ctx, GPAR_ERR_INVALID, "input dimension D=%d not supported (1..8)", ctx->D);
  CU(cudaSetDevice(ctx->device));
  const int64_t N = ctx->N; const int M = (int)ctx->M, DX = ctx->D;
  const double time_l = params[0], time_s = params[1] * params[1], out_l = params[2], out_s = params[3] * params[3], noise = params[4] * params[4];
  // time-major, S padded to a multiple of 128: fx | ys | smoothed means ; then mean | sd | W
  const int Sp = (S + 127) / 128 * 128;
  CU(ctx->kal_e.reserve(((size_t)3 * Sp * N + 2 * (size_t)N + (size_t)M * S) * sizeof(double)));
  double* fx = ctx->kal_e.as<double>(); double* ys = fx + (size_t)Sp * N; double* smean = ys + (size_t)Sp * N;
  double* dmean = smean + (size_t)Sp * N; double* dsd = dmean + N; double* dW = dsd + N;
  if (W) CU(cudaMemcpyAsync(dW, W, (size_t)M * S * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  else CU(cudaMemcpyAsync(dW, ctx->qW.p, (size_t)M * S * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
  if (Sp != S) CU(cudaMemsetAsync(fx, 0, (size_t)2 * Sp * N * sizeof(double), ctx->stream));      // padding sequences: zero data
  const double inv_l2 = 1.0 / (out_l * out_l);
  const double* X = ctx->X.as<double>(); const double* Z = ctx->Z.as<double>(); const double* y = ctx->y.as<double>();
  switch (k_out) {
    case GPAR_EQ: CHK(launch_fx<GPAR_EQ>(ctx, X, Z, N, M, DX, inv_l2, out_s, dW, S, y, fx, ys, Sp)); break;
    case GPAR_MATERN12: CHK(launch_fx<GPAR_MATERN12>(ctx, X, Z, N, M, DX, inv_l2, out_s, dW, S, y, fx, ys, Sp)); break;
    case GPAR_MATERN32: CHK(launch_fx<GPAR_MATERN32>(ctx, X, Z, N, M, DX, inv_l2, out_s, dW, S, y, fx, ys, Sp)); break;
    case GPAR_MATERN52: CHK(launch_fx<GPAR_MATERN52>(ctx, X, Z, N, M, DX, inv_l2, out_s, dW, S, y, fx, ys, Sp)); break;
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "unknown output kernel code %d", k_out);
  }
  // the S residual sequences share the LGSSM: gains and smoother gains once, then two affine passes per sequence
  CHK(lgssm_smooth_shared(ctx, k_time, time_l, time_s, noise, N, ctx->t.as<double>(), y, ctx->has_rvec ? ctx->rvec.as<double>() : nullptr,
                          ys, Sp, smean));
  LAUNCH(ctx, mc_reduce_kernel, (unsigned)((N + 7) / 8), 256, 0, fx, smean, N, S, Sp, dmean, dsd);
  timer.stop();
  ctx->res_a = dmean; ctx->res_b = dsd; ctx->res_len = N;      // stays resident for gpar_take_test
  if (mean) {
    CU(cudaMemcpyAsync(mean, dmean, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(sd, dsd, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  return GPAR_OK;
}
