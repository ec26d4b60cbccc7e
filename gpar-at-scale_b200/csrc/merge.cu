// merge.cu — the merge / sort / un-sort protocol around the smoother, on the device.
//
// Replaces, in get_sde_predictions (src/gp/temporal_gp_inference.jl:55-66,93-97,111-112) and
// get_gpar_scaled_predictions (src/gp/gpar_scaled_inference.jl:75-87,100-103,132-133):
//   latent = vcat(t, t*);  perm = sortperm(latent);  rev = sortperm(perm)
//   s_times = latent[perm];  s_outputs = vcat(y, zeros(N*))[perm];  s_inputs = vcat(X, X*)[perm]
//   s_noise = vcat(fill(sigma^2, N), fill(1e10, N*))[perm]
//   ... smooth / predict on the sorted arrays ...
//   result[rev][N+1:end]
// Julia's sortperm is stable, so a training time equal to a test time keeps the training point first;
// a stable LSD radix sort of (key = time, value = index) reproduces that order.  The sort itself is the
// library primitive cub::DeviceRadixSort (CUDA toolkit); gathers and the un-sort are kernels of this file.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {

__global__ void merge_keys_kernel(const double* __restrict__ t, int64_t N, const double* __restrict__ ts, int64_t Ns,
                                  double* __restrict__ keys, int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N + Ns) return;
  keys[i] = i < N ? t[i] : ts[i - N];
  idx[i] = (int32_t)i;
}

// sorted position j <- original index perm[j]: outputs (0 at test points), noise vector, inputs; test_pos[i] = j for test point i
__global__ void merge_gather_kernel(const int32_t* __restrict__ perm, int64_t N, int64_t Ns, int D, const double* __restrict__ y,
                                    const double* __restrict__ X, const double* __restrict__ Xs, double sigma2,
                                    double* __restrict__ ys, double* __restrict__ rs, double* __restrict__ Xo, int32_t* __restrict__ test_pos) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N + Ns) return;
  const int64_t o = perm[j];
  const bool train = o < N;
  ys[j] = train ? y[o] : 0.0;
  rs[j] = train ? sigma2 : 1e10;
  for (int d = 0; d < D; d++) Xo[j * D + d] = train ? X[o * D + d] : Xs[(o - N) * D + d];
  if (!train) test_pos[o - N] = (int32_t)j;
}

__global__ void take_test_kernel(const double* __restrict__ a, const double* __restrict__ b, const int32_t* __restrict__ test_pos,
                                 int64_t Ns, double* __restrict__ oa, double* __restrict__ ob) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Ns) return;
  const int64_t j = test_pos[i];
  oa[i] = a[j];
  if (b) ob[i] = b[j];
}

// column d of the merged, sorted inputs at the test locations <- vals (test order)
__global__ void scatter_test_column_kernel(double* __restrict__ X, int D, int d, const int32_t* __restrict__ test_pos,
                                           const double* __restrict__ vals, int64_t Ns) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Ns) X[(int64_t)test_pos[i] * D + d] = vals[i];
}

}  // namespace

// The test-ordered gather of the last result of a merged problem into dst (device, Ns doubles), on the context's stream:
// what gpar_take_test copies to the host and what gpar_group_broadcast sends down the chain.
int merged_gather_test(gpar_ctx* ctx, double* dst_a, double* dst_b) {
  const int64_t Ns = ctx->merged_Ns;
  LAUNCH(ctx, take_test_kernel, (int)((Ns + 255) / 256), 256, 0, ctx->res_a, dst_b ? ctx->res_b : nullptr, ctx->test_pos.as<int32_t>(), Ns, dst_a, dst_b);
  return GPAR_OK;
}

extern "C" {

int gpar_set_merged(gpar_ctx* ctx, const double* t, const double* y, const double* X, int64_t N,
                    const double* ts, const double* Xs, int64_t Ns, int32_t D, double sigma2) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!t || !y || !ts || N < 1 || Ns < 1 || D < 0 || (D > 0 && (!X || !Xs)))
    return gpar_fail(ctx, GPAR_ERR_INVALID, "set_merged: need t, y, ts (and X, Xs when D > 0), N >= 1, N* >= 1");
  if (N + Ns > 2147483647LL) return gpar_fail(ctx, GPAR_ERR_INVALID, "set_merged: N + N* exceeds 2^31 - 1");
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx); ctx->phase_valid = false;
  const int64_t T = N + Ns;
  const size_t Dn = (size_t)std::max<int>(D, 1);
  // staging: t | ts | y | X | Xs | unsorted keys ; ints: idx | perm
  CU(ctx->mrg.reserve(((size_t)N + Ns + N + Dn * T + T) * sizeof(double) + 2 * (size_t)T * sizeof(int32_t)));
  double* d_t = ctx->mrg.as<double>(); double* d_ts = d_t + N; double* d_y = d_ts + Ns; double* d_X = d_y + N;
  double* d_Xs = d_X + Dn * N; double* keys = d_Xs + Dn * Ns;
  int32_t* idx = reinterpret_cast<int32_t*>(keys + T); int32_t* perm = idx + T;
  CU(cudaMemcpyAsync(d_t, t, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_ts, ts, (size_t)Ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_y, y, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (D > 0) {
    CU(cudaMemcpyAsync(d_X, X, (size_t)N * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_Xs, Xs, (size_t)Ns * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(ctx->t.reserve((size_t)T * sizeof(double)));
  CU(ctx->y.reserve((size_t)T * sizeof(double)));
  CU(ctx->rvec.reserve((size_t)T * sizeof(double)));
  if (D > 0) CU(ctx->X.reserve((size_t)T * D * sizeof(double)));
  CU(ctx->test_pos.reserve((size_t)Ns * sizeof(int32_t)));
  LAUNCH(ctx, merge_keys_kernel, (int)((T + 255) / 256), 256, 0, d_t, N, d_ts, Ns, keys, idx);
  size_t tmp_bytes = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, ctx->t.as<double>(), idx, perm, (int)T, 0, 64, ctx->stream));
  CU(ctx->tailws.reserve(tmp_bytes));
  CU(cub::DeviceRadixSort::SortPairs(ctx->tailws.p, tmp_bytes, keys, ctx->t.as<double>(), idx, perm, (int)T, 0, 64, ctx->stream));
  LAUNCH(ctx, merge_gather_kernel, (int)((T + 255) / 256), 256, 0, perm, N, Ns, (int)D, d_y, d_X, d_Xs, sigma2,
         ctx->y.as<double>(), ctx->rvec.as<double>(), D > 0 ? ctx->X.as<double>() : nullptr, ctx->test_pos.as<int32_t>());
  timer.stop();
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->Nt = T; ctx->t_reg_dt = 0.0;
  ctx->Ny = T; ctx->ybatch = 1;
  ctx->Nr = T; ctx->has_rvec = true;
  if (D > 0) { ctx->N = T; ctx->D = D; }
  ctx->merged_N = N; ctx->merged_Ns = Ns;
  ctx->res_a = nullptr; ctx->res_b = nullptr; ctx->res_len = 0;
  return GPAR_OK;
}

int gpar_take_test(gpar_ctx* ctx, double* a_test, double* b_test) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!a_test) return gpar_fail(ctx, GPAR_ERR_INVALID, "take_test: the first output must not be NULL");
  if (ctx->merged_Ns < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "take_test: no merged problem is resident (gpar_set_merged)");
  if (!ctx->res_a || ctx->res_len != ctx->merged_N + ctx->merged_Ns)
    return gpar_fail(ctx, GPAR_ERR_INVALID, "take_test: no smoother / prediction result of the merged length is resident");
  if (b_test && !ctx->res_b) return gpar_fail(ctx, GPAR_ERR_INVALID, "take_test: the last call produced one result array only");
  CU(cudaSetDevice(ctx->device));
  const int64_t Ns = ctx->merged_Ns;
  CU(ctx->scal.reserve((size_t)2 * Ns * sizeof(double)));
  double* oa = ctx->scal.as<double>(); double* ob = oa + Ns;
  CHK(merged_gather_test(ctx, oa, b_test ? ob : nullptr));
  CU(cudaMemcpyAsync(a_test, oa, (size_t)Ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (b_test) CU(cudaMemcpyAsync(b_test, ob, (size_t)Ns * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GPAR_OK;
}

// Column d of the merged inputs at the Ns test locations <- col (host, test order) or, with col == NULL, the context's
// chain buffer (gpar_group_broadcast): the predicted means of an earlier output become the test-time input feature of a
// later output (GPAR_scaled_examples.jl:172 `[test_y1, y2_out]`) without a host round trip.
int gpar_set_merged_test_column(gpar_ctx* ctx, int32_t d, const double* col) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (ctx->merged_Ns < 1 || ctx->N != ctx->merged_N + ctx->merged_Ns)
    return gpar_fail(ctx, GPAR_ERR_INVALID, "set_merged_test_column: no merged problem with inputs is resident (gpar_set_merged with D > 0)");
  if (d < 0 || d >= ctx->D) return gpar_fail(ctx, GPAR_ERR_INVALID, "set_merged_test_column: column %d outside the resident inputs (D = %d)", d, ctx->D);
  CU(cudaSetDevice(ctx->device));
  const int64_t Ns = ctx->merged_Ns;
  if (col) {
    CU(ctx->chain.reserve((size_t)Ns * sizeof(double)));
    CU(cudaMemcpyAsync(ctx->chain.p, col, (size_t)Ns * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->chain_n = Ns;
  } else if (ctx->chain_n != Ns) {
    return gpar_fail(ctx, GPAR_ERR_INVALID, "set_merged_test_column: chain buffer holds %lld values, the merged problem has %lld test locations",
                     (long long)ctx->chain_n, (long long)Ns);
  }
  LAUNCH(ctx, scatter_test_column_kernel, (int)((Ns + 255) / 256), 256, 0, ctx->X.as<double>(), ctx->D, (int)d, ctx->test_pos.as<int32_t>(),
         ctx->chain.as<double>(), Ns);
  CU(cudaStreamSynchronize(ctx->stream));
  return GPAR_OK;
}

}  // extern "C"
