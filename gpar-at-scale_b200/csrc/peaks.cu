// peaks.cu — in-run roofline denominators (bench.py): the FP64 tensor (DMMA.8x8x4) and FP64 vector (DFMA) issue peaks
// of this device from register-only loops, and the HBM copy bandwidth.  ~60 ms in total.  No reference counterpart:
// measurement support for SURVEY 8d ("measure the FP64 DMMA peak on the box and use it as the roofline denominator").
#include "common.cuh"
#include "dmma_pipe.cuh"

namespace {

// mode 0: DFMA (16 independent chains per thread), mode 1: DMMA (8 independent accumulator pairs per warp)
__global__ void __launch_bounds__(512) fp64_issue_kernel(double* out, int iters, int mode, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = i;
  if (mode == 1) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma::dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double4* __restrict__ a, double4* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}

}  // namespace

extern "C" int gpar_measure_peaks(gpar_ctx* ctx, double* dmma_tflops, double* dfma_tflops, double* hbm_copy_gbs) {
  if (!ctx) return GPAR_ERR_INVALID;
  CU(cudaSetDevice(ctx->device));
  const int grid = ctx->num_sms, block = 512, iters = 20000;
  CU(ctx->scal.reserve((size_t)grid * block * sizeof(double)));
  double* out = ctx->scal.as<double>();
  cudaEvent_t e0 = ctx->pev[0], e1 = ctx->pev[3];
  double res[2] = {0.0, 0.0};
  for (int q = 0; q < 2; q++) {
    const int mode = 1 - q;      // DMMA first: the clocks ramp during the first launches
    LAUNCH(ctx, fp64_issue_kernel, grid, block, 0, out, 5000, mode, 1.0000001, 1e-9);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      CU(cudaEventRecord(e0, ctx->stream));
      LAUNCH(ctx, fp64_issue_kernel, grid, block, 0, out, iters, mode, 1.0000001, 1e-9);
      CU(cudaEventRecord(e1, ctx->stream));
      CU(cudaEventSynchronize(e1));
      float ms = 0; CU(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double warps = (double)grid * block / 32;
    const double fl = mode == 0 ? warps * (double)iters * 128 * 32 * 2 : warps * (double)iters * 8 * 256 * 2;
    res[mode] = fl / best / 1e9;
  }
  if (dfma_tflops) *dfma_tflops = res[0];
  if (dmma_tflops) *dmma_tflops = res[1];
  if (hbm_copy_gbs) {
    const size_t n = (size_t)1 << 30;
    DevBuf a, b;
    CU(a.reserve(n)); cudaError_t e = b.reserve(n);
    if (e != cudaSuccess) { a.release(); return gpar_fail(ctx, GPAR_ERR_NOMEM, "measure_peaks: copy buffers"); }
    cudaMemsetAsync(a.p, 1, n, ctx->stream);
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
      cudaEventRecord(e0, ctx->stream);
      copy_kernel<<<ctx->num_sms * 16, 512, 0, ctx->stream>>>(a.as<double4>(), b.as<double4>(), n / 32);
      ctx->launches++;
      cudaEventRecord(e1, ctx->stream);
      cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
    }
    *hbm_copy_gbs = 2.0 * n / best / 1e6;
    a.release(); b.release();
  }
  ctx->phase_valid = false;
  return GPAR_OK;
}
