// Step-0 roofline denominators for the FP64 path on B200 (sm_100a).
// Measures: DFMA peak, DMMA (mma.sync m8n8k4 f64) peak, DFMA+DMMA issued together
// (do they share a pipe?), FP64 exp() throughput, cuBLAS DGEMM, HBM copy.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/peaks.cu -lcublas -o tools/peaks
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: all warps DFMA; 1: all warps DMMA; 2: even warps DMMA, odd warps DFMA
__global__ void __launch_bounds__(512) k_mix(double *out, int iters, int mode, double a0, double b0) {
  int warp = threadIdx.x >> 5;
  bool do_mma = (mode == 1) || (mode == 2 && (warp & 1) == 0);
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  if (do_mma) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 8; r++)   // 8*16 = 128 DFMA per iter per thread = 128 MAC/lane
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) k_exp(double *out, int iters, double x0) {
  double x = x0 - (threadIdx.x & 31) * 0.37 - blockIdx.x * 1e-3;
  double acc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { acc += exp(x); x -= 1e-3; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void k_copy(const double4 *__restrict__ a, double4 *__restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) { float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms; }

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"", p.name, nsm, p.major, p.minor);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 4 * 512));
  const char *names[3] = {"dfma", "dmma", "mixed"};
  for (int mode = 0; mode < 3; mode++) {
    for (int bpsm = 1; bpsm <= 2; bpsm++) {
      int iters = 20000, grid = nsm * bpsm, block = 512;
      k_mix<<<grid, block>>>(out, 100, mode, 1.0000001, 1e-9); CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0)); k_mix<<<grid, block>>>(out, iters, mode, 1.0000001, 1e-9); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); best = fminf(best, time_ms(e0, e1));
      }
      // flops: DFMA warp: iters*128 FMA/lane*32 lanes*2; DMMA warp: iters*8*(8*8*4)*2
      double warps = (double)grid * block / 32;
      double f_fma = (double)iters * 128 * 32 * 2, f_mma = (double)iters * 8 * 256 * 2;
      double fl = mode == 0 ? warps * f_fma : mode == 1 ? warps * f_mma : warps / 2 * (f_fma + f_mma);
      printf(", \"%s_b%d_tflops\": %.3f", names[mode], bpsm, fl / best / 1e9);
    }
  }
  {
    int iters = 2000, grid = nsm * 8, block = 256;
    k_exp<<<grid, block>>>(out, 10, -0.1); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      CK(cudaEventRecord(e0)); k_exp<<<grid, block>>>(out, iters, -0.1); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); best = fminf(best, time_ms(e0, e1));
    }
    printf(", \"exp_gevals_per_s\": %.3f", (double)grid * block * iters * 8 / best / 1e6);
  }
  {
    size_t n = (size_t)1 << 30;  // 1 GiB each way
    double4 *a, *b; CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n)); CK(cudaMemset(a, 1, n));
    float best = 1e30f;
    for (int rep = 0; rep < 10; rep++) {
      CK(cudaEventRecord(e0)); k_copy<<<nsm * 16, 512>>>(a, b, n / 32); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); best = fminf(best, time_ms(e0, e1));
    }
    printf(", \"hbm_copy_gbs\": %.1f", 2.0 * n / best / 1e6);
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  {
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {4096, 8192}) {
      double *A, *B, *C; size_t bytes = sizeof(double) * n * n;
      CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
      std::vector<double> hA((size_t)n * n);
      for (size_t i = 0; i < hA.size(); i++) hA[i] = (double)rand() / RAND_MAX - 0.5;
      CK(cudaMemcpy(A, hA.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hA.data(), bytes, cudaMemcpyHostToDevice));
      double one = 1, zero = 0;
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0)); cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); best = fminf(best, time_ms(e0, e1));
      }
      printf(", \"dgemm_%d_tflops\": %.3f", n, 2.0 * n * n * n / best / 1e9);
      // sustained: loop for ~3 s
      if (n == 8192) {
        int reps = (int)(3000.0f / best) + 1;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; r++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        printf(", \"dgemm_8192_sustained_tflops\": %.3f", 2.0 * n * n * n * reps / time_ms(e0, e1) / 1e9);
        // DSYRK (what cuBLAS gets with symmetry exploited), k large: n=1024, k=262144
      }
      CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
    }
    {
      int m = 1024; size_t k = 262144; double *A, *C;
      CK(cudaMalloc(&A, sizeof(double) * m * k)); CK(cudaMalloc(&C, sizeof(double) * m * m)); CK(cudaMemset(A, 0, sizeof(double) * m * k));
      double one = 1, zero = 0;
      cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, m, (int)k, &one, A, m, &zero, C, m); CK(cudaDeviceSynchronize());
      float best = 1e30f;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0)); cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, m, (int)k, &one, A, m, &zero, C, m); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); best = fminf(best, time_ms(e0, e1));
      }
      printf(", \"dsyrk_1024x262144_ms\": %.3f, \"dsyrk_sym_tflops\": %.3f", best, (double)m * (m + 1) * k / best / 1e9);
      CK(cudaFree(A)); CK(cudaFree(C));
    }
    cublasDestroy(h);
  }
  printf("}\n");
  return 0;
}
