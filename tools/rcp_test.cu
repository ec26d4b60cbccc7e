// rcp_test.cu — accuracy of the reciprocal used inside the one-pass Kalman loop (lgssm_math.cuh: rcp_pos):
// MUFU.RCP64H seed, after the cubic Newton step (what the library uses), and after a further quadratic step.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/rcp_test tools/rcp_test.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rcp_accuracy_kernel(double* out3) {
  double m0 = 0.0, m3 = 0.0, m5 = 0.0;
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < (1 << 22); i += blockDim.x * gridDim.x) {
    const double s = (1.0 + i * (1.0 / (1 << 22))) * ((i & 1) ? 3.7e5 : 1.3e-4);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    const double ex = 1.0 / s;
    m0 = fmax(m0, fabs(r - ex) / ex);
    double e = fma(-s, r, 1.0); e = fma(e, e, e); r = fma(r, e, r);
    m3 = fmax(m3, fabs(r - ex) / ex);
    e = fma(-s, r, 1.0); r = fma(r, e, r);
    m5 = fmax(m5, fabs(r - ex) / ex);
  }
  for (int o = 16; o > 0; o >>= 1) { m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o)); m3 = fmax(m3, __shfl_xor_sync(0xffffffffu, m3, o)); m5 = fmax(m5, __shfl_xor_sync(0xffffffffu, m5, o)); }
  if ((threadIdx.x & 31) == 0) { atomicMax((unsigned long long*)out3, __double_as_longlong(m0)); atomicMax((unsigned long long*)out3 + 1, __double_as_longlong(m3)); atomicMax((unsigned long long*)out3 + 2, __double_as_longlong(m5)); }

int main() {
  double* d; cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
  rcp_accuracy_kernel<<<148, 256>>>(d);
  double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("max relative error over 4M arguments: rcp.approx seed %.3e, + cubic step %.3e, + quadratic step %.3e\n", h[0], h[1], h[2]);
  return 0;
}
