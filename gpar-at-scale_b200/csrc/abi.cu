// abi.cu — extern "C" entry points of libgpar_b200.so (include/gpar_b200.h): context, resident
// data, and the host-side orchestration of each hot-path call.  No torch types, no CPU fallback:
// every compute entry point runs the CUDA kernels of this library or returns an error status.
#include "common.cuh"
#include <algorithm>

int gpar_fail(gpar_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (c) c->err = buf;
  return code;
}

namespace {
int upload(gpar_ctx* ctx, DevBuf& buf, const double* src, size_t count) {
  CU(cudaSetDevice(ctx->device));
  CU(buf.reserve(std::max<size_t>(count, 1) * sizeof(double)));
  if (count) CU(cudaMemcpyAsync(buf.p, src, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GPAR_OK;
}

__global__ void sumsq_kernel(const double* __restrict__ y, int64_t n, double* __restrict__ part) {
  __shared__ double sh[32];
  double acc = 0.0;
  // fixed chunking: block b sums a contiguous range in fixed order (deterministic)
  int64_t per = (n + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = i0 + per < n ? i0 + per : n;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) acc = fma(y[i], y[i], acc);
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}
__global__ void sum_final_kernel(const double* part, int n, double* out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = r;
}
}  // namespace

// Sufficient statistics of the plain DTC objective over THIS context's resident slice of the data: panel producers and
// the DMMA SYRK are enqueued on the context's stream, nothing is read back.  *stats -> [G (M*M) | H (M*M) | g (Mpad) |
// h (Mpad) | y'y (1)], *count doubles.  Every term is a sum over data points, so slices on different devices add up
// (gpar_group_dtc_logpdf_sharded all-reduces this buffer).
int dtc_slice_stats(gpar_ctx* ctx, int kernel, const GpParams& p, bool want_grad, double** stats, size_t* count, int whiten_vfe) {
  if (ctx->N < 1 || ctx->M < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc: inputs (set_inputs) and pseudo-inputs (set_pseudo) must be set");
  if (ctx->D != ctx->Dz) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc: X has D=%d but Z has D=%d", ctx->D, ctx->Dz);
  if (ctx->Ny != ctx->N || ctx->ybatch < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc: outputs length %lld != N %lld", (long long)ctx->Ny, (long long)ctx->N);
  CU(cudaSetDevice(ctx->device));
  const int64_t N = ctx->N; const int M = (int)ctx->M;
  const int Mpad = (M + GPAR_TILE - 1) / GPAR_TILE * GPAR_TILE;
  const int64_t Npad = (N + GPAR_KT - 1) / GPAR_KT * GPAR_KT;
  const size_t panel_bytes = (size_t)Npad * Mpad * sizeof(double);
  CU(ctx->panelK.reserve(panel_bytes));
  if (want_grad) CU(ctx->panelD.reserve(panel_bytes));
  const int T = Mpad / GPAR_TILE;
  int nsplit = std::max(1, (ctx->num_sms * 16) / T);
  nsplit = (int)std::min<int64_t>(nsplit, std::max<int64_t>(1, Npad / 4));
  CU(ctx->gpart.reserve((size_t)nsplit * 2 * Mpad * sizeof(double) + 1024 * sizeof(double)));
  const size_t MM = (size_t)M * M;
  CU(ctx->kal_a.reserve((2 * MM + 2 * (size_t)Mpad + 8) * sizeof(double)));
  double* G = ctx->kal_a.as<double>(); double* H = G + MM; double* gh = H + MM; double* dyy = gh + 2 * Mpad;
  if (!want_grad) CU(cudaMemsetAsync(H, 0, MM * sizeof(double), ctx->stream));       // the whole buffer is reduced
  CHK(launch_kuf_panels(ctx, kernel, want_grad, p.l, p.s, ctx->panelK.as<double>(), ctx->panelD.as<double>(),
                        ctx->gpart.as<double>(), nsplit, Npad, Mpad));
  CHK(launch_reduce_gh(ctx, ctx->gpart.as<double>(), nsplit, Mpad, 2, gh));
  double* ypart = ctx->gpart.as<double>() + (size_t)nsplit * 2 * Mpad;
  LAUNCH(ctx, sumsq_kernel, 1024, 256, 0, ctx->y.as<double>(), N, ypart);
  LAUNCH(ctx, sum_final_kernel, 1, 256, 0, ypart, 1024, dyy);
  if (whiten_vfe >= 0) {      // poorly conditioned cov(u): this member's dtc_tail_prepare left L_u on its side stream — whiten the
    TailBufs tb;              // slice's panel(s) before the SYRK, so that the summed statistics are A A' (and A A_D') themselves
    CHK(tail_layout(ctx, want_grad, whiten_vfe, &tb));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
    CHK(panel_left_solve(ctx, ctx->panelK.as<double>(), Npad, Mpad, M, tb.Lu));
    if (want_grad) CHK(panel_left_solve(ctx, ctx->panelD.as<double>(), Npad, Mpad, M, tb.Lu));
  }
  CHK(panel_syrk_run(ctx, ctx->panelK.as<double>(), ctx->panelD.as<double>(), Npad, Mpad, M, want_grad, G, H));
  *stats = G; *count = 2 * MM + 2 * (size_t)Mpad + 1;
  return GPAR_OK;
}

extern "C" {

int gpar_abi_version(void) { return GPAR_ABI_VERSION; }

int gpar_ctx_create(int device, gpar_ctx** out) {
  if (!out) return GPAR_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return GPAR_ERR_CUDA;   // no GPU: fail loudly, no CPU fallback
  if (device < 0 || device >= ndev) return GPAR_ERR_INVALID;
  gpar_ctx* ctx = new gpar_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->pev[0]) != cudaSuccess || cudaEventCreate(&ctx->pev[1]) != cudaSuccess ||
      cudaEventCreate(&ctx->pev[2]) != cudaSuccess || cudaEventCreate(&ctx->pev[3]) != cudaSuccess) {
    delete ctx;
    return GPAR_ERR_CUDA;
  }
  ctx->num_sms = prop.multiProcessorCount;
  *out = ctx;
  return GPAR_OK;
}

int gpar_ctx_destroy(gpar_ctx* ctx) {
  if (!ctx) return GPAR_OK;
  for (gpar_ctx* lane : ctx->lanes) gpar_ctx_destroy(lane);
  ctx->lanes.clear();
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
  DevBuf* bufs[] = {&ctx->X, &ctx->Z, &ctx->t, &ctx->y, &ctx->rvec, &ctx->panelK, &ctx->panelD, &ctx->panelB, &ctx->panelA, &ctx->kal_f, &ctx->partial, &ctx->segs,
                    &ctx->jobs, &ctx->gpart, &ctx->scal, &ctx->dense, &ctx->tailws, &ctx->info,
                    &ctx->kal_a, &ctx->kal_b, &ctx->kal_c, &ctx->kal_d, &ctx->kal_e, &ctx->qW, &ctx->dla_ws, &ctx->dla_ws_side, &ctx->dla_ws2, &ctx->mrg, &ctx->test_pos, &ctx->shbuf, &ctx->chain};
  for (DevBuf* b : bufs) b->release();
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->small_pin) cudaFreeHost(ctx->small_pin);
  if (ctx->sgraph.exec) cudaGraphExecDestroy(ctx->sgraph.exec);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (int i = 0; i < 4; i++) if (ctx->pev[i]) cudaEventDestroy(ctx->pev[i]);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return GPAR_OK;
}

const char* gpar_last_error(const gpar_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gpar_last_timing(const gpar_ctx* ctx, double* device_ms, int64_t* kernel_launches) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (device_ms) *device_ms = ctx->last_ms;
  if (kernel_launches) *kernel_launches = ctx->last_launches;
  return GPAR_OK;
}

int gpar_last_profile(const gpar_ctx* ctx, double* phase_ms, int32_t n) {
  if (!ctx || !phase_ms || n < 3) return GPAR_ERR_INVALID;
  for (int i = 0; i < n; i++) phase_ms[i] = 0.0;
  if (!ctx->phase_valid) return GPAR_OK;
  float a = 0, b = 0;
  if (cudaEventElapsedTime(&a, ctx->ev0, ctx->pev[0]) != cudaSuccess) return GPAR_OK;
  if (cudaEventElapsedTime(&b, ctx->pev[1], ctx->pev[2]) != cudaSuccess) return GPAR_OK;
  phase_ms[0] = a; phase_ms[1] = b; phase_ms[2] = ctx->last_ms - a - b;
  return GPAR_OK;
}

int gpar_set_inputs(gpar_ctx* ctx, const double* X, int32_t D, int64_t N) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!X || D < 1 || N < 0) return gpar_fail(ctx, GPAR_ERR_INVALID, "set_inputs: need X != NULL, D >= 1, N >= 0");
  CHK(upload(ctx, ctx->X, X, (size_t)N * D));
  ctx->D = D; ctx->N = N;
  return GPAR_OK;
}
int gpar_set_pseudo(gpar_ctx* ctx, const double* Z, int32_t D, int64_t M) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!Z || D < 1 || M < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "set_pseudo: need Z != NULL, D >= 1, M >= 1");
  CHK(upload(ctx, ctx->Z, Z, (size_t)M * D));
  ctx->Dz = D; ctx->M = M;
  return GPAR_OK;
}
int gpar_set_times(gpar_ctx* ctx, const double* t, int64_t N) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!t || N < 0) return gpar_fail(ctx, GPAR_ERR_INVALID, "set_times: need t != NULL, N >= 0");
  CHK(upload(ctx, ctx->t, t, (size_t)N));
  ctx->Nt = N; ctx->t_reg_dt = 0.0; ctx->merged_Ns = 0;
  return GPAR_OK;
}
__global__ void fill_range_kernel(double* t, double t0, double dt, int64_t N) {
  int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < N) t[k] = fma((double)k, dt, t0);
}
int gpar_set_times_range(gpar_ctx* ctx, double t0, double dt, int64_t N) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (N < 0 || !(dt > 0.0)) return gpar_fail(ctx, GPAR_ERR_INVALID, "set_times_range: need N >= 0 and dt > 0");
  CU(cudaSetDevice(ctx->device));
  CU(ctx->t.reserve((size_t)std::max<int64_t>(N, 1) * sizeof(double)));
  if (N > 0) LAUNCH(ctx, fill_range_kernel, (int)((N + 255) / 256), 256, 0, ctx->t.as<double>(), t0, dt, N);
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->Nt = N; ctx->t_reg_dt = dt; ctx->merged_Ns = 0;
  return GPAR_OK;
}
int gpar_set_outputs(gpar_ctx* ctx, const double* y, int64_t N, int32_t batch) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!y || N < 0 || batch < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "set_outputs: need y != NULL, N >= 0, batch >= 1");
  CHK(upload(ctx, ctx->y, y, (size_t)N * batch));
  ctx->Ny = N; ctx->ybatch = batch;
  return GPAR_OK;
}
int gpar_set_noise_vector(gpar_ctx* ctx, const double* r, int64_t N) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!r) { ctx->has_rvec = false; ctx->Nr = 0; return GPAR_OK; }
  CHK(upload(ctx, ctx->rvec, r, (size_t)N));
  ctx->Nr = N; ctx->has_rvec = true;
  return GPAR_OK;
}

int gpar_dtc_logpdf(gpar_ctx* ctx, int kernel, const double theta[3], int vfe, double jitter, double* val, double* grad) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!theta || !val) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc_logpdf: theta and val must not be NULL");
  if (kernel < GPAR_EQ || kernel > GPAR_MATERN52) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc_logpdf: unknown kernel code %d", kernel);
  if (ctx->N < 1 || ctx->M < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc_logpdf: inputs (set_inputs) and pseudo-inputs (set_pseudo) must be set");
  if (ctx->D != ctx->Dz) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc_logpdf: X has D=%d but Z has D=%d", ctx->D, ctx->Dz);
  if (ctx->Ny != ctx->N || ctx->ybatch < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "dtc_logpdf: outputs length %lld != N %lld", (long long)ctx->Ny, (long long)ctx->N);
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx);
  const GpParams p = unpack_gp3(theta);
  const bool want_grad = grad != nullptr;
  const int64_t N = ctx->N; const int M = (int)ctx->M;
  const int Mpad = (M + GPAR_TILE - 1) / GPAR_TILE * GPAR_TILE;
  const int64_t Npad = (N + GPAR_KT - 1) / GPAR_KT * GPAR_KT;
  const size_t panel_bytes = (size_t)Npad * Mpad * sizeof(double);
  CU(ctx->panelK.reserve(panel_bytes));
  if (want_grad) CU(ctx->panelD.reserve(panel_bytes));
  const int T = Mpad / GPAR_TILE;
  int nsplit = std::max(1, (ctx->num_sms * 16) / T);   // 16 CTAs of 128 threads per SM: full occupancy at 48 registers
  nsplit = (int)std::min<int64_t>(nsplit, std::max<int64_t>(1, Npad / 4));
  CU(ctx->gpart.reserve((size_t)nsplit * 2 * Mpad * sizeof(double) + 1024 * sizeof(double)));
  const size_t MM = (size_t)M * M;
  CU(ctx->scal.reserve(64));
  // G, H, g/h, yy live in kal_a (not used by this entry point otherwise)
  CU(ctx->kal_a.reserve((2 * MM + 2 * (size_t)Mpad + 8) * sizeof(double)));
  double* G = ctx->kal_a.as<double>(); double* H = G + MM; double* gh = H + MM; double* dyy = gh + 2 * Mpad;
  // fork: cov(u), its Cholesky factor and (for gradients) its explicit inverse do not depend on the data
  // statistics — they run on the side stream underneath the producer and the SYRK
  CHK(dtc_tail_prepare(ctx, kernel, p, vfe, jitter, want_grad));
  CHK(launch_kuf_panels(ctx, kernel, want_grad, p.l, p.s, ctx->panelK.as<double>(), ctx->panelD.as<double>(),
                        ctx->gpart.as<double>(), nsplit, Npad, Mpad));
  CHK(launch_reduce_gh(ctx, ctx->gpart.as<double>(), nsplit, Mpad, 2, gh));
  double* ypart = ctx->gpart.as<double>() + (size_t)nsplit * 2 * Mpad;
  LAUNCH(ctx, sumsq_kernel, 1024, 256, 0, ctx->y.as<double>(), N, ypart);
  LAUNCH(ctx, sum_final_kernel, 1, 256, 0, ypart, 1024, dyy);
  // value-only calls: when cov(u) is poorly conditioned, whiten the panel by L_u before the SYRK (A = L_u^-1 Kuf as the
  // reference forms it, dtc_example.jl:14-16) instead of collapsing to Kuf Kfu first — see gpar_needs_whitened_panel
  bool whitened = false, whitened_grad = false;
  {
    TailBufs tb;
    CHK(tail_layout(ctx, want_grad, vfe, &tb));
    double mm[2] = {1.0, 1.0};
    CU(cudaMemcpyAsync(mm, tb.sc + 4, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream2));
    CU(cudaStreamSynchronize(ctx->stream2));              // the panel producer is already running on the main stream
    bool ill = gpar_needs_whitened_panel(mm);
    if (!want_grad && ill) {
      CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
      CHK(panel_left_solve(ctx, ctx->panelK.as<double>(), Npad, Mpad, M, tb.Lu));
      whitened = true;
    }
    if (want_grad) {
      // poorly conditioned cov(u): the collapsed analytic gradient (explicit (cov(u) + G)^-1) loses cond * eps and its Lambda
      // may not even factor — whiten BOTH panels by L_u (A = L_u^-1 Kuf, A_D = L_u^-1 D'), so that the SYRK yields
      // A A' and A A_D' directly, and finish in whitened coordinates (dtc_tail_whitened).
      // Testing knobs: GPAR_GRAD_FD=1 forces the 4-point stencil of the value path, 0 the collapsed analytic form;
      // GPAR_GRAD_WHITENED=1 forces the whitened form.
      bool fd = false; whitened_grad = ill;
      if (const char* e = getenv("GPAR_GRAD_FD")) { fd = atoi(e) != 0; whitened_grad = false; }
      if (const char* e = getenv("GPAR_GRAD_WHITENED")) { if (atoi(e) != 0) { whitened_grad = true; fd = false; } }
      if (fd) {
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
        CHK(gpar_dtc_logpdf(ctx, kernel, theta, vfe, jitter, val, nullptr));
        return gpar_fd_gradient([&](const double* th, double* v) { return gpar_dtc_logpdf(ctx, kernel, th, vfe, jitter, v, nullptr); }, theta, 3, grad);
      }
      if (whitened_grad) {
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
        CHK(panel_left_solve(ctx, ctx->panelK.as<double>(), Npad, Mpad, M, tb.Lu));
        CHK(panel_left_solve(ctx, ctx->panelD.as<double>(), Npad, Mpad, M, tb.Lu));
      }
    }
  }
  cudaEventRecord(ctx->pev[0], ctx->stream);
  CHK(panel_syrk_run(ctx, ctx->panelK.as<double>(), ctx->panelD.as<double>(), Npad, Mpad, M, want_grad, G, H));
  ctx->phase_valid = true;
  double yy = 0.0;
  CU(cudaMemcpyAsync(&yy, dyy, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (whitened_grad) return dtc_tail_whitened(ctx, p, vfe, jitter, N, G, H, gh, gh + Mpad, yy, val, grad, nullptr);
  return dtc_tail(ctx, kernel, p, vfe, jitter, N, G, H, gh, gh + Mpad, yy, val, grad, nullptr, whitened);
}

}  // extern "C"
