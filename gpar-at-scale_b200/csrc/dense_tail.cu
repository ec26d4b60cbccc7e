// dense_tail.cu — the M x M tail of the pseudo-point objectives.
//
// Replaces, for Sigma_y = sigma^2 I (Stheno dtc/elbo as restated at examples/dtc_example.jl:10-23)
//   A = chol(cov(u)).U' \ (Cfu / sigma)',  Lambda = chol(A A' + I),
//   dtc = -1/2 [N log 2pi + N log sigma^2 + logdet Lambda + y'y/sigma^2 - ||Lambda.U' \ (A y/sigma)||^2]
// evaluated from the collapsed statistics G = Kuf Kfu, g = Kuf y, y'y, so no
// N x M array is touched here:  B = L_u^{-1} G L_u^{-T}/sigma^2, Lambda = I + B,
// c = L_Lambda^{-1} L_u^{-1} g / sigma^2.
// The analytic gradient (NEW: the reference has none) uses Q = cov(u) + G/sigma^2 = L_u Lambda L_u',
// P = Q^{-1}, w = P g, T = P + w w'/sigma^4:
//   dF = -1/(2 sigma^2) tr(T dG) + w'dg/sigma^4 - 1/2 tr((T - cov(u)^{-1}) dcov(u)) + direct sigma^2 term
// with dG/dlog l = H + H' from the forward-mode panel (panel_syrk.cu).  DESIGN.md has the derivation;
// tests pin it on torch autograd of the oracle.
// The M^3-class pieces (Cholesky, triangular inverse / solves, M x M products) are the hand-written routines of
// dense_la.cu; the element-wise and reduction pieces are kernels of this file.  No library call anywhere.
#include "common.cuh"
#include <algorithm>

namespace {

// Kj = s kappa(Z,Z) + jitter I  (cov(u), dtc.jl:35,119);  dKu = s * l dkappa/dl
template <int KIND>
__global__ void kuu_kernel(const double* __restrict__ Z, int M, int D, double inv_l2, double s, double jitter,
                           double* __restrict__ Kj, double* __restrict__ dKu) {
  int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (a >= M) return;
  double d2 = 0.0;
  for (int d = 0; d < D; d++) { double df = Z[(int64_t)a * D + d] - Z[(int64_t)b * D + d]; d2 = fma(df, df, d2); }
  double ld; double k = base_kernel_dev<KIND, true>(d2 * inv_l2, ld);
  Kj[(int64_t)a + (int64_t)b * M] = s * k + (a == b ? jitter : 0.0);
  if (dKu) dKu[(int64_t)a + (int64_t)b * M] = s * ld;
}

// out[0] = tr(B); then B += I.   single block.
__global__ void trace_add_identity_kernel(double* B, int M, double* out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) { double v = B[(int64_t)i * M + i]; acc += v; B[(int64_t)i * M + i] = v + 1.0; }
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = r;
}

// out[0] = 2 sum log diag(L)
__global__ void logdet_kernel(const double* L, int M, double* out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) acc += log(L[(int64_t)i * M + i]);
  double r = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = 2.0 * r;
}

constexpr int NTR = GPAR_NTR;
// Element-wise trace sums over the M x M matrices (column-major; all symmetric except H).
// part[block][NTR]; reduced in fixed order by trace_final_kernel.
__global__ void __launch_bounds__(256)
trace_kernel(int M, const double* __restrict__ P, const double* __restrict__ Kinv, const double* __restrict__ G,
             const double* __restrict__ H, const double* __restrict__ Kj, const double* __restrict__ dKu,
             const double* __restrict__ Cm, const double* __restrict__ w, const double* __restrict__ g,
             const double* __restrict__ h, double* __restrict__ part) {
  __shared__ double sh[32];
  double acc[NTR];
#pragma unroll
  for (int i = 0; i < NTR; i++) acc[i] = 0.0;
  const int64_t total = (int64_t)M * M;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int a = (int)(e % M), b = (int)(e / M);
    double p = P[e], ki = Kinv[e], gg = G[e], hh = H[e], kj = Kj[e], dk = dKu[e];
    double ww = w[a] * w[b];
    acc[0] += p * hh;  acc[1] += ww * hh;
    acc[2] += p * gg;  acc[3] += ww * gg;
    acc[4] += p * dk;  acc[5] += ww * dk;  acc[6] += ki * dk;
    acc[7] += p * kj;  acc[8] += ww * kj;  acc[9] += ki * kj;
    acc[13] += ki * hh;
    if (Cm) { double c = Cm[e]; acc[14] += c * dk; if (a == b) acc[15] += c; }
    if (a == b) { acc[10] += p; acc[11] += ki; acc[12] += ww; }
    if (b == 0) { acc[16] += g[a] * w[a]; acc[17] += h[a] * w[a]; }
  }
#pragma unroll
  for (int i = 0; i < NTR; i++) {
    double r = block_sum(acc[i], sh);
    if (threadIdx.x == 0) part[(int64_t)blockIdx.x * NTR + i] = r;
  }
}
__global__ void trace_final_kernel(const double* part, int nblocks, double* out) {
  int i = threadIdx.x;
  if (i >= NTR) return;
  double acc = 0.0;
  for (int b = 0; b < nblocks; b++) acc += part[(int64_t)b * NTR + i];
  out[i] = acc;
}

// ---- whitened-coordinate gradient (ill-conditioned cov(u)) ----------------------------------------
// Q = Lambda^-1 - I + wt wt'  (= A (dF/dA)': the M x M matrix every cov(u)-derivative contracts with; eigenvalues of
// Lambda^-1 - I lie in (-1, 0], so nothing here is larger than the data term)
__global__ void whitened_q_kernel(const double* __restrict__ Li2, const double* __restrict__ wt, int M, double* __restrict__ Q) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * M) return;
  const int a = (int)(e % M), b = (int)(e / M);
  Q[e] = Li2[e] - (a == b ? 1.0 : 0.0) + wt[a] * wt[b];
}
// Element-wise sums of the whitened tail; same partial / final scheme (and slot count) as trace_kernel:
//  [0] tr Q  [1] <W,Q>  [2] <X,Q>  [3] <Lambda^-1,Hw>  [4] wt'Hw wt  [5] wt'hA  [6] tr Hw  [7] <X,GA>  [8] <W,GA>
// with W = L_u^-1 L_u^-T, X = L_u^-1 (l dKuu/dl) L_u^-T; Hw, hA, GA nullable.
__global__ void __launch_bounds__(256)
whitened_trace_kernel(int M, const double* __restrict__ Q, const double* __restrict__ W, const double* __restrict__ X,
                      const double* __restrict__ Li2, const double* __restrict__ Hw, const double* __restrict__ GA,
                      const double* __restrict__ wt, const double* __restrict__ hA, double* __restrict__ part) {
  __shared__ double sh[32];
  double acc[NTR];
#pragma unroll
  for (int i = 0; i < NTR; i++) acc[i] = 0.0;
  const int64_t total = (int64_t)M * M;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(e % M), b = (int)(e / M);
    const double q = Q[e], w = W[e], x = X[e];
    if (a == b) acc[0] += q;
    acc[1] += w * q;  acc[2] += x * q;
    if (Hw) { const double hh = Hw[e]; acc[3] += Li2[e] * hh; acc[4] += wt[a] * wt[b] * hh; if (a == b) acc[6] += hh; }
    if (GA) { const double ga = GA[e]; acc[7] += x * ga; acc[8] += w * ga; }
    if (hA && b == 0) acc[5] += wt[a] * hA[a];
  }
#pragma unroll
  for (int i = 0; i < NTR; i++) {
    double r = block_sum(acc[i], sh);
    if (threadIdx.x == 0) part[(int64_t)blockIdx.x * NTR + i] = r;
  }
}

}  // namespace

static const double LOG2PI = 1.8378770664093454835606594728112;

// Buffer layout of the tail (struct TailBufs, common.cuh), shared by dtc_tail_prepare, dtc_tail and
// the scaled-GPAR gradient (scaled.cu), which reads P, w, ... back from it.
int tail_layout(gpar_ctx* ctx, bool want_grad, int vfe, TailBufs* b) {
  const int M = (int)ctx->M;
  const size_t MM = (size_t)M * M;
  const int nmat = (want_grad && vfe) ? 10 : 8;
  CU(ctx->dense.reserve(nmat * MM * sizeof(double) + 8 * (size_t)M * sizeof(double) + 64 * sizeof(double)));
  double* base = ctx->dense.as<double>();
  b->Kj = base; b->Lu = base + MM; b->Bm = base + 2 * MM;
  b->dKu = want_grad ? base + 3 * MM : nullptr;
  b->V = base + 4 * MM; b->Kinv = base + 5 * MM; b->R = base + 6 * MM; b->Pm = base + 7 * MM;
  b->Tm = base + 8 * MM; b->Cm = base + 9 * MM;
  double* vecs = base + nmat * MM;
  b->cvec = vecs; b->wvec = vecs + M;
  b->sc = vecs + 8 * (size_t)M;   // device scalars: [0]=trB [1]=logdetLam [2]=cc [4,5]=min/max diag L_u [8..8+NTR) traces
  CU(ctx->info.reserve(4 * sizeof(int)));
  b->dinfo = ctx->info.as<int>();
  return GPAR_OK;
}

// G-independent part of the tail, enqueued on the SIDE stream (overlaps the producer and the SYRK):
// cov(u) = Kuu + jitter I (and l dKuu/dl), its Cholesky factor L_u and, for gradients, V = L_u^-1 and
// cov(u)^-1 = V'V.  Records ctx->ev_side.
int dtc_tail_prepare(gpar_ctx* ctx, int kind, const GpParams& p, int vfe, double jitter_in, bool want_grad) {
  const int M = (int)ctx->M, D = ctx->Dz;
  const size_t MM = (size_t)M * M;
  const double jitter = jitter_in < 0.0 ? p.noise : jitter_in;
  TailBufs b;
  CHK(tail_layout(ctx, want_grad, vfe, &b));
  cudaStream_t main_stream = ctx->stream;
  CU(cudaEventRecord(ctx->ev_fork, main_stream));
  CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
  ctx->stream = ctx->stream2;                 // LAUNCH() and the dense routines follow ctx->stream
  int rc = [&]() -> int {
    const double inv_l2 = 1.0 / (p.l * p.l);
    dim3 kgrid((M + 127) / 128, M);
    const double* Zd = ctx->Z.as<double>();
    switch (kind) {
      case GPAR_EQ: LAUNCH(ctx, kuu_kernel<GPAR_EQ>, kgrid, 128, 0, Zd, M, D, inv_l2, p.s, jitter, b.Kj, b.dKu); break;
      case GPAR_MATERN12: LAUNCH(ctx, kuu_kernel<GPAR_MATERN12>, kgrid, 128, 0, Zd, M, D, inv_l2, p.s, jitter, b.Kj, b.dKu); break;
      case GPAR_MATERN32: LAUNCH(ctx, kuu_kernel<GPAR_MATERN32>, kgrid, 128, 0, Zd, M, D, inv_l2, p.s, jitter, b.Kj, b.dKu); break;
      default: LAUNCH(ctx, kuu_kernel<GPAR_MATERN52>, kgrid, 128, 0, Zd, M, D, inv_l2, p.s, jitter, b.Kj, b.dKu); break;
    }
    CU(cudaMemcpyAsync(b.Lu, b.Kj, MM * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CHK(dla_potrf(ctx, M, b.Lu, M, b.dinfo));
    CHK(launch_diag_minmax(ctx, b.Lu, M, b.sc + 4));        // conditioning estimate for the value-only path (abi.cu)
    CHK(dla_trtri(ctx, M, b.Lu, M, b.V, M));                 // V = L_u^-1: B = V G V' below; cov(u)^-1 = V'V for gradients
    if (want_grad) CHK(dla_gemm(ctx, true, false, M, M, M, 1.0, b.V, M, b.V, M, 0.0, b.Kinv, M, DLA_A_UPPER | DLA_B_LOWER));
    return GPAR_OK;
  }();
  cudaEventRecord(ctx->ev_side, ctx->stream2);
  ctx->stream = main_stream;
  return rc;
}

int dtc_tail(gpar_ctx* ctx, int kind, const GpParams& p, int vfe, double jitter_in, int64_t N,
             const double* G, const double* H, const double* g, const double* h, double yy,
             double* val, double* grad, double* raw_out, bool whitened_G) {
  const int M = (int)ctx->M;
  const size_t MM = (size_t)M * M;
  const bool want_grad = grad != nullptr;
  const bool jit_is_noise = jitter_in < 0.0;
  const double jitter = jit_is_noise ? p.noise : jitter_in;
  const double ip = 1.0 / p.noise;
  TailBufs tb;
  CHK(tail_layout(ctx, want_grad, vfe, &tb));
  double *Kj = tb.Kj, *Lu = tb.Lu, *Bm = tb.Bm, *dKu = tb.dKu, *V = tb.V, *Kinv = tb.Kinv, *R = tb.R, *Pm = tb.Pm, *Tm = tb.Tm, *Cm = tb.Cm;
  double *cvec = tb.cvec, *wvec = tb.wvec, *sc = tb.sc;
  int* dinfo = tb.dinfo;
  (void)kind;
  CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));     // join: L_u, V = L_u^-1 (and cov(u)^-1) are ready
  int trace_blocks = 0;
  // B = L_u^-1 G L_u^-T / sigma^2 = V G V' / sigma^2 (lower tiles: only chol(Lambda) reads it); with a panel whitened by
  // L_u before the SYRK, G is already A A' (sigma^2 apart)
  if (whitened_G) {
    CU(cudaMemcpyAsync(Bm, G, MM * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CHK(dla_scal(ctx, (long long)MM, ip, Bm));
  } else {
    CHK(dla_gemm(ctx, false, false, M, M, M, 1.0, V, M, G, M, 0.0, Pm, M, DLA_A_LOWER));                       // Pm = V G (scratch)
    CHK(dla_gemm(ctx, false, true, M, M, M, ip, Pm, M, V, M, 0.0, Bm, M, DLA_B_UPPER | DLA_LOWER_TILES));       // B
  }
  LAUNCH(ctx, trace_add_identity_kernel, 1, 256, 0, Bm, M, sc + 0);
  CHK(dla_potrf(ctx, M, Bm, M, dinfo + 1));
  LAUNCH(ctx, logdet_kernel, 1, 256, 0, Bm, M, sc + 1);
  // c = L_Lambda^-1 L_u^-1 g / sigma^2 (and w below) through backward-stable substitutions, NOT through explicit
  // inverses: the gradient sums cancel terms of size g'w / sigma^4 and need w to satisfy Q w = g to working accuracy
  // (with w = P g the s-derivative at N = 1M, M = 1024 was off by 1 %).
  CU(cudaMemcpyAsync(cvec, g, M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CHK(dla_trsv(ctx, false, M, Lu, M, cvec));
  CHK(dla_trsv(ctx, false, M, Bm, M, cvec, ip));
  if (want_grad) {
    //   R = L_Lambda^-1 V (= (L_u L_Lambda)^-1);  P = Q^-1 = R'R,  Q = cov(u) + G/sigma^2 = L_u Lambda L_u'
    CHK(dla_trtri(ctx, M, Bm, M, Pm, M));                                                                        // Pm = L_Lambda^-1 (scratch)
    CHK(dla_gemm(ctx, false, false, M, M, M, 1.0, Pm, M, V, M, 0.0, R, M, DLA_A_LOWER | DLA_B_LOWER));
    CHK(dla_gemm(ctx, true, false, M, M, M, 1.0, R, M, R, M, 0.0, Pm, M, DLA_A_UPPER | DLA_B_LOWER));
    CU(cudaMemcpyAsync(wvec, cvec, M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CHK(dla_trsv(ctx, true, M, Bm, M, wvec));
    CHK(dla_trsv(ctx, true, M, Lu, M, wvec, p.noise));
  }
  CHK(dla_dot(ctx, M, cvec, cvec, sc + 2));
  if (want_grad) {
    if (vfe) {
      CHK(dla_gemm(ctx, false, false, M, M, M, 1.0, Kinv, M, G, M, 0.0, Tm, M));
      CHK(dla_gemm(ctx, false, false, M, M, M, 1.0, Tm, M, Kinv, M, 0.0, Cm, M));
    }
    trace_blocks = (int)std::min<size_t>((MM + 255) / 256, (size_t)ctx->num_sms * 4);
    CU(ctx->scal.reserve((size_t)trace_blocks * NTR * sizeof(double)));
    // reverse-mode callers (scaled.cu) have no forward-mode H / h: G / g stand in and those sums are ignored
    LAUNCH(ctx, trace_kernel, trace_blocks, 256, 0, M, Pm, Kinv, G, H ? H : G, Kj, dKu, vfe ? Cm : nullptr, wvec, g, h ? h : g, ctx->scal.as<double>());
    LAUNCH(ctx, trace_final_kernel, 1, 32, 0, ctx->scal.as<double>(), trace_blocks, sc + 8);
  }
  double hs[8 + NTR]; int hinfo[2];
  CU(cudaMemcpyAsync(hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hinfo, dinfo, sizeof(hinfo), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (hinfo[0] != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(cov(u)) failed: leading minor %d is not positive definite", hinfo[0]);
  if (hinfo[1] != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(A*A' + I) failed: leading minor %d is not positive definite", hinfo[1]);
  if (raw_out) memcpy(raw_out, hs, sizeof(hs));
  const double trB = hs[0], logdetL = hs[1], cc = hs[2];
  double v = -0.5 * ((double)N * LOG2PI + (double)N * log(p.noise) + logdetL + yy * ip - cc);
  if (vfe) v += -0.5 * ((double)N * p.s * ip - trB);
  *val = v;
  if (want_grad) {
    const double* t = hs + 8;
    const double ip2 = ip * ip;
    const double trTH = t[0] + t[1] * ip2, trTG = t[2] + t[3] * ip2;
    const double trTKdK = t[4] + t[5] * ip2 - t[6];
    const double trTKK = t[7] + t[8] * ip2 - t[9];
    const double trTK = t[10] + t[12] * ip2 - t[11];
    const double gw = t[16], hw = t[17];
    double dlogl = -ip * trTH + ip2 * hw - 0.5 * trTKdK;
    double ds = (-ip * trTG + ip2 * gw - 0.5 * (trTKK - jitter * trTK)) / p.s;
    double dn = -0.5 * ((double)N * ip - yy * ip2 + 2.0 * gw * ip2 * ip - trTG * ip2);
    if (jit_is_noise) dn += -0.5 * trTK;
    if (vfe) {
      const double SKinvH = t[13], SCdK = t[14], trC = t[15];
      dlogl += -0.5 * ip * (-2.0 * SKinvH + SCdK);
      ds += -0.5 * ((double)N * ip - (trB + jitter * ip * trC) / p.s);
      dn += -0.5 * (-(double)N * p.s * ip2 + trB * ip);
      if (jit_is_noise) dn += -0.5 * ip * trC;
    }
    grad[0] = dlogl * p.dl / p.l;
    grad[1] = ds * p.ds_dv;
    grad[2] = dn * p.dn;
  }
  return GPAR_OK;
}

// The tail in WHITENED coordinates, for cov(u) too poorly conditioned for the collapsed statistic.
// Inputs are the statistics of the panels whitened by L_u BEFORE the SYRK: Gw = A A' sigma^2 (A = L_u^-1 Kuf / sigma as
// the reference forms it, dtc.jl:119-120), Hw = L_u^-1 (Kuf D) L_u^-T (forward-mode callers; nullable), and the raw
// g = Kuf y, h = D'y.  With Lambda = I + A A', a = A y / sigma, wt = Lambda^-1 a:
//   dF/dA = -Lambda^-1 A + wt e',  e = y / sigma - A'wt;   dA = L_u^-1 dKuf / sigma - (L_u^-1 dL_u) A;
//   L_u^-1 dL_u = Phi(L_u^-1 dcov(u) L_u^-T)  =>  cov(u)-part of dF = -1/2 <L_u^-1 dcov(u) L_u^-T, Q>,  Q = Lambda^-1 - I + wt wt'
// — every matrix that is formed is conditioned like Lambda; L_u enters through triangular solves and V = L_u^-1 only.
// grad (nullable): the three plain-DTC / VFE derivatives.  out (nullable): what the reverse-mode caller (scaled.cu) needs.
int dtc_tail_whitened(gpar_ctx* ctx, const GpParams& p, int vfe, double jitter_in, int64_t N,
                      const double* Gw, const double* Hw, const double* g, const double* h, double yy,
                      double* val, double* grad, WhitenedTail* out) {
  const int M = (int)ctx->M;
  const size_t MM = (size_t)M * M;
  const bool jit_is_noise = jitter_in < 0.0;
  const double jitter = jit_is_noise ? p.noise : jitter_in;
  const double ip = 1.0 / p.noise;
  TailBufs tb;
  CHK(tail_layout(ctx, true, vfe, &tb));
  double *Lu = tb.Lu, *Bm = tb.Bm, *V = tb.V, *W = tb.Kinv, *Li2 = tb.R, *Q = tb.Pm, *T1 = tb.Kj, *X = tb.dKu;
  double* GA = vfe ? tb.Cm : nullptr;
  double *cvec = tb.cvec, *wvec = tb.wvec, *wt = tb.cvec + 2 * (size_t)M, *hA = tb.cvec + 3 * (size_t)M, *sc = tb.sc;
  int* dinfo = tb.dinfo;
  CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
  CU(cudaMemcpyAsync(Bm, Gw, MM * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (ip != 1.0) CHK(dla_scal(ctx, (long long)MM, ip, Bm));
  if (GA) CU(cudaMemcpyAsync(GA, Bm, MM * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  LAUNCH(ctx, trace_add_identity_kernel, 1, 256, 0, Bm, M, sc + 0);
  CHK(dla_potrf(ctx, M, Bm, M, dinfo + 1));
  LAUNCH(ctx, logdet_kernel, 1, 256, 0, Bm, M, sc + 1);
  CU(cudaMemcpyAsync(cvec, g, M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CHK(dla_trsv(ctx, false, M, Lu, M, cvec));
  CHK(dla_trsv(ctx, false, M, Bm, M, cvec, ip));                       // c = L_Lambda^-1 a
  CHK(dla_dot(ctx, M, cvec, cvec, sc + 2));
  CU(cudaMemcpyAsync(wt, cvec, M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CHK(dla_trsv(ctx, true, M, Bm, M, wt));                              // wt = Lambda^-1 a
  CU(cudaMemcpyAsync(wvec, wt, M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CHK(dla_trsv(ctx, true, M, Lu, M, wvec));                            // w = L_u^-T wt
  if (h) {
    CU(cudaMemcpyAsync(hA, h, M * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    CHK(dla_trsv(ctx, false, M, Lu, M, hA, ip));                       // hA = L_u^-1 h / sigma^2
  }
  CHK(dla_trtri(ctx, M, Bm, M, Q, M));                                                                          // L_Lambda^-1 (scratch)
  CHK(dla_gemm(ctx, true, false, M, M, M, 1.0, Q, M, Q, M, 0.0, Li2, M, DLA_A_UPPER | DLA_B_LOWER));            // Lambda^-1
  LAUNCH(ctx, whitened_q_kernel, (int)((MM + 255) / 256), 256, 0, Li2, wt, M, Q);
  CHK(dla_gemm(ctx, false, true, M, M, M, 1.0, V, M, V, M, 0.0, W, M, DLA_A_LOWER | DLA_B_UPPER));              // W = V V'
  CHK(dla_gemm(ctx, false, false, M, M, M, 1.0, V, M, X, M, 0.0, T1, M, DLA_A_LOWER));                          // V (l dKuu/dl)
  CHK(dla_gemm(ctx, false, true, M, M, M, 1.0, T1, M, V, M, 0.0, X, M, DLA_B_UPPER));                           // X
  const int trace_blocks = (int)std::min<size_t>((MM + 255) / 256, (size_t)ctx->num_sms * 4);
  CU(ctx->scal.reserve((size_t)trace_blocks * NTR * sizeof(double)));
  LAUNCH(ctx, whitened_trace_kernel, trace_blocks, 256, 0, M, Q, W, X, Li2, Hw, GA, wt, h ? hA : (const double*)nullptr, ctx->scal.as<double>());
  LAUNCH(ctx, trace_final_kernel, 1, 32, 0, ctx->scal.as<double>(), trace_blocks, sc + 8);
  // the operand of S = A'(Lambda^-1 L_u^-1): out_n = C in_n with C = V' Lambda^-1 (panel_gemm.cu's orientation)
  if (out) CHK(dla_gemm(ctx, true, false, M, M, M, 1.0, V, M, Li2, M, 0.0, T1, M, DLA_A_UPPER));
  double hs[8 + NTR]; int hinfo[2];
  CU(cudaMemcpyAsync(hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(hinfo, dinfo, sizeof(hinfo), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (hinfo[0] != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(cov(u)) failed: leading minor %d is not positive definite", hinfo[0]);
  if (hinfo[1] != 0) return gpar_fail(ctx, GPAR_ERR_NOT_POSDEF, "cholesky(A*A' + I) failed: leading minor %d is not positive definite", hinfo[1]);
  const double trB = hs[0], logdetL = hs[1], cc = hs[2];
  const double* t = hs + 8;
  double v = -0.5 * ((double)N * LOG2PI + (double)N * log(p.noise) + logdetL + yy * ip - cc);
  if (vfe) v += -0.5 * ((double)N * p.s * ip - trB);
  *val = v;
  if (out) { out->Cop = T1; out->wt = wt; out->w = wvec; out->trQ = t[0]; out->WQ = t[1]; out->XQ = t[2]; out->cc = cc; }
  if (grad) {
    double dlogl = -ip * t[3] + t[5] - ip * t[4] - 0.5 * t[2];
    double ds = 0.5 * (t[0] + jitter * t[1]) / p.s;
    double dn = 0.5 * ip * (-(double)N - t[0] + yy * ip - cc);
    if (jit_is_noise) dn += -0.5 * t[1];
    if (vfe) {
      dlogl += 0.5 * (2.0 * ip * t[6] - t[7]);
      ds += -0.5 * (double)N * ip + 0.5 * (trB + jitter * t[8]) / p.s;
      dn += 0.5 * (double)N * p.s * ip * ip - 0.5 * trB * ip;
      if (jit_is_noise) dn += -0.5 * t[8];
    }
    grad[0] = dlogl * p.dl / p.l;
    grad[1] = ds * p.ds_dv;
    grad[2] = dn * p.dn;
  }
  return GPAR_OK;
}
