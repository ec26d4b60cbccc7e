// kalman.cu — Matern LGSSM Kalman filter / RTS smoother as a temporally-parallel associative scan.
//
// Replaces TemporalGPs' sequential `logpdf(lgssm, y)` (src/gp/temporal_gp_inference.jl:78),
// `decorrelate` (src/gp/dtc.jl:106,115; gpar_scaled_inference.jl:175,183) and `smooth`
// (temporal_gp_inference.jl:109; gpar_scaled_inference.jl:117) for batches of sequences on one
// time grid, each with its own (l, s, sigma^2) or all sharing one.
//
// Chunked three-phase scan (Sarkka & Garcia-Fernandez 2021 elements):
//  P1 kf_chunk_summary : one thread per chunk of L steps builds the chunk's filtering element
//     (A, b, C, eta, J) in registers with a *sequential* pass (a Kalman filter whose mean is affine in
//     the unknown chunk-start state: Phi x0 + b) — ~2x a plain filter step instead of one full
//     element combine per step.
//  P2 scan_up / scan_down : warp-shuffle Kogge-Stone scan of the chunk elements with the full
//     associative combine, 32 elements per warp, recursive over levels (32^k chunks), identity-
//     padded per sequence so no segmentation logic is needed.
//  P3 kf_chunk_filter : one thread per chunk restarts the ordinary filter from the scanned prefix
//     state and emits alpha_k, sum log S_k, sum alpha_k^2 (and, for smoothing, the filtered states
//     and the chunk's smoothing element (E, g, L), composed forward).
//  P4 the same scan machinery over the smoothing elements in reversed time; P5 ks_backward walks
//     each chunk backwards from the scanned smoothed state of the next chunk.
// All state lives in registers; FP64 ALU bound: ~16-24 B of HBM traffic per step.
#include "lgssm_math.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int WARPS_PER_BLOCK = 4;
#ifndef KF_MIN_BLOCKS
#define KF_MIN_BLOCKS 4
#endif
__device__ constexpr double kSmoothJitter = 1e-12;   // TemporalGPs smooth: cholesky(P_pred + 1e-12 I)

struct Level { double* base; int n; int P; };   // n valid elements per sequence, padded to P (multiple of 32)

// An element is stored as Elem::NFD double fields (value and tangent components of each entry).
template <class Elem>
__device__ __forceinline__ void load_elem(Elem& e, const double* base, int64_t fstride, int64_t off) {
  typedef Scalar<typename Elem::scalar_t> SC;
#pragma unroll
  for (int f = 0; f < Elem::NF; f++)
#pragma unroll
    for (int c = 0; c < SC::NC; c++) SC::comp(e.v[f], c) = base[(f * SC::NC + c) * fstride + off];
}
template <class Elem>
__device__ __forceinline__ void store_elem(const Elem& e, double* base, int64_t fstride, int64_t off) {
  typedef Scalar<typename Elem::scalar_t> SC;
#pragma unroll
  for (int f = 0; f < Elem::NF; f++)
#pragma unroll
    for (int c = 0; c < SC::NC; c++) base[(f * SC::NC + c) * fstride + off] = SC::comp(e.v[f], c);
}
template <class Elem>
__device__ __forceinline__ void shfl_up_elem(Elem& o, const Elem& e, int d) {
  typedef Scalar<typename Elem::scalar_t> SC;
#pragma unroll
  for (int f = 0; f < Elem::NF; f++)
#pragma unroll
    for (int c = 0; c < SC::NC; c++) SC::comp(o.v[f], c) = __shfl_up_sync(0xffffffffu, SC::comp(e.v[f], c), d);
}
// parameter with a unit tangent in direction `dir` (none if dir < 0)
template <class F> struct Seed { __device__ __forceinline__ static double make(double x, int) { return x; } };
template <int NT> struct Seed<Dual<NT>> {
  __device__ __forceinline__ static Dual<NT> make(double x, int dir) {
    Dual<NT> r(x);
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = (i == dir) ? 1.0 : 0.0;
    return r;
  }
};

// inclusive prefix (scan order) of element idx of sequence b, from level 0 (local prefixes) and the
// finalised level 1 (if any)
template <class Elem>
__device__ __forceinline__ Elem inclusive_prefix(const Level l0, const Level l1, int batch, int b, int idx) {
  Elem loc; load_elem(loc, l0.base, (int64_t)batch * l0.P, (int64_t)b * l0.P + idx);
  int T = idx >> 5;
  if (T > 0 && l1.base) {
    Elem up; load_elem(up, l1.base, (int64_t)batch * l1.P, (int64_t)b * l1.P + (T - 1));
    return Elem::scan_combine(up, loc);
  }
  return loc;
}

// Up-sweep: one warp per 32-element tile; stores the within-tile inclusive prefixes in place and
// the tile total to the upper level (identity beyond the valid tiles).
template <class Elem>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
scan_up_kernel(Level cur, Level up, int batch) {
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  const int ntiles = cur.P >> 5;
  const int tile_lim = up.base ? up.P : ntiles;
  if (tile >= tile_lim) return;
  if (tile >= ntiles) {   // padding slot of the upper level
    if (lane == 0) { Elem id; id.set_identity(); store_elem(id, up.base, (int64_t)batch * up.P, (int64_t)b * up.P + tile); }
    return;
  }
  Elem e;
  if (tile * 32 + lane < cur.n) load_elem(e, cur.base, (int64_t)batch * cur.P, (int64_t)b * cur.P + tile * 32 + lane);
  else e.set_identity();                  // padding slots need not have been written
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Elem o;
    shfl_up_elem(o, e, d);
    if (lane >= d) e = Elem::scan_combine(o, e);
  }
  store_elem(e, cur.base, (int64_t)batch * cur.P, (int64_t)b * cur.P + tile * 32 + lane);
  if (up.base && lane == 31) store_elem(e, up.base, (int64_t)batch * up.P, (int64_t)b * up.P + tile);
}

// Down-sweep for levels >= 1: element i of tile T > 0 becomes carry(T-1) o local(i).
template <class Elem>
__global__ void scan_down_kernel(Level cur, Level up, int batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= cur.P || i < 32) return;
  Elem loc; load_elem(loc, cur.base, (int64_t)batch * cur.P, (int64_t)b * cur.P + i);
  Elem c; load_elem(c, up.base, (int64_t)batch * up.P, (int64_t)b * up.P + ((i >> 5) - 1));
  Elem r = Elem::scan_combine(c, loc);
  store_elem(r, cur.base, (int64_t)batch * cur.P, (int64_t)b * cur.P + i);
}

// Block-level scan with a running carry: block (x, b) owns elements [x span, (x+1) span) of sequence b and walks them
// in rounds of NW*32*K.  Every lane owns K consecutive elements: it combines them serially, the lane totals go through a
// warp-shuffle scan, the NW warp totals are scanned by warp 0 through shared memory (with the carry of the earlier
// rounds), and the lane walks its K elements again from its exclusive prefix.  Per lane that is 2K + 6 + log2(NW)
// dependent combines — the scans of this file are latency-bound (a FiltElem<3> combine is ~300 dependent-ish FP64
// instructions), so depth, not work, is what is minimised.  Writes the block-local inclusive prefixes in place and the
// block total to `up` (if any).  With span >= cur.n one block finishes a whole sequence in one launch; two launches
// cover 10M-step sequences that the 32-ary scan_up / scan_down pyramid needs six for.
template <class Elem, int NW>
__global__ void __launch_bounds__(NW * 32)
scan_span_kernel(Level cur, Level up, int batch, int span, int K) {
  constexpr int NFD = Elem::NFD;
  __shared__ double sh[(NW + 1) * NFD];             // slot w < NW: warp totals of the round; slot NW: carry of the earlier rounds
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.y;
  const int start = blockIdx.x * span, end = min(start + span, cur.n);
  const int64_t fstride = (int64_t)batch * cur.P, off0 = (int64_t)b * cur.P;
  if (threadIdx.x == 0) { Elem id; id.set_identity(); store_elem(id, sh, NW + 1, NW); }
  pdl_trigger();
  pdl_wait();                                      // (launched as a programmatic dependent of the pass that writes `cur`)
  __syncthreads();
  for (int base = start; base < end; base += NW * 32 * K) {
    const int my0 = base + threadIdx.x * K, my1 = min(my0 + K, end);
    Elem tot; tot.set_identity();
#pragma unroll 1
    for (int idx = my0; idx < my1; idx++) {
      Elem e; load_elem(e, cur.base, fstride, off0 + idx);
      tot = (idx == my0) ? e : Elem::scan_combine(tot, e);
    }
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
      Elem o;
      shfl_up_elem(o, tot, d);
      if (lane >= d) tot = Elem::scan_combine(o, tot);
    }
    if (lane == 31) store_elem(tot, sh, NW + 1, w);
    __syncthreads();
    if (w == 0) {                                   // sh[w] <- carry o (warp totals 0..w)
      Elem tt;
      if (lane < NW) load_elem(tt, sh, NW + 1, lane); else tt.set_identity();
#pragma unroll 1
      for (int d = 1; d < NW; d <<= 1) {
        Elem o;
        shfl_up_elem(o, tt, d);
        if (lane >= d) tt = Elem::scan_combine(o, tt);
      }
      Elem cy; load_elem(cy, sh, NW + 1, NW);
      tt = Elem::scan_combine(cy, tt);
      __syncwarp();
      if (lane < NW) store_elem(tt, sh, NW + 1, lane);
    }
    __syncthreads();
    // exclusive prefix of this lane: (everything before its warp) o (lane totals before it)
    Elem p; shfl_up_elem(p, tot, 1);
    if (lane == 0) p.set_identity();
    { Elem wp; load_elem(wp, sh, NW + 1, w > 0 ? w - 1 : NW); p = Elem::scan_combine(wp, p); }
#pragma unroll 1
    for (int idx = my0; idx < my1; idx++) {
      Elem e; load_elem(e, cur.base, fstride, off0 + idx);
      p = Elem::scan_combine(p, e);
      store_elem(p, cur.base, fstride, off0 + idx);
      if (up.base && idx == end - 1) store_elem(p, up.base, (int64_t)batch * up.P, (int64_t)b * up.P + blockIdx.x);
    }
    __syncthreads();
    if (threadIdx.x < NFD) sh[threadIdx.x * (NW + 1) + NW] = sh[threadIdx.x * (NW + 1) + NW - 1];      // next round's carry
    __syncthreads();
  }
}
// inclusive prefix of element idx after scan_span passes: level 0 holds block-local prefixes over `span` elements,
// level 1 (if any) the finished prefixes of the block totals
template <class Elem>
__device__ __forceinline__ Elem inclusive_prefix_span(const Level l0, const Level l1, int span, int batch, int b, int idx) {
  Elem loc; load_elem(loc, l0.base, (int64_t)batch * l0.P, (int64_t)b * l0.P + idx);
  const int T = idx / span;
  if (T > 0 && l1.base) {
    Elem upe; load_elem(upe, l1.base, (int64_t)batch * l1.P, (int64_t)b * l1.P + (T - 1));
    return Elem::scan_combine(upe, loc);
  }
  return loc;
}

// dir_*: tangent slot of each parameter in the Dual instantiations (-1: not differentiated)
// reg_dt > 0: regular time grid with that spacing (gpar_set_times_range) — the transition matrix is
// then constant per sequence (what TemporalGPs does for a `range` input) and is hoisted out of the loops.
struct SeqParams { const double *l, *s, *noise; int nparam; int dir_l, dir_s, dir_n; double reg_dt; };

// inputs of one step (time, observation, per-step noise), loaded ahead of use
struct StepIn {
  double t = 0.0, y = 0.0, r = 0.0;
  __device__ __forceinline__ void load(const double* __restrict__ tp, const double* __restrict__ yp, const double* __restrict__ rp,
                                       int64_t k, int64_t N, bool reg) {
    if (k < N) { y = __ldg(yp + k); if (!reg) t = __ldg(tp + k); if (rp) r = __ldg(rp + k); }
  }
  // the same without a branch around the loads: indices beyond the sequence re-read its last step
  __device__ __forceinline__ void load_clamped(const double* __restrict__ tp, const double* __restrict__ yp, const double* __restrict__ rp,
                                               int64_t k, int64_t N, bool reg) {
    k = k < N ? k : N - 1;
    y = __ldg(yp + k); if (!reg) t = __ldg(tp + k); if (rp) r = __ldg(rp + k);
  }
};

// P1: chunk filtering element.
template <int D, class F>
__global__ void __launch_bounds__(128, Scalar<F>::NC == 1 ? KF_MIN_BLOCKS : 1)
kf_chunk_summary_kernel(const double* __restrict__ t, const double* __restrict__ y, const double* __restrict__ rvec,
                        SeqParams sp, int64_t N, int L, int nC, Level l0, int batch, int64_t ystride) {
  typedef FiltElem<D, F> E;
  // threads are numbered over (sequence, chunk) jointly: no lane idles when the chunks of a sequence do not fill whole blocks
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= (int64_t)batch * l0.P) return;
  const int b = (int)(gidx / l0.P), c = (int)(gidx % l0.P);
  E e;
  if (c >= nC) { e.set_identity(); store_elem(e, l0.base, (int64_t)batch * l0.P, (int64_t)b * l0.P + c); return; }
  const int pb = sp.nparam == 1 ? 0 : b;
  const F il = 1.0 / Seed<F>::make(sp.l[pb], sp.dir_l), s = Seed<F>::make(sp.s[pb], sp.dir_s), noise = Seed<F>::make(sp.noise[pb], sp.dir_n);
  F P0[NSYM<D>]; lgssm_pinf<D>(P0);
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) P0[i] = P0[i] * s;
  F* Phi = e.v; F* bv = e.v + E::OB; F* C = e.v + E::OC; F* eta = e.v + E::OE; F* J = e.v + E::OJ;
  e.set_identity();
  if (c == 0) {
#pragma unroll
    for (int i = 0; i < D * D; i++) Phi[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) C[i] = P0[i];
  }
  const int64_t k0 = (int64_t)c * L, k1 = (k0 + L < N) ? k0 + L : N;
  double tprev = (k0 == 0) ? __ldg(t) - 1.0 : __ldg(t + k0 - 1);
  const double* yb = y + (int64_t)b * ystride;
  const bool reg = sp.reg_dt > 0.0;
  F A[D * D];
  if (reg) lgssm_transition<D>(sp.reg_dt * il, A);
  // each thread walks its own chunk, so its loads are strided across the warp: they are issued two steps
  // ahead of their use (ncu: long_scoreboard was the top stall of these kernels)
  StepIn in0, in1;
  in0.load(t, yb, rvec, k0, N, reg); in1.load(t, yb, rvec, k0 + 1, N, reg);
  for (int64_t k = k0; k < k1; k++) {
    F T[D * D], u[D], Cn[NSYM<D>];
    const StepIn cur = in0; in0 = in1; in1.load(t, yb, rvec, k + 2, N, reg);
    if (!reg) { lgssm_transition<D>((cur.t - tprev) * il, A); tprev = cur.t; }
    else if (k <= 1) lgssm_transition<D>((k == 0 ? 1.0 : sp.reg_dt) * il, A);     // step 0 follows the t[0] - 1 prefix
    matmul<D>(A, Phi, T);
#pragma unroll
    for (int i = 0; i < D * D; i++) Phi[i] = T[i];
    matvec<D>(A, bv, u);
#pragma unroll
    for (int i = 0; i < D; i++) bv[i] = u[i];
    predict_cov<D>(A, C, P0, Cn);
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) C[i] = Cn[i];
    const F S = rvec ? C[0] + cur.r : C[0] + noise;
    const F iS = 1.0 / S;
    const F r = cur.y - bv[0];
    F h[D], Kg[D];
#pragma unroll
    for (int i = 0; i < D; i++) { h[i] = Phi[i]; Kg[i] = SYM(C, i, 0) * iS; }
#pragma unroll
    for (int i = 0; i < D; i++) {
      eta[i] = fma(iS * r, h[i], eta[i]);
#pragma unroll
      for (int j = i; j < D; j++) SYM(J, i, j) = fma(iS * h[i], h[j], SYM(J, i, j));
    }
#pragma unroll
    for (int i = 0; i < D; i++) {
#pragma unroll
      for (int j = 0; j < D; j++) Phi[i * D + j] = fma(-Kg[i], h[j], Phi[i * D + j]);
      bv[i] = fma(Kg[i], r, bv[i]);
    }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) SYM(C, i, j) = fma(-S * Kg[i], Kg[j], SYM(C, i, j));
  }
  store_elem(e, l0.base, (int64_t)batch * l0.P, (int64_t)b * l0.P + c);
}

// P3: restart the ordinary filter from the scanned prefix state.
// part: [b][c][2 NC] = (sum log S, sum alpha^2), each with its tangents.  SMOOTH additionally stores
// the filtered states fs (f < D + NSYM, warp-coalesced) and the chunk smoothing element at reversed
// index.  Dual instantiations also emit d alpha (dalpha[(j*batch + b)*N + k]) and the tangent rows of
// the step table: dtable row k = [sqrt(S_k), 0, then per tangent: dPhi (D*D), dK (D), dHA (D), dlog rs].
template <int D, bool SMOOTH, class F>
__global__ void __launch_bounds__(128, Scalar<F>::NC == 1 ? (SMOOTH ? 2 : KF_MIN_BLOCKS) : 1)
kf_chunk_filter_kernel(const double* __restrict__ t, const double* __restrict__ y, const double* __restrict__ rvec,
                       SeqParams sp, int64_t N, int L, int nC, Level l0, Level l1, int batch,
                       double* __restrict__ alpha, double* __restrict__ part, double* __restrict__ fs, Level s0,
                       double* __restrict__ table, double* __restrict__ dalpha, double* __restrict__ dtable,
                       int64_t ystride, double* __restrict__ fstate, int64_t astride, int64_t tstride) {
  typedef Scalar<F> SC;
  constexpr int NC = SC::NC;
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (sequence, chunk) jointly, as in the summary pass
  if (gidx >= (int64_t)batch * nC) return;
  const int b = (int)(gidx / nC), c = (int)(gidx % nC);
  const int pb = sp.nparam == 1 ? 0 : b;
  const F il = 1.0 / Seed<F>::make(sp.l[pb], sp.dir_l), s = Seed<F>::make(sp.s[pb], sp.dir_s), noise = Seed<F>::make(sp.noise[pb], sp.dir_n);
  F P0[NSYM<D>]; lgssm_pinf<D>(P0);
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) P0[i] = P0[i] * s;
  F m[D], P[NSYM<D>];
  if (c == 0) {
#pragma unroll
    for (int i = 0; i < D; i++) m[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) P[i] = P0[i];
  } else {
    FiltElem<D, F> pre = inclusive_prefix<FiltElem<D, F>>(l0, l1, batch, b, c - 1);
#pragma unroll
    for (int i = 0; i < D; i++) m[i] = pre.v[FiltElem<D, F>::OB + i];
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) P[i] = pre.v[FiltElem<D, F>::OC + i];
  }
  SmoothElem<D, F> comp; if (SMOOTH) comp.set_identity();
  const int64_t k0 = (int64_t)c * L, k1 = (k0 + L < N) ? k0 + L : N;
  double tprev = (k0 == 0) ? __ldg(t) - 1.0 : __ldg(t + k0 - 1);
  const double* yb = y + (int64_t)b * ystride;
  F sum_logS = 0.0, sum_a2 = 0.0, prodS = 1.0;
  const int64_t kend = SMOOTH ? k1 + 1 : k1;    // one extra predict closes the chunk's last smoothing element
  const bool reg = sp.reg_dt > 0.0;
  F A[D * D];
  if (reg) lgssm_transition<D>(sp.reg_dt * il, A);
  StepIn in0, in1;
  in0.load(t, yb, rvec, k0, N, reg); in1.load(t, yb, rvec, k0 + 1, N, reg);
  for (int64_t k = k0; k < kend; k++) {
    F mp[D], Pp[NSYM<D>];
    const StepIn cur = in0; in0 = in1; in1.load(t, yb, rvec, k + 2, N, reg);
    if (k < N) {
      if (!reg) { lgssm_transition<D>((cur.t - tprev) * il, A); tprev = cur.t; }
      else if (k <= 1) lgssm_transition<D>((k == 0 ? 1.0 : sp.reg_dt) * il, A);   // step 0 follows the t[0] - 1 prefix
      matvec<D>(A, m, mp);
      predict_cov<D>(A, P, P0, Pp);
    }
    if constexpr (SMOOTH) {
      if (k > k0) {
        // smoothing element of step k-1: G = P A^T (Pp + eps I)^{-1}, g = m - G mp, Ls = P - G Pp G^T
        SmoothElem<D, F> el;
        F* G = el.v; F* g = el.v + SmoothElem<D, F>::OG; F* Ls = el.v + SmoothElem<D, F>::OL;
        if (k < N) {
          F W[D * D];
#pragma unroll
          for (int i = 0; i < D; i++)
#pragma unroll
            for (int j = 0; j < D; j++) { F v = 0.0;
#pragma unroll
              for (int q = 0; q < D; q++) v = fma(SYM(P, i, q), A[j * D + q], v);
              W[i * D + j] = v; }
          solve_spd_right<D>(W, Pp, kSmoothJitter, G);
          F u[D], R[NSYM<D>];
          matvec<D>(G, mp, u);
#pragma unroll
          for (int i = 0; i < D; i++) g[i] = m[i] - u[i];
          asat<D>(G, Pp, R);
#pragma unroll
          for (int i = 0; i < NSYM<D>; i++) Ls[i] = P[i] - R[i];
        } else {   // last step of the sequence: (0, m_N, P_N)
#pragma unroll
          for (int i = 0; i < D * D; i++) G[i] = 0.0;
#pragma unroll
          for (int i = 0; i < D; i++) g[i] = m[i];
#pragma unroll
          for (int i = 0; i < NSYM<D>; i++) Ls[i] = P[i];
        }
        comp = SmoothElem<D, F>::combine(comp, el);
      }
    }
    if (k >= k1) break;
    const F S = rvec ? Pp[0] + cur.r : Pp[0] + noise;
    const F rs = rsqrt(S);          // one reciprocal square root instead of sqrt + D+1 divisions
    const F a = (cur.y - mp[0]) * rs;
    F Bv[D];
#pragma unroll
    for (int i = 0; i < D; i++) { Bv[i] = SYM(Pp, 0, i) * rs; m[i] = fma(Bv[i], a, mp[i]); }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) SYM(P, i, j) = fma(-Bv[i], Bv[j], SYM(Pp, i, j));
    // sum log S_k as the log of a running product, flushed every 8 steps (S in [1e-12, 1e10]: no over/underflow)
    prodS = prodS * S;
    if (((k - k0) & 7) == 7) { sum_logS += log(prodS); prodS = 1.0; }
    sum_a2 = fma(a, a, sum_a2);
    if (alpha) alpha[(int64_t)b * astride + k] = value_of(a);      // (astride != ystride when the sequences share one y: candidates)
    if constexpr (NC > 1) {
      if (dalpha) {
#pragma unroll
        for (int j = 1; j < NC; j++) dalpha[((int64_t)(j - 1) * batch + b) * N + k] = SC::comp(a, j);
      }
    }
    {
      if (table) {   // shared-model step table for the affine mean scans (scaled.cu, smooth_shared.cu)
        constexpr int TS = D * D + 2 * D + 1;
        double* row = table + ((int64_t)b * tstride + k) * TS;      // tstride = N: one table per sequence (candidates); 0: batch == 1
        double* drow = nullptr;
        if constexpr (NC > 1) { if (dtable) drow = dtable + k * (2 + (NC - 1) * TS); }
        const F iS = 1.0 / S;
        F ent[TS];
#pragma unroll
        for (int i = 0; i < D; i++) {
          const F Kg = SYM(Pp, i, 0) * iS;
#pragma unroll
          for (int j = 0; j < D; j++) ent[i * D + j] = fma(-Kg, A[j], A[i * D + j]);
          ent[D * D + i] = Kg;
          ent[D * D + D + i] = A[i];
        }
        ent[D * D + 2 * D] = rs;
#pragma unroll
        for (int i = 0; i < TS; i++) row[i] = value_of(ent[i]);
        if constexpr (NC > 1) {
          if (drow) {
            drow[0] = value_of(S) * value_of(rs);       // sqrt(S): turns beta back into the innovation
            drow[1] = 0.0;                              // (pad: keeps the tangent rows 16-byte aligned)
            const double irs = 1.0 / value_of(rs);
#pragma unroll
            for (int j = 1; j < NC; j++) {
#pragma unroll
              for (int i = 0; i < TS - 1; i++) drow[2 + (j - 1) * TS + i] = SC::comp(ent[i], j);
              drow[2 + (j - 1) * TS + TS - 1] = SC::comp(rs, j) * irs;   // d log rs
            }
          }
        }
      }
    }
    if constexpr (SMOOTH) {   // warp-coalesced layout: [sequence][group of 32 chunks][field][step in chunk][chunk % 32]
      double* fsw = fs + (((int64_t)b * ((nC + 31) >> 5) + (c >> 5)) * (D + NSYM<D>) * L + (k - k0)) * 32 + (c & 31);
#pragma unroll
      for (int i = 0; i < D; i++) fsw[(int64_t)i * L * 32] = value_of(m[i]);
#pragma unroll
      for (int i = 0; i < NSYM<D>; i++) fsw[(int64_t)(D + i) * L * 32] = value_of(P[i]);
    }
  }
  sum_logS += log(prodS);
#pragma unroll
  for (int j = 0; j < NC; j++) {
    part[((int64_t)b * nC + c) * 2 * NC + j] = SC::comp(sum_logS, j);
    part[((int64_t)b * nC + c) * 2 * NC + NC + j] = SC::comp(sum_a2, j);
  }
  if constexpr (SMOOTH) store_elem(comp, s0.base, (int64_t)batch * s0.P, (int64_t)b * s0.P + (nC - 1 - c));
  if constexpr (NC == 1 && !SMOOTH) {
    if (fstate && c == nC - 1) {   // filtered state after the last step: hand-over to the steady-state path
#pragma unroll
      for (int i = 0; i < D; i++) fstate[(int64_t)b * (D + NSYM<D>) + i] = m[i];
#pragma unroll
      for (int i = 0; i < NSYM<D>; i++) fstate[(int64_t)b * (D + NSYM<D>) + D + i] = P[i];
    }
  }
}

// pads the reversed smoothing level 0 beyond nC with identities
template <int D>
__global__ void smooth_pad_kernel(Level s0, int nC, int batch) {
  const int i = nC + blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= s0.P) return;
  SmoothElem<D> id; id.set_identity();
  store_elem(id, s0.base, (int64_t)batch * s0.P, (int64_t)b * s0.P + i);
}

// P5: backward walk inside each chunk from the smoothed state at the start of the next chunk.
template <int D>
__global__ void __launch_bounds__(128)
ks_backward_kernel(const double* __restrict__ t, SeqParams sp, int64_t N, int L, int nC, Level s0, Level s1, int batch,
                   const double* __restrict__ fs, double* __restrict__ mean, double* __restrict__ var, double* __restrict__ table2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= nC) return;
  const int pb = sp.nparam == 1 ? 0 : b;
  const double il = 1.0 / sp.l[pb], s = sp.s[pb];
  double P0[NSYM<D>]; lgssm_pinf<D>(P0);
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) P0[i] *= s;
  const int64_t k0 = (int64_t)c * L, k1 = (k0 + L < N) ? k0 + L : N;
  double ms[D], Ps[NSYM<D>];
  int64_t k = k1 - 1;
  const double* fsb = fs + (((int64_t)b * ((nC + 31) >> 5) + (c >> 5)) * (D + NSYM<D>) * L) * 32 + (c & 31);
  if (c == nC - 1) {
#pragma unroll
    for (int i = 0; i < D; i++) ms[i] = fsb[((int64_t)i * L + (k - k0)) * 32];
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) Ps[i] = fsb[((int64_t)(D + i) * L + (k - k0)) * 32];
    mean[(int64_t)b * N + k] = ms[0]; var[(int64_t)b * N + k] = Ps[0];
    if (table2) {      // last step: m^s_N = m_N  (B = I, G = 0)
#pragma unroll
      for (int i = 0; i < D * D; i++) { table2[k * 2 * D * D + i] = (i / D == i % D) ? 1.0 : 0.0; table2[k * 2 * D * D + D * D + i] = 0.0; }
    }
    k--;
  } else {
    SmoothElem<D> suf = inclusive_prefix<SmoothElem<D>>(s0, s1, batch, b, nC - 2 - c);
#pragma unroll
    for (int i = 0; i < D; i++) ms[i] = suf.v[SmoothElem<D>::OG + i];
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) Ps[i] = suf.v[SmoothElem<D>::OL + i];
  }
  for (; k >= k0; k--) {
    double m[D], P[NSYM<D>], A[D * D], mp[D], Pp[NSYM<D>], W[D * D], G[D * D];
#pragma unroll
    for (int i = 0; i < D; i++) m[i] = fsb[((int64_t)i * L + (k - k0)) * 32];
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) P[i] = fsb[((int64_t)(D + i) * L + (k - k0)) * 32];
    lgssm_transition<D>((sp.reg_dt > 0.0 ? sp.reg_dt : __ldg(t + k + 1) - __ldg(t + k)) * il, A);
    matvec<D>(A, m, mp);
    predict_cov<D>(A, P, P0, Pp);
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = 0; j < D; j++) { double v = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) v = fma(SYM(P, i, q), A[j * D + q], v);
        W[i * D + j] = v; }
    solve_spd_right<D>(W, Pp, kSmoothJitter, G);
    if (table2) {      // shared-model smoothing of further sequences: m^s_k = B_k m_k + G_k m^s_{k+1},  B_k = I - G_k A_{k+1}
      double GA[D * D];
      matmul<D>(G, A, GA);
#pragma unroll
      for (int i = 0; i < D * D; i++) { table2[k * 2 * D * D + i] = ((i / D == i % D) ? 1.0 : 0.0) - GA[i]; table2[k * 2 * D * D + D * D + i] = G[i]; }
    }
    double dm[D], dP[NSYM<D>], u[D], R[NSYM<D>];
#pragma unroll
    for (int i = 0; i < D; i++) dm[i] = ms[i] - mp[i];
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) dP[i] = Ps[i] - Pp[i];
    matvec<D>(G, dm, u);
    asat<D>(G, dP, R);
#pragma unroll
    for (int i = 0; i < D; i++) ms[i] = m[i] + u[i];
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) Ps[i] = P[i] + R[i];
    mean[(int64_t)b * N + k] = ms[0]; var[(int64_t)b * N + k] = Ps[0];
  }
}

// lml[b] = -1/2 (N log 2pi + sum log S + sum alpha^2), fixed-order two-stage reduction:
// stage 1: block (s, b) sums a contiguous slice of sequence b's chunk partials; stage 2: one block per
// sequence sums the slices (a single long sequence would otherwise be reduced by one block).
// NC = 1 + number of tangents; per chunk the partials are [sum log S (NC), sum alpha^2 (NC)].
template <int NC>
__global__ void __launch_bounds__(256)
lml_partial_kernel(const double* __restrict__ part, int nC, int nslice, double* __restrict__ part2) {
  __shared__ double sh[32];
  constexpr int NP = 2 * NC;
  const int b = blockIdx.y, s = blockIdx.x;
  const int per = (nC + nslice - 1) / nslice, c0 = s * per, c1 = min(nC, c0 + per);
  double a[NP];
#pragma unroll
  for (int i = 0; i < NP; i++) a[i] = 0.0;
  for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
#pragma unroll
    for (int i = 0; i < NP; i++) a[i] += part[((int64_t)b * nC + c) * NP + i];
  }
#pragma unroll
  for (int i = 0; i < NP; i++) {
    double r = block_sum(a[i], sh);
    if (threadIdx.x == 0) part2[((int64_t)b * nslice + s) * NP + i] = r;
  }
}
// lml[b]; dlml[b*(NC-1) + j] = d lml / d tangent j; sums[b*2NC + ...] = the raw sums.
template <int NC>
__global__ void __launch_bounds__(64)
lml_reduce_kernel(const double* __restrict__ part2, int nslice, int64_t N, double* __restrict__ lml, double* __restrict__ dlml,
                  double* __restrict__ sums) {
  __shared__ double sh[32];
  constexpr int NP = 2 * NC;
  const int b = blockIdx.x;
  double a[NP];
#pragma unroll
  for (int i = 0; i < NP; i++) a[i] = 0.0;
  for (int s = threadIdx.x; s < nslice; s += blockDim.x) {
#pragma unroll
    for (int i = 0; i < NP; i++) a[i] += part2[((int64_t)b * nslice + s) * NP + i];
  }
  double r[NP];
#pragma unroll
  for (int i = 0; i < NP; i++) r[i] = block_sum(a[i], sh);
  if (threadIdx.x == 0) {
    if (lml) lml[b] = -0.5 * ((double)N * 1.8378770664093454835606594728112 + r[0] + r[NC]);
    if (dlml) {
#pragma unroll
      for (int j = 1; j < NC; j++) dlml[b * (NC - 1) + (j - 1)] = -0.5 * (r[j] + r[NC + j]);
    }
    if (sums) {
#pragma unroll
      for (int i = 0; i < NP; i++) sums[b * NP + i] = r[i];
    }
  }
}

// ------------------------------------------------------------------------------------------
struct LevelPlan { std::vector<Level> lv; size_t doubles = 0; };
LevelPlan plan_levels(int n0, int NF, int batch) {
  LevelPlan p;
  int n = n0;
  for (;;) {
    int P = (n + 31) / 32 * 32;
    p.lv.push_back(Level{nullptr, n, P});
    p.doubles += (size_t)NF * batch * P;
    if (P <= 32) break;
    n = P / 32;
  }
  return p;
}
void bind_levels(LevelPlan& p, double* base, int NF, int batch) {
  for (auto& l : p.lv) { l.base = base; base += (size_t)NF * batch * l.P; }
}

template <class Elem>
int run_scan(gpar_ctx* ctx, LevelPlan& p, int batch) {
  const int nl = (int)p.lv.size();
  for (int l = 0; l < nl; l++) {
    Level up = (l + 1 < nl) ? p.lv[l + 1] : Level{nullptr, 0, 0};
    int tiles = up.base ? up.P : p.lv[l].P / 32;
    dim3 grid((tiles + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, batch);
    LAUNCH(ctx, scan_up_kernel<Elem>, grid, WARPS_PER_BLOCK * 32, 0, p.lv[l], up, batch);
  }
  for (int l = nl - 2; l >= 1; l--) {
    dim3 grid((p.lv[l].P + 127) / 128, batch);
    LAUNCH(ctx, scan_down_kernel<Elem>, grid, 128, 0, p.lv[l], p.lv[l + 1], batch);
  }
  return GPAR_OK;
}

#ifndef KF1_PD
#define KF1_PD 4          // steps between the issue of a cp.async group and its use
#endif
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
// predicated form: no branch around the copy (a branch would split the loop body the scheduler works on)
__device__ __forceinline__ void cp_async8_if(double* smem, const double* gmem, bool pred) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %2, 0; @p cp.async.ca.shared.global [%0], [%1], 8; }" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"((int)pred));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

// ---- log-pdf only: ONE pass over the data ------------------------------------------------------------
// The value temporal_gp_inference.jl:78 needs is a sum over steps, and a chunk's share of it is a closed
// form of the chunk's filtering element and the state the chunk starts from, so the second walk over the
// data (P3) is not needed.  Conditional on the start state x0 the chunk's observations have
//   log p(y_c | x0) = -1/2 [ sum_k log 2 pi S0_k + sum_k (v0_k - h_k' x0)^2 / S0_k ]
//                   = -1/2 (n_c log 2 pi + sum log S0 + sum v0^2/S0) + eta' x0 - 1/2 x0' J x0,
// where (S0, v0) are the innovation variances / innovations of the filter started from x0 = 0 EXACTLY
// known and (eta, J) are the element's information pair — everything P1 accumulates anyway.  With
// x0 ~ N(m, P) from the scanned prefix and P = Lc Lc', B = I + Lc' J Lc, a = Lc'(eta - J m):
//   log p(y_c | y_<c) = log p(y_c | x0 = m) + 1/2 a' B^-1 a - 1/2 log det B.
// B has eigenvalues >= 1, so its Cholesky factorisation is unconditionally stable.  The two sums the
// callers use (sum log S_k and sum alpha_k^2 of the real filter) are recovered exactly: the determinant
// and the quadratic form of the chunk's marginal covariance factor both ways.
// kf_chunk_element: P1 in deviation form (dC = C - P_inf: one congruence per step, no Q), threads numbered
// densely over (sequence, chunk); aux rows (field-major, stride batch * nC): sum log S0, sum v0^2/S0, eta, J.
template <int D, class F, int TPB, int MINB, bool REG, int PD>
__global__ void __launch_bounds__(TPB, MINB)
kf_chunk_element_kernel(const double* __restrict__ t, const double* __restrict__ y, const double* __restrict__ rvec,
                        SeqParams sp, int64_t N, int L, int nC, Level l0, int batch, int64_t ystride, double* __restrict__ aux, int* __restrict__ tickets) {
  typedef FiltElem<D, F> E;
  typedef Scalar<F> SC;
  constexpr int NC = SC::NC;
  const int64_t gidx = (int64_t)blockIdx.x * TPB + threadIdx.x, ntot = (int64_t)batch * nC;
  pdl_trigger();                                   // the scan's CTAs may become resident underneath this pass
  if (gidx >= ntot) return;
  if (gidx < batch) tickets[gidx] = 0;            // per-sequence tickets of the fused final reduction (kf_chunk_lml_kernel)
  const int b = (int)(gidx / nC), c = (int)(gidx % nC);
  const int pb = sp.nparam == 1 ? 0 : b;
  const F il = 1.0 / Seed<F>::make(sp.l[pb], sp.dir_l), s = Seed<F>::make(sp.s[pb], sp.dir_s), noise = Seed<F>::make(sp.noise[pb], sp.dir_n);
  F P0[NSYM<D>]; lgssm_pinf_jordan<D, F>(s, P0);        // the whole pass runs in Jordan coordinates (lgssm_math.cuh)
  E e;
  F* Phi = e.v; F* bv = e.v + E::OB; F* dC = e.v + E::OC; F* eta = e.v + E::OE; F* J = e.v + E::OJ;
  e.set_identity();
  if (c == 0) {           // the prior: x_0 ~ N(0, P_inf) whatever came before
#pragma unroll
    for (int i = 0; i < D * D; i++) Phi[i] = 0.0;
  } else {                // start state known exactly: C = 0
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) dC[i] = -P0[i];
  }
  const int64_t k0 = (int64_t)c * L, k1 = (k0 + L < N) ? k0 + L : N;
  const double* yb = y + (int64_t)b * ystride;
  F sum_q = 0.0, prodS = 1.0;
  int eacc = 0;            // sum log S = log(prodS) + eacc log 2: the exponent of the running product is peeled off every step
  // Every lane streams its own chunk, so its (t, y, R) loads are lane-strided and their latency is a full trip to L2 /
  // HBM every fourth step.  A register pipeline needs a move per stage and iteration — which waits on the load it
  // moves (ncu: 19 % of the kernel's stall samples sat on that one instruction) — so the inputs travel through a
  // per-thread ring in shared memory instead: 8-byte cp.async copies issued KF1_PD steps ahead, read back with LDS.
  // Group j = { t[k0+j+1], y[k0+j], R[k0+j] } (the time of the NEXT step: its exponential is formed one step early).
  constexpr int RING = PD + 1;
  __shared__ double ring[(REG ? 2 : 3) * RING * TPB];
  double* sy = ring + threadIdx.x; double* sr = sy + RING * TPB; double* st = sr + RING * TPB;
  // copies are predicated on staying inside the sequence (only its last chunk can run past the end; a slot that is not
  // refilled keeps stale data that no step reads) and addressed by a 32-bit step offset from the chunk's base pointers
  const double* yc = yb + k0; const double* rc = rvec ? rvec + k0 : nullptr; const double* tc = t + k0 + 1;
  const int navail = (int)(N - k0 < (int64_t)L + PD + 1 ? N - k0 : (int64_t)L + PD + 1);     // steps of (y, R) readable from k0 on
  auto issue = [&](int j, int slot) {
    cp_async8_if(sy + slot, yc + j, j < navail);
    cp_async8_if(sr + slot, rc + j, rc != nullptr && j < navail);
    if constexpr (!REG) cp_async8_if(st + slot, tc + j, j + 1 < navail);
    cp_async_commit();
  };
#pragma unroll
  for (int j = 0; j < PD; j++) issue(j, j * TPB);
  // irregular grid: (a, e = exp(-lam a)) of the step about to run; the pair of the step after it is formed inside the
  // iteration, where its long dependent chain (range reduction + polynomial) overlaps the state recursion
  const double lam = lgssm_lambda<D>();
  const F a_reg = sp.reg_dt * il, e_reg = REG ? exp_nonpos(-lam * a_reg) : F(1.0);
  double t_cur = REG ? 0.0 : __ldg(t + k0);
  F a_cur = REG ? (k0 == 0 ? il : a_reg) : (t_cur - ((k0 == 0) ? t_cur - 1.0 : __ldg(t + k0 - 1))) * il;     // step 0 follows the t[0] - 1 prefix
  if (value_of(a_cur) > 1e4) a_cur = F(1e4);
  F e_cur = exp_nonpos(-lam * a_cur);
  const int nsteps = (int)(k1 - k0);
  int slot = 0, wslot = PD * TPB;           // ring positions (in doubles) of the group read now / issued now
  // Phi is carried as g * PhiHat: the scalar e^{-lam a} of every transition goes into g (one multiplication) instead of
  // into the nine entries, and is folded back every 16 steps (PhiHat grows at most by (1 + a + a^2/2) per step; a is
  // clamped at 1e4, far beyond the point where e^{-lam a} has underflowed, so 16 steps cannot overflow)
  F g = 1.0;
  for (int j0 = 0; j0 < nsteps; j0 += 16) {
  const int j1 = j0 + 16 < nsteps ? j0 + 16 : nsteps;
#pragma unroll 1
  for (int j = j0; j < j1; j++) {
    F T[D * D], u[D], Cn[NSYM<D>], col[D], Kg[D], hr[D];
    cp_async_wait<PD - 1>();
    const double y_cur = sy[slot];
    const F r_cur = rvec ? F(sr[slot]) : noise;
    const F a = a_cur, ee = e_cur, h = 0.5 * a * a;
    if constexpr (!REG) {
      const double t_next = st[slot];                  // (beyond the sequence: the clamped copy gives a = 0, never used)
      a_cur = (t_next - t_cur) * il; t_cur = t_next;
      if (value_of(a_cur) > 1e4) a_cur = F(1e4);
      e_cur = exp_nonpos(-lam * a_cur);
    } else { a_cur = a_reg; e_cur = e_reg; }
    issue(j + PD, wslot);
    slot = slot + TPB == RING * TPB ? 0 : slot + TPB; wslot = wslot + TPB == RING * TPB ? 0 : wslot + TPB;
    jordan_rows<D, F>(a, h, Phi, T);                      // the scalar e is applied where the rows are used
    jordan_vec<D, F>(a, h, bv, u);
    g = g * ee;
#pragma unroll
    for (int i = 0; i < D; i++) { u[i] = u[i] * ee; hr[i] = g * T[i]; }         // hr = H A Phi
    jordan_congruence<D, F>(a, h, ee * ee, dC, Cn);
#pragma unroll
    for (int i = 0; i < D; i++) col[i] = SYM(Cn, i, 0) + SYM(P0, i, 0);
    const F S = col[0] + r_cur;
    const F iS = rcp_pos(S);
    const F r = y_cur - u[0], w = iS * r;
#pragma unroll
    for (int i = 0; i < D; i++) Kg[i] = col[i] * iS;
#pragma unroll
    for (int i = 0; i < D; i++) {
      eta[i] = fma(w, hr[i], eta[i]);
      const F hi = iS * hr[i];
#pragma unroll
      for (int j = i; j < D; j++) SYM(J, i, j) = fma(hi, hr[j], SYM(J, i, j));
    }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int jj = 0; jj < D; jj++) Phi[i * D + jj] = fma(-Kg[i], T[jj], T[i * D + jj]);       // PhiHat <- (I - K H) U PhiHat
#pragma unroll
    for (int i = 0; i < D; i++) bv[i] = fma(Kg[i], r, u[i]);
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) SYM(dC, i, j) = fma(-Kg[i], col[j], SYM(Cn, i, j));
    sum_q = fma(w, r, sum_q);
    prodS = prodS * S;         // S > 0 and normal: mantissa stays in [1, 2), the exponent goes to the integer accumulator
    peel_exponent(prodS, eacc);
  }
#pragma unroll
  for (int i = 0; i < D * D; i++) Phi[i] = Phi[i] * g;
  g = 1.0;
  }
  cp_async_wait<0>();
  const F sum_logS = log(prodS) + (double)eacc * 0.693147180559945309417232121458;
  // aux rows: field-major, each field with its NC components
#pragma unroll
  for (int cc = 0; cc < NC; cc++) {
    aux[(0 * NC + cc) * ntot + gidx] = SC::comp(sum_logS, cc); aux[(1 * NC + cc) * ntot + gidx] = SC::comp(sum_q, cc);
#pragma unroll
    for (int i = 0; i < D; i++) aux[((2 + i) * NC + cc) * ntot + gidx] = SC::comp(eta[i], cc);
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) aux[((2 + D + i) * NC + cc) * ntot + gidx] = SC::comp(J[i], cc);
  }
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) dC[i] = dC[i] + P0[i];
  store_elem(e, l0.base, (int64_t)batch * l0.P, (int64_t)b * l0.P + c);
}

// Every chunk's share of (sum log S, sum alpha^2) from its aux row and the scanned state it starts from;
// block (x, b) sums 128 chunks of sequence b in fixed order -> part2[b][x][2].
template <int D, class F>
__global__ void __launch_bounds__(128)
kf_chunk_lml_kernel(Level l0, Level l1, int span, int nC, int batch, const double* __restrict__ aux, double* __restrict__ part2,
                    int* __restrict__ tickets, int64_t N, double* __restrict__ lml, double* __restrict__ dlml, double* __restrict__ sums) {
  typedef Scalar<F> SC;
  constexpr int NC = SC::NC;
  __shared__ double sh[32];
  __shared__ int last;
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  pdl_wait();
  F r0 = 0.0, r1 = 0.0;
  if (c < nC) {
    const int64_t ntot = (int64_t)batch * nC, idx = (int64_t)b * nC + c;
    auto ld = [&](int f) { F x;
#pragma unroll
      for (int cc = 0; cc < NC; cc++) SC::comp(x, cc) = aux[((int64_t)f * NC + cc) * ntot + idx];
      return x; };
    r0 = ld(0); r1 = ld(1);
    if (c > 0) {
      typedef FiltElem<D, F> E;
      const E pre = inclusive_prefix_span<E>(l0, l1, span, batch, b, c - 1);
      const F* m = pre.v + E::OB; const F* P = pre.v + E::OC;
      F eta[D], J[NSYM<D>];
#pragma unroll
      for (int i = 0; i < D; i++) eta[i] = ld(2 + i);
#pragma unroll
      for (int i = 0; i < NSYM<D>; i++) J[i] = ld(2 + D + i);
      // P = Lc Lc' (lower; a non-positive pivot — P singular to working precision — zeroes its column)
      F Lc[D * D];
#pragma unroll
      for (int j = 0; j < D; j++) {
        F inv = 0.0;
#pragma unroll
        for (int i = j; i < D; i++) {
          F v = SYM(P, i, j);
#pragma unroll
          for (int q = 0; q < j; q++) v = v - Lc[i * D + q] * Lc[j * D + q];
          if (i == j) {
            const bool pos = value_of(v) > 0.0;
            const F d = pos ? sqrt(v) : F(0.0); Lc[j * D + j] = d; inv = pos ? 1.0 / d : F(0.0);
          } else Lc[i * D + j] = v * inv;
        }
      }
      F Jm[D], a[D], W[D * D], Bm[NSYM<D>];
      symvec<D>(J, m, Jm);
      F em = 0.0, mJm = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) { em = fma(eta[i], m[i], em); mJm = fma(m[i], Jm[i], mJm); }
#pragma unroll
      for (int j = 0; j < D; j++) { F v = 0.0;
#pragma unroll
        for (int i = j; i < D; i++) v = fma(Lc[i * D + j], eta[i] - Jm[i], v);
        a[j] = v; }
#pragma unroll
      for (int i = 0; i < D; i++)           // W = J Lc
#pragma unroll
        for (int j = 0; j < D; j++) { F v = 0.0;
#pragma unroll
          for (int q = j; q < D; q++) v = fma(SYM(J, i, q), Lc[q * D + j], v);
          W[i * D + j] = v; }
#pragma unroll
      for (int i = 0; i < D; i++)           // B = I + Lc' W (symmetric)
#pragma unroll
        for (int j = i; j < D; j++) { F v = (i == j) ? 1.0 : 0.0;
#pragma unroll
          for (int q = i; q < D; q++) v = fma(Lc[q * D + i], W[q * D + j], v);
          SYM(Bm, i, j) = v; }
      // B = Lb Lb': log det B and |Lb^-1 a|^2
      F Lb[D * D], z[D], det = 1.0, zz = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) {
        F inv = 0.0;
#pragma unroll
        for (int i = j; i < D; i++) {
          F v = SYM(Bm, i, j);
#pragma unroll
          for (int q = 0; q < j; q++) v = v - Lb[i * D + q] * Lb[j * D + q];
          if (i == j) { det = det * v; const F d = sqrt(v); Lb[j * D + j] = d; inv = 1.0 / d; }
          else Lb[i * D + j] = v * inv;
        }
        F v = a[j];
#pragma unroll
        for (int q = 0; q < j; q++) v = v - Lb[j * D + q] * z[q];
        z[j] = v * inv; zz = fma(z[j], z[j], zz);
      }
      r0 = r0 + log(det);
      r1 = r1 + (mJm - 2.0 * em - zz);
    }
  }
  // the last block of a sequence to arrive sums the sequence's block partials in fixed order (no separate reduce launch)
  constexpr int NP = 2 * NC;
  double rr[NP];
#pragma unroll
  for (int cc = 0; cc < NC; cc++) { rr[cc] = block_sum(SC::comp(r0, cc), sh); rr[NC + cc] = block_sum(SC::comp(r1, cc), sh); }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NP; i++) part2[((int64_t)b * gridDim.x + blockIdx.x) * NP + i] = rr[i];
    __threadfence();
    last = atomicAdd(tickets + b, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x >= 32) return;
  __threadfence();
  double acc[NP];
#pragma unroll
  for (int i = 0; i < NP; i++) acc[i] = 0.0;
  for (int x = threadIdx.x; x < (int)gridDim.x; x += 32) {
#pragma unroll
    for (int i = 0; i < NP; i++) acc[i] += __ldcg(part2 + ((int64_t)b * gridDim.x + x) * NP + i);
  }
#pragma unroll
  for (int i = 0; i < NP; i++)
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
  if (threadIdx.x == 0) {
    if (lml) lml[b] = -0.5 * ((double)N * 1.8378770664093454835606594728112 + acc[0] + acc[NC]);
    if (dlml) {
#pragma unroll
      for (int j = 1; j < NC; j++) dlml[b * (NC - 1) + (j - 1)] = -0.5 * (acc[j] + acc[NC + j]);
    }
    if (sums) {
#pragma unroll
      for (int i = 0; i < NP; i++) sums[b * NP + i] = acc[i];
    }
    tickets[b] = 0;
  }
}

// Outputs of one filter / smoother run (device pointers, all nullable except lml or sums).
struct LgssmOut {
  double* alpha = nullptr; double* lml = nullptr; double* mean = nullptr; double* var = nullptr;
  double* table = nullptr; double* sums = nullptr;
  double* dlml = nullptr; double* dalpha = nullptr; double* dtable = nullptr;    // tangent outputs (Dual runs)
  double* fstate = nullptr;      // per sequence (m, P) after the last step
  int64_t ystride = 0;           // distance between sequences in y / alpha (0: N)
  double* table_fwd = nullptr;   // smoother runs: the forward (filter) step table next to the backward one in `table`
  bool ybroadcast = false;       // every "sequence" of the batch reads the SAME y (hyper-parameter candidates on one sequence)
};

// Inclusive prefixes of a sequence's chunk elements by ONE thread walking the chunks in order.  A prefix that starts at the
// sequence's first element is just a filtering state (A = 0, eta = 0, J = 0: the first chunk starts from the prior), so
// applying the next element is  M = (I + P J)^-1,  m' = A M (m + P eta) + b,  P' = A M P A' + C  — half a general combine
// and no shuffles.  With many sequences and few chunks each (tangent runs: 1024 x 18) this replaces the warp-shuffle
// pyramid, whose 81-double Dual elements spill (same 0.2 ms at 1024 x 10k with 18 chunks per sequence, a third of the code;
// a warp per block and a prefetched next element were measured: slower).
template <int D, class F>
__global__ void __launch_bounds__(64)
kf_prefix_seq_kernel(Level l0, int batch) {
  typedef FiltElem<D, F> E;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int64_t fstride = (int64_t)batch * l0.P, off0 = (int64_t)b * l0.P;
  E pre; load_elem(pre, l0.base, fstride, off0);
  for (int c = 1; c < l0.n; c++) {
    E e; load_elem(e, l0.base, fstride, off0 + c);
    const F* A = e.v; const F* bb = e.v + E::OB; const F* C = e.v + E::OC; const F* eta = e.v + E::OE; const F* J = e.v + E::OJ;
    F* m = pre.v + E::OB; F* P = pre.v + E::OC;
    F Mx[D * D], Mi[D * D], X[D * D], T[D * D], v[D], u[D];
    symsym<D>(P, J, Mx);
#pragma unroll
    for (int i = 0; i < D; i++) Mx[i * D + i] += 1.0;
    inv_general<D>(Mx, Mi);
    matmul<D>(A, Mi, X);
    symvec<D>(P, eta, v);
#pragma unroll
    for (int i = 0; i < D; i++) v[i] = v[i] + m[i];
    matvec<D>(X, v, u);
    matsym<D>(X, P, T);
    F Pn[NSYM<D>];
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) { F a = SYM(C, i, j);
#pragma unroll
        for (int k = 0; k < D; k++) a = fma(T[i * D + k], A[j * D + k], a);
        SYM(Pn, i, j) = a; }
#pragma unroll
    for (int i = 0; i < D; i++) m[i] = u[i] + bb[i];
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) P[i] = Pn[i];
    // the prefix as an element: (A, eta, J) stay zero
#pragma unroll
    for (int i = 0; i < D * D; i++) pre.v[i] = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) pre.v[E::OE + i] = 0.0;
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) pre.v[E::OJ + i] = 0.0;
    store_elem(pre, l0.base, fstride, off0 + c);
  }
}

// chunk length of the one-pass log-pdf: the element pass should fill the device in whole waves of resident
// threads (every thread walks the same number of steps, so a partly filled last wave costs a full one)
int onepass_chunk_length(int64_t N, int batch, int64_t resident_threads, int64_t lmax) {
  const int64_t total = N * (int64_t)batch;
  if (total <= (1 << 19)) return 8;
  if (total <= (1 << 21)) return 16;
  int64_t L = 32;
  for (int w = 1; w <= 4096; w++) {
    const int64_t nCmax = w * resident_threads / batch;
    if (nCmax < 1) continue;
    L = (N + nCmax - 1) / nCmax;
    if (L <= lmax) break;
  }
  return (int)std::max<int64_t>(L, 32);
}

template <int D, class F>
int lgssm_logpdf_onepass(gpar_ctx* ctx, SeqParams sp, int batch, int64_t N, const double* t, const double* y, const double* rvec,
                         const LgssmOut& o) {
  typedef FiltElem<D, F> FE;
  constexpr int NC = Scalar<F>::NC;
  const int64_t ystride = o.ybroadcast ? 0 : (o.ystride > 0 ? o.ystride : N);
  int variant = NC == 1 ? 1 : 2;          // 3 resident blocks of 128 threads (no register cap) measured slightly ahead of 4 capped ones
  if (const char* e = getenv("GPAR_KF1_VARIANT")) variant = atoi(e);       // tuning knob: threads x resident blocks of the element pass
  const int tpb = variant == 3 ? 64 : 128, minb = variant == 0 ? 4 : (variant == 2 ? 1 : (variant == 3 ? 4 : (variant == 4 ? 2 : 3)));
  // tangent runs: longer chunks — the scan of their 81-double elements (255 registers, spills) costs more than the element pass
  int L = onepass_chunk_length(N, batch, (int64_t)ctx->num_sms * tpb * minb, NC == 1 ? 192 : 640);
  if (const char* e = getenv("GPAR_KF_L")) { int v = atoi(e); if (v >= 4 && v <= 4096) L = v; }
  const int nC = (int)((N + L - 1) / L);
  // scan plan (values): one block per sequence when it has at most 2048 chunks, else blocks of 256 chunks (2 per lane) +
  // one block over their totals.  Tangent runs keep the 32-ary pyramid: the block-level kernel does not survive register
  // allocation with 81-double elements.
  LevelPlan fp;
  if constexpr (NC > 1) fp = plan_levels(nC, FE::NFD, batch);
  const bool two_level = NC == 1 && nC > 2048;
  const int span = NC > 1 ? 32 : (two_level ? 256 : nC), n1 = two_level ? (nC + span - 1) / span : 0;
  Level f0{nullptr, nC, (nC + 31) / 32 * 32}, f1{nullptr, n1, (n1 + 31) / 32 * 32};
  const size_t lev_doubles = NC > 1 ? fp.doubles : (size_t)FE::NFD * batch * ((size_t)f0.P + f1.P);
  const int64_t ntot = (int64_t)batch * nC;
  const int nblk = (nC + 127) / 128;
  constexpr int NAUX = (2 + D + NSYM<D>) * NC;
  CU(ctx->kal_a.reserve((lev_doubles + (size_t)NAUX * ntot + (size_t)2 * NC * batch * nblk + (size_t)(batch + 1) / 2) * sizeof(double)));
  double* base = ctx->kal_a.as<double>();
  if constexpr (NC > 1) {
    bind_levels(fp, base, FE::NFD, batch);
    f0 = fp.lv[0]; f1 = fp.lv.size() > 1 ? fp.lv[1] : Level{nullptr, 0, 0};
  } else { f0.base = base; if (two_level) f1.base = base + (size_t)FE::NFD * batch * f0.P; }
  double* aux = base + lev_doubles;
  double* part2 = aux + (size_t)NAUX * ntot;
  int* tickets = reinterpret_cast<int*>(part2 + (size_t)2 * NC * batch * nblk);
  const Level none{nullptr, 0, 0};
  const unsigned g1 = (unsigned)((ntot + tpb - 1) / tpb);
#define KF1_LAUNCH(TPB_, MINB_, PD_)                                                                                              \
  do {                                                                                                                      \
    if (sp.reg_dt > 0.0) LAUNCH(ctx, (kf_chunk_element_kernel<D, F, TPB_, MINB_, true, PD_>), g1, TPB_, 0, t, y, rvec, sp, N, L, nC, f0, batch, ystride, aux, tickets); \
    else LAUNCH(ctx, (kf_chunk_element_kernel<D, F, TPB_, MINB_, false, PD_>), g1, TPB_, 0, t, y, rvec, sp, N, L, nC, f0, batch, ystride, aux, tickets);               \
  } while (0)
  switch (variant) {
    case 0: KF1_LAUNCH(128, 4, KF1_PD); break;
    case 2: KF1_LAUNCH(128, 1, KF1_PD); break;
    case 3: KF1_LAUNCH(64, 4, KF1_PD); break;
    case 4: KF1_LAUNCH(128, 2, KF1_PD); break;
    default: KF1_LAUNCH(128, 3, KF1_PD); break;
  }
#undef KF1_LAUNCH
  bool seq_prefix = false;
  if constexpr (NC > 1) {
    seq_prefix = nC <= 64 && batch >= 128;        // many sequences, few chunks each: one thread per sequence (kf_prefix_seq_kernel)
    if (const char* e = getenv("GPAR_KF_SEQ_PREFIX")) seq_prefix = atoi(e) != 0 && nC <= 4096;      // testing knob
    if (seq_prefix) { if (nC > 1) LAUNCH(ctx, (kf_prefix_seq_kernel<D, F>), (batch + 63) / 64, 64, 0, f0, batch); }
    else if (nC > 1) CHK(run_scan<FE>(ctx, fp, batch));
  } else if (two_level) {
    LAUNCH_PDL(ctx, (scan_span_kernel<FE, 4>), dim3(n1, batch), dim3(128), 0, f0, f1, batch, span, 2);
    LAUNCH_PDL(ctx, (scan_span_kernel<FE, 8>), dim3(1, batch), dim3(256), 0, f1, none, batch, n1, std::min(8, (n1 + 255) / 256));
  } else if (nC > 1) {        // one block per sequence; few sequences: more warps, many sequences: more elements per lane
    const int want_lanes = batch >= 256 ? (nC + 2) / 3 : (nC + 1) / 2;
    if (want_lanes > 128) LAUNCH_PDL(ctx, (scan_span_kernel<FE, 8>), dim3(1, batch), dim3(256), 0, f0, none, batch, span, std::min(8, (nC + 255) / 256));
    else if (want_lanes > 64) LAUNCH_PDL(ctx, (scan_span_kernel<FE, 4>), dim3(1, batch), dim3(128), 0, f0, none, batch, span, (nC + 127) / 128);
    else if (want_lanes > 64 / 2) LAUNCH_PDL(ctx, (scan_span_kernel<FE, 2>), dim3(1, batch), dim3(64), 0, f0, none, batch, span, (nC + 63) / 64);
    else LAUNCH_PDL(ctx, (scan_span_kernel<FE, 1>), dim3(1, batch), dim3(32), 0, f0, none, batch, span, (nC + 31) / 32);
  }
  // (after the sequential prefix pass every level-0 entry is a complete prefix: no upper level)
  LAUNCH_PDL(ctx, (kf_chunk_lml_kernel<D, F>), dim3(nblk, batch), dim3(128), 0, f0, seq_prefix ? none : f1, seq_prefix ? (nC > 0 ? nC : 1) : span, nC, batch, aux, part2, tickets, N, o.lml, o.dlml, o.sums);
  return GPAR_OK;
}

template <int D, class F>
int lgssm_run_d(gpar_ctx* ctx, SeqParams sp, int batch, int64_t N, const double* t, const double* y, const double* rvec, const LgssmOut& o) {
  constexpr int NC = Scalar<F>::NC;
  typedef FiltElem<D, F> FE;
  const bool smooth = o.mean != nullptr;
  if (o.ybroadcast && o.mean) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm: candidates on one sequence: no smoother");
  {      // nothing per step asked for: the log-pdf (and its tangents) needs one pass over the data
    bool onepass = !smooth && !o.alpha && !o.table && !o.fstate && !o.dalpha && !o.dtable;
    if (const char* e = getenv("GPAR_KF_ONEPASS")) onepass = onepass && atoi(e) != 0;       // testing knob: 0 = three-phase path
    if (onepass) return lgssm_logpdf_onepass<D, F>(ctx, sp, batch, N, t, y, rvec, o);
  }
  const int64_t ystride = o.ybroadcast ? 0 : (o.ystride > 0 ? o.ystride : N);
  const int64_t astride = o.ybroadcast ? N : ystride, tstride = batch > 1 ? N : 0;
  // chunk length: long chunks amortise the scan (P2), short chunks keep small problems parallel
  const int64_t total_steps = N * (int64_t)batch;
  int L = total_steps <= (1 << 19) ? 8 : (total_steps <= (1 << 21) ? 16 : 32);
  if (N / 32 > 100000) L = 128; else if (N / 32 > 30000) L = 64;      // very long sequences: fewer scan levels
  if (const char* e = getenv("GPAR_KF_L")) { int v = atoi(e); if (v >= 4 && v <= 256) L = v; }   // tuning knob
  const int nC = (int)((N + L - 1) / L);
  LevelPlan fp = plan_levels(nC, FE::NFD, batch);
  LevelPlan spn = smooth ? plan_levels(nC, SmoothElem<D>::NFD, batch) : LevelPlan{};
  const int nslice = std::max(1, std::min(64, (nC + 2047) / 2048));
  const size_t part_doubles = ((size_t)batch * nC + (size_t)batch * nslice) * 2 * NC;
  const size_t fs_doubles = smooth ? (size_t)(D + NSYM<D>) * batch * ((size_t)((nC + 31) / 32) * 32) * L : 0;
  CU(ctx->kal_a.reserve((fp.doubles + spn.doubles + part_doubles) * sizeof(double)));
  if (smooth) CU(ctx->kal_b.reserve(fs_doubles * sizeof(double)));
  double* base = ctx->kal_a.as<double>();
  bind_levels(fp, base, FE::NFD, batch);
  if (smooth) bind_levels(spn, base + fp.doubles, SmoothElem<D>::NFD, batch);
  double* part = base + fp.doubles + spn.doubles;
  double* fs = smooth ? ctx->kal_b.as<double>() : nullptr;
  const Level none{nullptr, 0, 0};
  const Level f0 = fp.lv[0], f1 = fp.lv.size() > 1 ? fp.lv[1] : none;
  const unsigned g1 = (unsigned)(((int64_t)batch * f0.P + 127) / 128);
  LAUNCH(ctx, (kf_chunk_summary_kernel<D, F>), g1, 128, 0, t, y, rvec, sp, N, L, nC, f0, batch, ystride);
  CHK(run_scan<FE>(ctx, fp, batch));
  const unsigned g3 = (unsigned)(((int64_t)batch * nC + 127) / 128);
  dim3 g3b((nC + 127) / 128, batch);
  if constexpr (NC == 1) {
    if (smooth) {
      const Level s0 = spn.lv[0], s1 = spn.lv.size() > 1 ? spn.lv[1] : none;
      LAUNCH(ctx, (kf_chunk_filter_kernel<D, true, double>), g3, 128, 0, t, y, rvec, sp, N, L, nC, f0, f1, batch, o.alpha, part, fs, s0,
             o.table_fwd, (double*)nullptr, (double*)nullptr, ystride, (double*)nullptr, astride, tstride);
      if (s0.P > nC) { dim3 gp((s0.P - nC + 127) / 128, batch); LAUNCH(ctx, smooth_pad_kernel<D>, gp, 128, 0, s0, nC, batch); }
      CHK(run_scan<SmoothElem<D>>(ctx, spn, batch));
      LAUNCH(ctx, ks_backward_kernel<D>, g3b, 128, 0, t, sp, N, L, nC, s0, s1, batch, fs, o.mean, o.var, o.table);     // (smoother: o.table = backward table)
    } else {
      LAUNCH(ctx, (kf_chunk_filter_kernel<D, false, double>), g3, 128, 0, t, y, rvec, sp, N, L, nC, f0, f1, batch, o.alpha, part, fs, none,
             o.table, (double*)nullptr, (double*)nullptr, ystride, o.fstate, astride, tstride);
    }
  } else {
    if (smooth) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm: the smoother has no tangent mode");
    LAUNCH(ctx, (kf_chunk_filter_kernel<D, false, F>), g3, 128, 0, t, y, rvec, sp, N, L, nC, f0, f1, batch, o.alpha, part, fs, none,
           o.table, o.dalpha, o.dtable, ystride, (double*)nullptr, astride, tstride);
  }
  double* part2 = part + (size_t)batch * nC * 2 * NC;
  LAUNCH(ctx, lml_partial_kernel<NC>, dim3(nslice, batch), 256, 0, part, nC, nslice, part2);
  LAUNCH(ctx, lml_reduce_kernel<NC>, batch, 64, 0, part2, nslice, N, o.lml, o.dlml, o.sums);
  return GPAR_OK;
}

// ---- steady-state path for regular grids with scalar noise -----------------------------------------
// On a regular grid the model is time invariant, the covariance recursion converges geometrically to
// the fixed point of the Riccati map, and from then on every step uses the same (Phi, K, S): the filter
// reduces to the affine mean recursion  m_k = Phi m_{k-1} + K y_k,  alpha_k = (y_k - HA m_{k-1}) / sqrt(S)
// (the textbook steady-state Kalman filter).  The general scan handles the first KTR steps, the
// hand-over kernel verifies |P' - P| <= 1e-12 |P| for every sequence (otherwise the caller falls back
// to the general path), and the remaining steps run in two HBM-bound passes (chunk responses ->
// affine scan over chunks -> emit), with y / alpha moved through shared memory so that global accesses
// are coalesced although each thread walks its own chunk.
template <int D> struct SSLayout {      // per-sequence constants
  static constexpr int PHI = 0, K = D * D, HA = K + D, RS = HA + D, LOGS = RS + 1, M0 = LOGS + 1, PHIL = M0 + D, OK = PHIL + D * D, SIZE = OK + 1;
};

template <int D>
__global__ void __launch_bounds__(128)
ss_setup_kernel(SeqParams sp, int batch, const double* __restrict__ fstate, int L, double* __restrict__ ssc) {
  typedef SSLayout<D> SL;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const int pb = sp.nparam == 1 ? 0 : b;
  const double il = 1.0 / sp.l[pb], s = sp.s[pb], noise = sp.noise[pb];
  double P0[NSYM<D>]; lgssm_pinf<D>(P0);
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) P0[i] *= s;
  double A[D * D]; lgssm_transition<D>(sp.reg_dt * il, A);
  double P[NSYM<D>], Pp[NSYM<D>], Kg[D], S = 1.0;
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) P[i] = fstate[(int64_t)b * (D + NSYM<D>) + D + i];
  bool ok = true;
  for (int it = 0; it < 2; it++) {     // two more Riccati steps: both must leave P unchanged to 1e-12
    predict_cov<D>(A, P, P0, Pp);
    S = Pp[0] + noise;
    const double iS = 1.0 / S;
    double Pn[NSYM<D>], dmax = 0.0, pmax = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) Kg[i] = SYM(Pp, i, 0) * iS;
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) SYM(Pn, i, j) = fma(-S * Kg[i], Kg[j], SYM(Pp, i, j));
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) { dmax = fmax(dmax, fabs(Pn[i] - P[i])); pmax = fmax(pmax, fabs(P[i])); P[i] = Pn[i]; }
    ok = ok && (dmax <= 1e-12 * pmax);
  }
  double* o = ssc + (int64_t)b * SL::SIZE;
  double Phi[D * D];
#pragma unroll
  for (int i = 0; i < D; i++) {
#pragma unroll
    for (int j = 0; j < D; j++) Phi[i * D + j] = fma(-Kg[i], A[j], A[i * D + j]);
    o[SL::K + i] = Kg[i]; o[SL::HA + i] = A[i]; o[SL::M0 + i] = fstate[(int64_t)b * (D + NSYM<D>) + i];
  }
  o[SL::RS] = rsqrt(S); o[SL::LOGS] = log(S); o[SL::OK] = ok ? 1.0 : 0.0;
  double Sq[D * D], Pw[D * D], T[D * D];
#pragma unroll
  for (int i = 0; i < D * D; i++) { o[SL::PHI + i] = Phi[i]; Sq[i] = Phi[i]; Pw[i] = (i / D == i % D) ? 1.0 : 0.0; }
  for (int e = L; e > 0; e >>= 1) {      // Phi^L by square-and-multiply
    if (e & 1) {
      matmul<D>(Sq, Pw, T);
#pragma unroll
      for (int i = 0; i < D * D; i++) Pw[i] = T[i];
    }
    matmul<D>(Sq, Sq, T);
#pragma unroll
    for (int i = 0; i < D * D; i++) Sq[i] = T[i];
  }
#pragma unroll
  for (int i = 0; i < D * D; i++) o[SL::PHIL + i] = Pw[i];
}

constexpr int SS_WARPS = 4;
__device__ __forceinline__ void ss_cp_async8(double* dst, const double* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// EMIT = false: chunk response r_c (zero start) -> affine element (Phi^L, r_c) at level 0 of the scan.
// EMIT = true : restart from the scanned state, emit alpha (optional) and the chunk's sum alpha^2.
// A warp owns 32 consecutive chunks; the 32 x 32 block of their next steps is fetched with cp.async
// (each row a coalesced 256-byte run) into a shared tile, consumed column-wise, and alpha leaves through
// the same tile.  (ncu: the kernel is latency bound at the 12 warps/SM a double-buffered tile allows;
// a single tile per warp doubles the resident warps, which hides more than the prefetch did.)
template <int D, bool EMIT>
__global__ void __launch_bounds__(SS_WARPS * 32, 5)
ss_chunk_kernel(const double* __restrict__ y, int64_t ystride, int64_t k_lo, int64_t N, int L, int nC, const double* __restrict__ ssc,
                Level l0, Level l1, int batch, double* __restrict__ alpha, double* __restrict__ part) {
  typedef SSLayout<D> SL;
  typedef AffineElem<D> AE;
  __shared__ double sm[SS_WARPS][32][33];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * (SS_WARPS * 32) + threadIdx.x, b = blockIdx.y;
  const int cw = c - lane;                                 // first chunk of this warp
  const double* cst = ssc + (int64_t)b * SL::SIZE;
  double Phi[D * D], Kg[D], ha[D], x[D];
#pragma unroll
  for (int i = 0; i < D * D; i++) Phi[i] = cst[SL::PHI + i];
#pragma unroll
  for (int i = 0; i < D; i++) { Kg[i] = cst[SL::K + i]; ha[i] = cst[SL::HA + i]; x[i] = 0.0; }
  const double rs = cst[SL::RS];
  const double* yb = y + (int64_t)b * ystride;
  double* ab = (EMIT && alpha) ? alpha + (int64_t)b * ystride : nullptr;
  const int64_t kw = k_lo + (int64_t)cw * L + lane;        // this lane's column in row 0 of the warp's tile
  const int nrow = min(32, nC - cw);                        // rows (chunks) of this warp's tile that exist
  auto issue = [&](int j0) {
    const double* src = yb + kw + j0;
    const int64_t room = N - (kw + j0);                    // row r is in range iff r * L < room
#pragma unroll
    for (int r = 0; r < 32; r++) {
      double* dst = &sm[w][r][lane];
      if (r < nrow && (int64_t)r * L < room) ss_cp_async8(dst, src + (int64_t)r * L); else *dst = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (EMIT && c < nC) {
    if (c == 0) {
#pragma unroll
      for (int i = 0; i < D; i++) x[i] = cst[SL::M0 + i];
    } else {      // state entering chunk c = prefix map of chunks 0..c-1 applied to the hand-over state
      AE pre = inclusive_prefix<AE>(l0, l1, batch, b, c - 1);
      double m0[D];
#pragma unroll
      for (int i = 0; i < D; i++) m0[i] = cst[SL::M0 + i];
      matvec<D>(pre.v, m0, x);
#pragma unroll
      for (int i = 0; i < D; i++) x[i] += pre.v[AE::OR + i];
    }
  }
  const int64_t kc = k_lo + (int64_t)c * L;                // first step of this thread's chunk
  double a2 = 0.0;
  for (int j0 = 0; j0 < L; j0 += 32) {
    issue(j0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (c < nC) {
      double* row = sm[w][lane];
      const int64_t left = N - (kc + j0);
      const int nv = left >= 32 ? 32 : (left > 0 ? (int)left : 0);     // valid steps of this thread in the block
#pragma unroll 8
      for (int j = 0; j < 32; j++) {
        if (j < nv) {
          const double yv = row[j];
          if (EMIT) {
            double pred = 0.0;
#pragma unroll
            for (int q = 0; q < D; q++) pred = fma(ha[q], x[q], pred);
            const double a = (yv - pred) * rs;
            a2 = fma(a, a, a2);
            row[j] = a;
          }
          double nx[D];
#pragma unroll
          for (int i = 0; i < D; i++) { double v = Kg[i] * yv;
#pragma unroll
            for (int q = 0; q < D; q++) v = fma(Phi[i * D + q], x[q], v);
            nx[i] = v; }
#pragma unroll
          for (int i = 0; i < D; i++) x[i] = nx[i];
        }
      }
    }
    __syncwarp();
    if (EMIT && ab) {
#pragma unroll 8
      for (int r = 0; r < 32; r++) {
        const int64_t k = kw + (int64_t)r * L + j0;
        if (r < nrow && k < N) ab[k] = sm[w][r][lane];
      }
      __syncwarp();
    }
  }
  if (EMIT) {
    if (c < nC) { part[((int64_t)b * nC + c) * 2] = 0.0; part[((int64_t)b * nC + c) * 2 + 1] = a2; }
  } else if (c < l0.P) {
    AE e;
    if (c < nC) {
#pragma unroll
      for (int i = 0; i < D * D; i++) e.v[i] = cst[SL::PHIL + i];
#pragma unroll
      for (int i = 0; i < D; i++) e.v[AE::OR + i] = x[i];
    } else e.set_identity();
    store_elem(e, l0.base, (int64_t)batch * l0.P, (int64_t)b * l0.P + c);
  }
}

// lml[b] = -1/2 (N log 2pi + [transient sum log S + (N - k_lo) log S_ss] + [transient + steady sum alpha^2]); flags[b] = converged
template <int D>
__global__ void ss_finish_kernel(const double* __restrict__ sums_tr, const double* __restrict__ sums_ss, const double* __restrict__ ssc,
                                 int64_t N, int64_t k_lo, int batch, double* __restrict__ lml, double* __restrict__ sums, double* __restrict__ flags) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* cst = ssc + (int64_t)b * SSLayout<D>::SIZE;
  const double slog = sums_tr[2 * b] + (double)(N - k_lo) * cst[SSLayout<D>::LOGS];
  const double sa2 = sums_tr[2 * b + 1] + sums_ss[2 * b + 1];
  if (lml) lml[b] = -0.5 * ((double)N * 1.8378770664093454835606594728112 + slog + sa2);
  if (sums) { sums[2 * b] = slog; sums[2 * b + 1] = sa2; }
  flags[b] = cst[SSLayout<D>::OK];
}

// ---- single-pass steady-state path for LONG sequences (few sequences, N >> transient) -----------------
// The two-pass scheme above reads y twice and needs ~18 launches.  For a long sequence the mean recursion
// m_k = Phi m_{k-1} + K y_k forgets its start state geometrically (rho(Phi) < 1), so a block that starts W
// steps BEFORE its segment from a zero state arrives with the exact state to working precision once
// ||Phi^W|| <= 1e-18 ("burn-in"; W is found by repeated squaring).  Every block is then independent:
// one pass over y (plus W / segment re-reads), no inter-block scan, three launches in total:
//   ss2_setup : per sequence, the steady-state covariance by DOUBLING — the filtering element of one step
//               combined with itself (FiltElem::combine) covers 2, 4, 8, ... steps and its C converges to
//               the fixed point of the Riccati map quadratically; one Riccati step verifies it (1e-12);
//               steady constants, the powers Phi^(32 2^j) of the block scan, the 4-step blocking constants, W;
//   ss2_main  : block = 256 sub-chunks of 32 steps staged in shared memory (cp.async, coalesced);
//               sub-chunk responses (4 steps per dependent mat-vec) -> block-level scan with the constant
//               powers (warp shuffles) -> emit alpha, sum alpha^2.  The first block of a sequence walks the
//               Riccati transient from the prior sequentially (one lane) until P has stopped moving (1e-13,
//               twice) and hands the exact state to its remaining sub-chunks — it runs underneath the others;
//   ss2_finish: fixed-order sums -> lml.
constexpr int SS2_LS = 32, SS2_THREADS = 256, SS2_STEPS = SS2_LS * SS2_THREADS;     // 8192 steps per block (incl. burn-in)
constexpr int SS2_WCAP = 2048, SS2_WFIX = 1024;
template <int D> struct SS2Layout {      // per-sequence constants (doubles)
  static constexpr int TR = D * D + 2 * D + 2;                                     // Phi, K, HA, rs, log S
  static constexpr int NPOW = 6;                                                    // Phi^(LS 2^j), j = 0..5 (sub-chunk ... warp)
  static constexpr int ROW = 0, POWS = TR, PHI4 = POWS + NPOW * D * D, G4 = PHI4 + D * D, KSTAR = G4 + 4 * D, W = KSTAR + 1, OK = W + 1,
                       SIZE = OK + 1;
};

// Last-CTA-done epilogue of the single pass (ss3 variants): every CTA of the main and of the head launch takes a ticket
// after publishing its partial sums; the holder of the last ticket sums all sequences in fixed order, writes lml / sums
// and folds the "cannot handle this model" flags — no separate finish launch.  ticket == nullptr: epilogue off.
struct SSFin { int* ticket; int total; int batch; double* lml; double* sums; double* flag_out; };

template <int D>
__device__ __forceinline__ void ss_ticket_finish(const SSFin fin, const double* part, int nseg, const double* cst, int64_t N, int* sh_last) {
  typedef SS2Layout<D> SL;
  if (!fin.ticket) return;
  __syncthreads();                                           // the CTA's partials were written by thread 0
  if (threadIdx.x == 0) {
    __threadfence();
    *sh_last = atomicAdd(fin.ticket, 1) == fin.total - 1;
  }
  __syncthreads();
  if (!*sh_last) return;
  __threadfence();
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int b = threadIdx.x >> 5; b < fin.batch; b += nw) {
    const double* c = cst + (int64_t)b * SL::SIZE;
    double a0 = 0.0, a1 = 0.0;
    for (int sg = lane; sg < nseg; sg += 32) { a0 += __ldcg(part + ((int64_t)b * nseg + sg) * 2); a1 += __ldcg(part + ((int64_t)b * nseg + sg) * 2 + 1); }
    for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    if (lane == 0) {
      const int64_t kstar = (int64_t)__ldcg(c + SL::KSTAR);
      const double slog = a0 + (double)(N > kstar ? N - kstar : 0) * __ldcg(c + SL::ROW + D * D + 2 * D + 1);
      if (fin.lml) fin.lml[b] = -0.5 * ((double)N * 1.8378770664093454835606594728112 + slog + a1);
      if (fin.sums) { fin.sums[2 * b] = slog; fin.sums[2 * b + 1] = a1; }
    }
  }
  int bad = 0;
  for (int q = threadIdx.x; q < fin.batch; q += blockDim.x) bad |= __ldcg(cst + (int64_t)q * SL::SIZE + SL::OK) != 1.0;
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    fin.flag_out[0] = bad ? 0.0 : 1.0;
    *fin.ticket = 0;                                         // ready for the next call
  }
}

template <int D>
__global__ void __launch_bounds__(32)
ss2_setup_kernel(SeqParams sp, int batch, double* __restrict__ cst, int* __restrict__ ticket) {
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;");        // the main pass may start loading its first tiles now (it waits before reading cst)
#endif
  typedef SS2Layout<D> SL;
  typedef FiltElem<D> FE;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0 && ticket) *ticket = 0;
  if (b >= batch) return;
  const int pb = sp.nparam == 1 ? 0 : b;
  const double il = 1.0 / sp.l[pb], s = sp.s[pb], noise = sp.noise[pb];
  double P0[NSYM<D>], Zs[NSYM<D>], Q[NSYM<D>], A[D * D];
  lgssm_pinf<D>(P0);
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) { P0[i] *= s; Zs[i] = 0.0; }
  lgssm_transition<D>(sp.reg_dt * il, A);
  predict_cov<D>(A, Zs, P0, Q);                              // Q = P0 - A P0 A'
  // filtering element of one step, data parts zero
  FE e; e.set_identity();
  {
    const double S = Q[0] + noise, iS = 1.0 / S;
    double Kg[D];
#pragma unroll
    for (int i = 0; i < D; i++) Kg[i] = SYM(Q, i, 0) * iS;
#pragma unroll
    for (int i = 0; i < D; i++) {
#pragma unroll
      for (int j = 0; j < D; j++) e.v[i * D + j] = fma(-Kg[i], A[j], A[i * D + j]);
#pragma unroll
      for (int j = i; j < D; j++) { SYM((e.v + FE::OC), i, j) = fma(-S * Kg[i], Kg[j], SYM(Q, i, j)); SYM((e.v + FE::OJ), i, j) = A[i] * A[j] * iS; }
    }
  }
  bool conv = false;
  for (int it = 0; it < 16 && !conv; it++) {                 // 2^16 steps: far beyond any model the burn-in cap admits
    FE n = FE::combine(e, e);
    double dmax = 0.0, pmax = 0.0;
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) { dmax = fmax(dmax, fabs(n.v[FE::OC + i] - e.v[FE::OC + i])); pmax = fmax(pmax, fabs(n.v[FE::OC + i])); }
    conv = dmax <= 1e-14 * pmax;
    e = n;
  }
  double P[NSYM<D>], Pp[NSYM<D>], Kg[D], row[SL::TR];
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) P[i] = e.v[FE::OC + i];
  predict_cov<D>(A, P, P0, Pp);
  const double S = Pp[0] + noise, iS = 1.0 / S;
  double dmax = 0.0, pmax = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) Kg[i] = SYM(Pp, i, 0) * iS;
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = i; j < D; j++) { const double pn = fma(-S * Kg[i], Kg[j], SYM(Pp, i, j)); dmax = fmax(dmax, fabs(pn - SYM(P, i, j))); pmax = fmax(pmax, fabs(pn)); }
  bool ok = conv && dmax <= 1e-12 * pmax;
#pragma unroll
  for (int i = 0; i < D; i++) {
#pragma unroll
    for (int j = 0; j < D; j++) row[i * D + j] = fma(-Kg[i], A[j], A[i * D + j]);
    row[D * D + i] = Kg[i]; row[D * D + D + i] = A[i];
  }
  row[D * D + 2 * D] = rsqrt(S); row[D * D + 2 * D + 1] = log(S);
  double* o = cst + (int64_t)b * SL::SIZE;
#pragma unroll
  for (int i = 0; i < SL::TR; i++) o[SL::ROW + i] = row[i];
  // 4-step blocking: x_{k+4} = Phi^4 x_k + sum_i (Phi^(3-i) K) y_{k+i}
  {
    double v[D], u[D];
#pragma unroll
    for (int i = 0; i < D; i++) { v[i] = Kg[i]; o[SL::G4 + 3 * D + i] = v[i]; }
    for (int q = 2; q >= 0; q--) {
      matvec<D>(row, v, u);
#pragma unroll
      for (int i = 0; i < D; i++) { v[i] = u[i]; o[SL::G4 + q * D + i] = v[i]; }
    }
  }
  // powers Phi^(LS 2^j) for the block-level scan, and the burn-in length W: first W >= LS with ||Phi^W||_inf <= 1e-18
  double M[D * D], T[D * D];
#pragma unroll
  for (int i = 0; i < D * D; i++) M[i] = row[i];
  int W = 0, cur = 1;
  for (int q = 0; q < 12; q++) {                             // cur = 2^q
    if (cur == 4) {
#pragma unroll
      for (int i = 0; i < D * D; i++) o[SL::PHI4 + i] = M[i];
    }
    for (int j = 0; j < SL::NPOW; j++)
      if (cur == (SS2_LS << j)) {
#pragma unroll
        for (int i = 0; i < D * D; i++) o[SL::POWS + j * D * D + i] = M[i];
      }
    double nrm = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) { double r = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) r += fabs(M[i * D + j]);
      nrm = fmax(nrm, r); }
    if (W == 0 && nrm <= 1e-18 && cur >= SS2_LS) W = cur;
    if (cur >= (SS2_LS << (SL::NPOW - 1)) && (W > 0 || cur >= SS2_WCAP)) break;
    matmul<D>(M, M, T);
#pragma unroll
    for (int i = 0; i < D * D; i++) M[i] = T[i];
    cur *= 2;
  }
  if (W == 0) { W = SS2_WCAP; ok = false; }                    // start state not forgotten within the cap: the caller falls back
  if (W > SS2_WFIX) ok = false;                                // the fixed burn-in of the main pass would be too short
  o[SL::KSTAR] = 0.0; o[SL::W] = (double)W; o[SL::OK] = ok ? 1.0 : 0.0;
}

// HEAD = true: the first block of every sequence (transient + hand-over), launched on the side stream so that
// it runs underneath the HEAD = false blocks (all the others); scratchS: SS2_STEPS doubles per sequence.
template <int D, bool HEAD>
__global__ void __launch_bounds__(SS2_THREADS, HEAD ? 1 : 3)
ss2_main_kernel(const double* __restrict__ y, int64_t ystride, int64_t N, SeqParams sp, double* __restrict__ cst,
                int W, int nseg, double* __restrict__ alpha, double* __restrict__ part, double* __restrict__ scratchS, SSFin fin) {
  typedef SS2Layout<D> SL;
  extern __shared__ double ysm[];                            // SS2_STEPS values, padded by one per sub-chunk
  __shared__ double wtot[SS2_THREADS / 32][D];
  __shared__ int fin_last;
  __shared__ double inject[D], trsum[2], red[32];
  __shared__ int kst_sh;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y, seg = HEAD ? 0 : blockIdx.x + 1;
  const double* c = cst + (int64_t)b * SL::SIZE;
  const int payload = SS2_STEPS - W;
  // block 0 of a sequence: window [0, 8192) (transient + exact hand-over); block s >= 1: W burn-in steps, then its payload
  const int64_t load_start = HEAD ? 0 : (int64_t)SS2_STEPS + (int64_t)(seg - 1) * payload - W;
  if (load_start + (HEAD ? 0 : W) >= N) return;               // nothing to emit (uniform per block)
  const double* yb = y + (int64_t)b * ystride;
  {
    static_assert(SS2_THREADS % SS2_LS == 0, "a pass of the block covers whole sub-chunks");
    const double* src = yb + load_start + tid;
    double* dst = ysm + tid + tid / SS2_LS;
    constexpr int ROWS_PER_PASS = SS2_THREADS / SS2_LS;
    if (load_start + SS2_STEPS <= N) {
#pragma unroll 8
      for (int i = 0; i < SS2_LS; i++) ss_cp_async8(dst + i * (SS2_THREADS + ROWS_PER_PASS), src + i * SS2_THREADS);
    } else {
      for (int i = 0; i < SS2_LS; i++) {
        const int64_t k = load_start + tid + (int64_t)i * SS2_THREADS;
        if (k < N) ss_cp_async8(dst + i * (SS2_THREADS + ROWS_PER_PASS), src + i * SS2_THREADS);
        else dst[i * (SS2_THREADS + ROWS_PER_PASS)] = 0.0;
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  double Phi[D * D], Kg[D], ha[D];
#pragma unroll
  for (int i = 0; i < D * D; i++) Phi[i] = c[SL::ROW + i];
#pragma unroll
  for (int i = 0; i < D; i++) { Kg[i] = c[SL::ROW + D * D + i]; ha[i] = c[SL::ROW + D * D + D + i]; }
  const double rs = c[SL::ROW + D * D + 2 * D];
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  int kst = W;                                                // first emitting step of this block's window (local index)
  if constexpr (HEAD) {
    double* Sk = scratchS + (int64_t)b * SS2_STEPS;
    if (tid == 0) {
      // the Riccati transient from the prior, sequentially: covariance + mean recursion only; the innovations
      // replace y in place and S_k goes to scratch — alpha_k and log S_k are formed afterwards by all threads
      const int pb = sp.nparam == 1 ? 0 : b;
      const double il = 1.0 / sp.l[pb], s = sp.s[pb], noise = sp.noise[pb];
      double P0[NSYM<D>], P[NSYM<D>], A[D * D], x[D];
      lgssm_pinf<D>(P0);
#pragma unroll
      for (int i = 0; i < NSYM<D>; i++) { P0[i] *= s; P[i] = P0[i]; }
#pragma unroll
      for (int i = 0; i < D; i++) x[i] = 0.0;
      lgssm_transition<D>(1.0 * il, A);                      // step 0 follows the t[0] - 1 prefix
      int conv = 0, found = -1;
      const int kmax = (int)(N < SS2_STEPS - SS2_LS ? N : SS2_STEPS - SS2_LS);
      int k = 0;
      for (; k < kmax; k++) {
        if (k == 1) lgssm_transition<D>(sp.reg_dt * il, A);
        double Pp[NSYM<D>], mp[D], Kt[D];
        predict_cov<D>(A, P, P0, Pp);
        matvec<D>(A, x, mp);
        const double S = Pp[0] + noise, iS = 1.0 / S, inn = ysm[k + k / SS2_LS] - mp[0];
        ysm[k + k / SS2_LS] = inn;
        Sk[k] = S;
        double Pn[NSYM<D>];
#pragma unroll
        for (int i = 0; i < D; i++) { Kt[i] = SYM(Pp, i, 0) * iS; x[i] = fma(Kt[i], inn, mp[i]); }
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
          for (int j = i; j < D; j++) SYM(Pn, i, j) = fma(-S * Kt[i], Kt[j], SYM(Pp, i, j));
        // "P has stopped moving" is only looked at on the last two steps of a sub-chunk (a hand-over is only
        // possible at its end), so that the max-reductions stay off the sequential critical path elsewhere
        if ((k + 2) % SS2_LS < 2 && k >= 2) {
          double dmax = 0.0, pmax = 0.0;
#pragma unroll
          for (int i = 0; i < NSYM<D>; i++) { dmax = fmax(dmax, fabs(Pn[i] - P[i])); pmax = fmax(pmax, fabs(Pn[i])); }
          conv = dmax <= 1e-13 * pmax ? conv + 1 : 0;
        } else conv = 0;
#pragma unroll
        for (int i = 0; i < NSYM<D>; i++) P[i] = Pn[i];
        if (conv >= 2 && (k + 1) % SS2_LS == 0) { found = k + 1; break; }
      }
      if (found < 0 && k >= N) found = (int)((N + SS2_LS - 1) / SS2_LS * SS2_LS);   // the sequence ended inside the transient
      if (found < 0) { found = SS2_STEPS; cst[(int64_t)b * SL::SIZE + SL::OK] = 0.0; }   // not converged: the caller falls back
      kst_sh = found;
      cst[(int64_t)b * SL::SIZE + SL::KSTAR] = (double)(found < N ? found : N);
#pragma unroll
      for (int i = 0; i < D; i++) inject[i] = x[i];
      __threadfence_block();
    }
    __syncthreads();
    kst = kst_sh;
    // alpha_k = innovation / sqrt(S_k), sum log S_k, sum alpha_k^2 over the transient, in parallel
    double sl = 0.0, sa = 0.0;
    const int ntr = (int)(kst < N ? kst : N);
    for (int k = tid; k < ntr; k += SS2_THREADS) {
      const double S = Sk[k], a = ysm[k + k / SS2_LS] * rsqrt(S);
      ysm[k + k / SS2_LS] = a; sl += log(S); sa = fma(a, a, sa);
    }
    const double r0 = block_sum(sl, red);
    const double r1 = block_sum(sa, red);
    if (tid == 0) { trsum[0] = r0; trsum[1] = r1; }
    __syncthreads();
  }
  // phase 1: response of sub-chunk tid (zero start state; the hand-over sub-chunk of block 0 starts exactly)
  const int64_t k0 = load_start + (int64_t)tid * SS2_LS;     // first step of this sub-chunk
  const bool steady = tid * SS2_LS >= kst || !HEAD;           // block 0: sub-chunks before kst were done by the transient
  const bool active = k0 < N && steady;
  const bool pay = active && tid * SS2_LS >= kst;             // emits (burn-in sub-chunks only warm up)
  const bool injected = HEAD && tid * SS2_LS == kst;
  const int nv = !active ? 0 : (N - k0 >= SS2_LS ? SS2_LS : (int)(N - k0));
  const double* row = ysm + tid * (SS2_LS + 1);
  double rsp[D];                                             // state at the end of the sub-chunk for a zero state at its start
#pragma unroll
  for (int i = 0; i < D; i++) rsp[i] = 0.0;
  if (active && nv == SS2_LS) {                               // (a partial last sub-chunk has no successor)
    double P4[D * D], G[4][D];
#pragma unroll
    for (int i = 0; i < D * D; i++) P4[i] = c[SL::PHI4 + i];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int i = 0; i < D; i++) G[q][i] = c[SL::G4 + q * D + i];
    if (injected) {
#pragma unroll
      for (int i = 0; i < D; i++) rsp[i] = inject[i];        // (everything before it in the block is identically zero)
    }
#pragma unroll
    for (int j = 0; j < SS2_LS; j += 4) {
      double nx[D];
#pragma unroll
      for (int i = 0; i < D; i++) {
        double v = G[0][i] * row[j];
        v = fma(G[1][i], row[j + 1], v); v = fma(G[2][i], row[j + 2], v); v = fma(G[3][i], row[j + 3], v);
#pragma unroll
        for (int q = 0; q < D; q++) v = fma(P4[i * D + q], rsp[q], v);
        nx[i] = v;
      }
#pragma unroll
      for (int i = 0; i < D; i++) rsp[i] = nx[i];
    }
  }
  // phase 2: block-level scan of the sub-chunk responses.  All sub-chunks span SS2_LS steps, so the combine
  // at distance 2^j only needs the constant power Phi^(LS 2^j):  x_i <- Phi^(LS 2^j) x_(i - 2^j) + x_i.
  const double* pw = c + SL::POWS;
  auto warp_scan = [&](double* v) {
#pragma unroll
    for (int j = 0; j < 5; j++) {
      double o[D];
#pragma unroll
      for (int i = 0; i < D; i++) o[i] = __shfl_up_sync(0xffffffffu, v[i], 1 << j);
      if (lane >= (1 << j)) {
#pragma unroll
        for (int i = 0; i < D; i++) { double a = v[i];
#pragma unroll
          for (int q = 0; q < D; q++) a = fma(__ldg(pw + j * D * D + i * D + q), o[q], a);
          v[i] = a; }
      }
    }
  };
  double tot[D];
#pragma unroll
  for (int i = 0; i < D; i++) tot[i] = rsp[i];
  warp_scan(tot);                                             // lane 31: state at the end of the warp for a zero state at its start
  if (lane == 31) {
#pragma unroll
    for (int i = 0; i < D; i++) wtot[wid][i] = tot[i];
  }
  __syncthreads();
  double carry[D];                                            // state at the start of this warp (block start state: zero)
#pragma unroll
  for (int i = 0; i < D; i++) carry[i] = 0.0;
  for (int w2 = 0; w2 < wid; w2++) {
    double nx[D];
#pragma unroll
    for (int i = 0; i < D; i++) { double a = wtot[w2][i];
#pragma unroll
      for (int q = 0; q < D; q++) a = fma(__ldg(pw + 5 * D * D + i * D + q), carry[q], a);
      nx[i] = a; }
#pragma unroll
    for (int i = 0; i < D; i++) carry[i] = nx[i];
  }
  if (lane == 0) {                                            // the warp's carry-in enters through its first sub-chunk
#pragma unroll
    for (int i = 0; i < D; i++) { double a = rsp[i];
#pragma unroll
      for (int q = 0; q < D; q++) a = fma(__ldg(pw + i * D + q), carry[q], a);
      rsp[i] = a; }
  }
  warp_scan(rsp);                                             // inclusive: state at the end of sub-chunk tid
  double xs[D];                                               // state at the start of sub-chunk tid
#pragma unroll
  for (int i = 0; i < D; i++) { xs[i] = __shfl_up_sync(0xffffffffu, rsp[i], 1); if (lane == 0) xs[i] = carry[i]; }
  // phase 3: payload sub-chunks restart from their start state and emit
  double a2 = 0.0;
  if (pay) {
    double x[D];
#pragma unroll
    for (int i = 0; i < D; i++) x[i] = injected ? inject[i] : xs[i];
    double* rw = ysm + tid * (SS2_LS + 1);
#pragma unroll 8
    for (int j = 0; j < SS2_LS; j++) {
      if (j < nv) {
        const double yv = rw[j];
        double pred = 0.0, nx[D];
#pragma unroll
        for (int q = 0; q < D; q++) pred = fma(ha[q], x[q], pred);
        const double a = (yv - pred) * rs;
        a2 = fma(a, a, a2);
        rw[j] = a;
#pragma unroll
        for (int i = 0; i < D; i++) { double v = Kg[i] * yv;
#pragma unroll
          for (int q = 0; q < D; q++) v = fma(Phi[i * D + q], x[q], v);
          nx[i] = v; }
#pragma unroll
        for (int i = 0; i < D; i++) x[i] = nx[i];
      }
    }
  }
  __syncthreads();
  if (alpha) {
    double* ab = alpha + (int64_t)b * ystride;
    for (int e2 = (HEAD ? 0 : W) + tid; e2 < SS2_STEPS; e2 += SS2_THREADS) {
      const int64_t k = load_start + e2;
      if (k < N) ab[k] = ysm[e2 + e2 / SS2_LS];
    }
  }
  const double a2tot = block_sum(a2, red);
  if (tid == 0) {
    part[((int64_t)b * nseg + seg) * 2] = HEAD ? trsum[0] : 0.0;
    part[((int64_t)b * nseg + seg) * 2 + 1] = a2tot + (HEAD ? trsum[1] : 0.0);
  }
  ss_ticket_finish<D>(fin, part, nseg, cst, N, &fin_last);
}

template <int D>
__global__ void ss2_flags_kernel(const double* __restrict__ cst, int batch, double* __restrict__ out) {
  if (threadIdx.x != 0) return;
  double ok = 1.0;
  for (int b = 0; b < batch; b++) if (cst[(int64_t)b * SS2Layout<D>::SIZE + SS2Layout<D>::OK] != 1.0) ok = 0.0;
  out[0] = ok;
}

template <int D>
__global__ void __launch_bounds__(256)
ss2_finish_kernel(const double* __restrict__ part, int nseg, const double* __restrict__ cst, int64_t N, int W, double* __restrict__ lml,
                  double* __restrict__ sums) {
  typedef SS2Layout<D> SL;
  __shared__ double sh[32];
  const int b = blockIdx.x;
  const double* c = cst + (int64_t)b * SL::SIZE;
  const int64_t kstar = (int64_t)c[SL::KSTAR];
  const int payload = SS2_STEPS - W;
  const int64_t used = N <= SS2_STEPS ? 1 : 1 + (N - SS2_STEPS + payload - 1) / payload;     // blocks that did work
  double a0 = 0.0, a1 = 0.0;
  for (int64_t sg = threadIdx.x; sg < used && sg < nseg; sg += blockDim.x) { a0 += part[((int64_t)b * nseg + sg) * 2]; a1 += part[((int64_t)b * nseg + sg) * 2 + 1]; }
  const double r0 = block_sum(a0, sh);
  const double r1 = block_sum(a1, sh);
  if (threadIdx.x == 0) {
    const double slog = r0 + (double)(N > kstar ? N - kstar : 0) * c[SL::ROW + D * D + 2 * D + 1];
    if (lml) lml[b] = -0.5 * ((double)N * 1.8378770664093454835606594728112 + slog + r1);
    if (sums) { sums[2 * b] = slog; sums[2 * b + 1] = r1; }
  }
}

// ---- ss3: the non-head part of the single pass as a PERSISTENT, software-pipelined kernel ---------------------
// ss2_main<false> gives every block one 8192-step window: load it all, then compute — the three co-resident
// blocks of an SM start together, so the SM alternates between "all waiting for HBM" and "all computing"
// (measured 1.8 TB/s, FP64 pipe 39 %), and every block pays the W-step burn-in (14 % re-reads).
// Here a CTA owns ONE contiguous range of a sequence and walks it tile by tile (THREADS sub-chunks of 32 steps):
//   * tiles arrive through an NBUF-deep cp.async ring (16-byte copies, rows padded to 34 doubles so that both the
//     coalesced fill and the per-thread LDS.128 reads are bank-conflict free) — the loads of tiles i+1.. are in
//     flight while tile i is computed;
//   * the state is carried from tile to tile, so the burn-in is paid once per CTA (W / range = 3 % at 1 x 10M);
//   * the emit pass is blocked by 4 steps like the response pass: the four predictions of a group are
//     h_j . x + (ha Phi^m K) y terms, the state advances with Phi^4 — 12.75 instead of 16 FMA per step and one
//     dependent mat-vec per four steps;
//   * one warp scan per tile: the warp's carry enters through the per-lane constant power Phi^(32 (lane+1)).
// Algorithmic traffic: 8 B/step read (+ 8 B/step when alpha is written).
constexpr int SS3_LS = 32, SS3_ROW = SS3_LS + 2;
__device__ __forceinline__ void ss_cp_async16(double* dst, const double* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <int NB> __device__ __forceinline__ void ss_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NB) : "memory"); }

template <int D, int THREADS, int NBUF, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
ss3_main_kernel(const double* __restrict__ y, int64_t ystride, int64_t N, const double* __restrict__ cst, int W, int64_t R, int nseg,
                double* __restrict__ alpha, double* __restrict__ part, SSFin fin) {
  typedef SS2Layout<D> SL;
  constexpr int TILE = THREADS * SS3_LS, BUF = THREADS * SS3_ROW, NW = THREADS / 32, WT = 32 * SS3_LS;
  extern __shared__ __align__(16) double ysm[];              // NBUF buffers of THREADS padded rows
  __shared__ double wtot[2][NW][D], xin[2][D], red[32], pws[SL::NPOW * D * D];
  __shared__ int fin_last;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y, seg = blockIdx.x + 1;
  const double* c = cst + (int64_t)b * SL::SIZE;
  const int64_t r0 = (int64_t)SS2_STEPS + (int64_t)blockIdx.x * R;            // first emitted step of this CTA
  const int64_t r1 = r0 + R < N ? r0 + R : N;                                 // one past its last
  if (r0 >= N) {
    if (tid == 0) { part[((int64_t)b * nseg + seg) * 2] = 0.0; part[((int64_t)b * nseg + seg) * 2 + 1] = 0.0; }
#if __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    ss_ticket_finish<D>(fin, part, nseg, cst, N, &fin_last);
    return;
  }
  const int64_t s0 = r0 - W;                                                  // burn-in start (r0 >= SS2_STEPS > W)
  const int ntile = (int)((r1 - s0 + TILE - 1) / TILE);
  const double* yb = y + (int64_t)b * ystride;
  const bool al16 = ((reinterpret_cast<uintptr_t>(yb + s0) & 15) == 0);
  // Every warp fills (and later drains) only ITS OWN 32 rows = one contiguous 8 KB run of the sequence, so that
  // "tile landed" and "buffer free" are warp-local conditions: no block barrier around the ring.
  auto issue = [&](int ti) {                                                  // rows of this warp in buffer ti % NBUF (zeros beyond r1)
    if (ti < ntile) {
      double* wbuf = ysm + (ti % NBUF) * BUF + wid * 32 * SS3_ROW;
      const int64_t base = s0 + (int64_t)ti * TILE + (int64_t)wid * WT;
      if (al16 && base + WT <= r1) {
#pragma unroll
        for (int i = 0; i < SS3_LS / 2; i++) {
          const int q = lane + i * 32;                                        // 16-byte chunk of the warp's run
          ss_cp_async16(wbuf + (q >> 4) * SS3_ROW + ((q & 15) << 1), yb + base + 2 * q);
        }
      } else {
        for (int i = 0; i < SS3_LS; i++) {
          const int e = lane + i * 32;
          double* dst = wbuf + (e >> 5) * SS3_ROW + (e & 31);
          if (base + e < r1) ss_cp_async8(dst, yb + base + e); else *dst = 0.0;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int i = 0; i < NBUF - 1; i++) issue(i);
  // programmatic dependent launch: everything above needs only y; the set-up kernel's constants (and the zeroed
  // ticket) are complete and visible after this wait
#if __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  // constants (once per CTA): 4-step blocking of the state and of the predictions, scan powers
  double P4[D * D], G[4][D], hj[4][D], cm[3], PL[D * D];
  const double rs = c[SL::ROW + D * D + 2 * D];
  {
    double Phi[D * D];
#pragma unroll
    for (int i = 0; i < D * D; i++) { Phi[i] = c[SL::ROW + i]; P4[i] = c[SL::PHI4 + i]; }
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int i = 0; i < D; i++) G[q][i] = c[SL::G4 + q * D + i];
#pragma unroll
    for (int i = 0; i < D; i++) hj[0][i] = c[SL::ROW + D * D + D + i];
#pragma unroll
    for (int j = 1; j < 4; j++)                                                // h_(j+1) = h_j Phi
#pragma unroll
      for (int i = 0; i < D; i++) { double v = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) v = fma(hj[j - 1][q], Phi[q * D + i], v);
        hj[j][i] = v; }
#pragma unroll
    for (int m = 0; m < 3; m++) { double v = 0.0;                              // c_m = ha Phi^m K = ha . G[3 - m]
#pragma unroll
      for (int q = 0; q < D; q++) v = fma(hj[0][q], G[3 - m][q], v);
      cm[m] = v; }
#pragma unroll
    for (int m = 0; m < 3; m++) cm[m] *= -rs;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int i = 0; i < D; i++) hj[j][i] *= -rs;
    // PL = Phi^(32 (lane + 1)) from the binary powers Phi^(32 2^j)
#pragma unroll
    for (int i = 0; i < D * D; i++) PL[i] = (i / D == i % D) ? 1.0 : 0.0;
    for (int j = 0; j < SL::NPOW; j++)
      if (((lane + 1) >> j) & 1) {
        double Pj[D * D], T[D * D];
#pragma unroll
        for (int i = 0; i < D * D; i++) Pj[i] = c[SL::POWS + j * D * D + i];
        matmul<D>(PL, Pj, T);
#pragma unroll
        for (int i = 0; i < D * D; i++) PL[i] = T[i];
      }
  }
  for (int i = tid; i < SL::NPOW * D * D; i += THREADS) pws[i] = c[SL::POWS + i];
  if (tid < D) { xin[0][tid] = 0.0; xin[1][tid] = 0.0; }
  __syncthreads();
  double a2s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int ti = 0; ti < ntile; ti++) {
    ss_cp_wait<NBUF - 2>();
    __syncwarp();                                             // the warp's rows of tile ti landed; it is done with its rows of tile ti - 1
    issue(ti + NBUF - 1);
    double* row = ysm + (ti % NBUF) * BUF + tid * SS3_ROW;
    const int64_t k0 = s0 + (int64_t)ti * TILE + (int64_t)tid * SS3_LS;
    // phase 1: zero-start response of the sub-chunk (4 steps per dependent mat-vec)
    double rsp[D];
#pragma unroll
    for (int i = 0; i < D; i++) rsp[i] = 0.0;
#pragma unroll
    for (int j = 0; j < SS3_LS; j += 4) {
      const double2 ya = *reinterpret_cast<const double2*>(row + j), yc = *reinterpret_cast<const double2*>(row + j + 2);
      double nx[D];
#pragma unroll
      for (int i = 0; i < D; i++) {
        double u = G[0][i] * ya.x;                            // (independent of the state: off the dependent chain)
        u = fma(G[1][i], ya.y, u); u = fma(G[2][i], yc.x, u); u = fma(G[3][i], yc.y, u);
#pragma unroll
        for (int q = 0; q < D; q++) u = fma(P4[i * D + q], rsp[q], u);
        nx[i] = u;
      }
#pragma unroll
      for (int i = 0; i < D; i++) rsp[i] = nx[i];
    }
    // phase 2: inclusive warp scan with the constant powers, warp carries, per-lane fix-up
#pragma unroll
    for (int j = 0; j < 5; j++) {
      double o[D], pj[D * D];
#pragma unroll
      for (int i = 0; i < D * D; i++) pj[i] = pws[j * D * D + i];
#pragma unroll
      for (int i = 0; i < D; i++) o[i] = __shfl_up_sync(0xffffffffu, rsp[i], 1 << j);
      if (lane >= (1 << j)) {
#pragma unroll
        for (int i = 0; i < D; i++) { double a = rsp[i];
#pragma unroll
          for (int q = 0; q < D; q++) a = fma(pj[i * D + q], o[q], a);
          rsp[i] = a; }
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int i = 0; i < D; i++) wtot[ti & 1][wid][i] = rsp[i];
    }
    __syncthreads();                                          // the only block barrier of a tile (wtot, xin: parity double-buffered)
    double carry[D];                                          // state at the start of this warp's first sub-chunk
#pragma unroll
    for (int i = 0; i < D; i++) carry[i] = xin[ti & 1][i];
    for (int w2 = 0; w2 < wid; w2++) {
      double nx[D];
#pragma unroll
      for (int i = 0; i < D; i++) { double a = wtot[ti & 1][w2][i];
#pragma unroll
        for (int q = 0; q < D; q++) a = fma(pws[5 * D * D + i * D + q], carry[q], a);
        nx[i] = a; }
#pragma unroll
      for (int i = 0; i < D; i++) carry[i] = nx[i];
    }
#pragma unroll
    for (int i = 0; i < D; i++) { double a = rsp[i];          // true state at the end of sub-chunk tid
#pragma unroll
      for (int q = 0; q < D; q++) a = fma(PL[i * D + q], carry[q], a);
      rsp[i] = a; }
    if (tid == THREADS - 1) {
#pragma unroll
      for (int i = 0; i < D; i++) xin[(ti + 1) & 1][i] = rsp[i];
    }
    double x[D];                                              // true state at its start
#pragma unroll
    for (int i = 0; i < D; i++) { x[i] = __shfl_up_sync(0xffffffffu, rsp[i], 1); if (lane == 0) x[i] = carry[i]; }
    // phase 3: payload sub-chunks emit, again four steps per dependent mat-vec
    const bool pay = k0 >= r0 && k0 < r1;
    if (pay) {
      const int nv = r1 - k0 >= SS3_LS ? SS3_LS : (int)(r1 - k0);
#pragma unroll
      for (int j = 0; j < SS3_LS; j += 4) {
        const double2 ya = *reinterpret_cast<const double2*>(row + j), yc = *reinterpret_cast<const double2*>(row + j + 2);
        // alpha_k = rs (y_k - pred_k) with rs and the sign folded into the constants (hj, cm hold -rs h_j, -rs c_m)
        double al[4] = {rs * ya.x, rs * ya.y, rs * yc.x, rs * yc.y};
        al[1] = fma(cm[0], ya.x, al[1]); al[2] = fma(cm[1], ya.x, al[2]); al[3] = fma(cm[2], ya.x, al[3]);
        al[2] = fma(cm[0], ya.y, al[2]); al[3] = fma(cm[1], ya.y, al[3]); al[3] = fma(cm[0], yc.x, al[3]);
#pragma unroll
        for (int q = 0; q < D; q++) {
#pragma unroll
          for (int r = 0; r < 4; r++) al[r] = fma(hj[r][q], x[q], al[r]);
        }
        if (alpha) { *reinterpret_cast<double2*>(row + j) = make_double2(al[0], al[1]); *reinterpret_cast<double2*>(row + j + 2) = make_double2(al[2], al[3]); }
        if (nv != SS3_LS) {
#pragma unroll
          for (int q = 0; q < 4; q++) if (j + q >= nv) al[q] = 0.0;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) a2s[q] = fma(al[q], al[q], a2s[q]);      // four independent accumulators
        double nx[D];
#pragma unroll
        for (int i = 0; i < D; i++) {
          double u = G[0][i] * ya.x;
          u = fma(G[1][i], ya.y, u); u = fma(G[2][i], yc.x, u); u = fma(G[3][i], yc.y, u);
#pragma unroll
          for (int q = 0; q < D; q++) u = fma(P4[i * D + q], x[q], u);
          nx[i] = u;
        }
#pragma unroll
        for (int i = 0; i < D; i++) x[i] = nx[i];
      }
    }
    if (alpha) {                                              // alpha leaves through the warp's own rows, coalesced
      __syncwarp();
      double* ab = alpha + (int64_t)b * ystride;
      const double* wbuf = ysm + (ti % NBUF) * BUF + wid * 32 * SS3_ROW;
      const int64_t base = s0 + (int64_t)ti * TILE + (int64_t)wid * WT;
      if (al16 && base >= r0 && base + WT <= r1) {
#pragma unroll 4
        for (int i = 0; i < SS3_LS / 2; i++) {
          const int q = lane + i * 32;
          *reinterpret_cast<double2*>(ab + base + 2 * q) = *reinterpret_cast<const double2*>(wbuf + (q >> 4) * SS3_ROW + ((q & 15) << 1));
        }
      } else {
        for (int i = 0; i < SS3_LS; i++) {
          const int e = lane + i * 32;
          const int64_t k = base + e;
          if (k >= r0 && k < r1) ab[k] = wbuf[(e >> 5) * SS3_ROW + (e & 31)];
        }
      }
    }
  }
  ss_cp_wait<0>();
  const double a2tot = block_sum((a2s[0] + a2s[1]) + (a2s[2] + a2s[3]), red);
  if (tid == 0) { part[((int64_t)b * nseg + seg) * 2] = 0.0; part[((int64_t)b * nseg + seg) * 2 + 1] = a2tot; }
  ss_ticket_finish<D>(fin, part, nseg, cst, N, &fin_last);
}

// all sequences in one launch: warp w sums the partials of sequence w in fixed order; thread 0 folds the flags
template <int D>
__global__ void __launch_bounds__(512)
ss3_finish_kernel(const double* __restrict__ part, int nseg, const double* __restrict__ cst, int64_t N, int batch,
                  double* __restrict__ lml, double* __restrict__ sums, double* __restrict__ flag_out) {
  typedef SS2Layout<D> SL;
  const int b = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (b < batch) {
    const double* c = cst + (int64_t)b * SL::SIZE;
    double a0 = 0.0, a1 = 0.0;
    for (int sg = lane; sg < nseg; sg += 32) { a0 += part[((int64_t)b * nseg + sg) * 2]; a1 += part[((int64_t)b * nseg + sg) * 2 + 1]; }
    for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    if (lane == 0) {
      const int64_t kstar = (int64_t)c[SL::KSTAR];
      const double slog = a0 + (double)(N > kstar ? N - kstar : 0) * c[SL::ROW + D * D + 2 * D + 1];
      if (lml) lml[b] = -0.5 * ((double)N * 1.8378770664093454835606594728112 + slog + a1);
      if (sums) { sums[2 * b] = slog; sums[2 * b + 1] = a1; }
    }
  }
  if (threadIdx.x == 0) {
    double ok = 1.0;
    for (int q = 0; q < batch; q++) if (cst[(int64_t)q * SL::SIZE + SL::OK] != 1.0) ok = 0.0;
    flag_out[0] = ok;
  }
}

template <int D, int THREADS, int NBUF, int MINB>
int ss3_launch(gpar_ctx* ctx, int batch, int64_t N, const double* y, int64_t ystride, const double* cst, int W, double* alpha, double* part,
               int nc, int64_t R, SSFin fin, bool pdl) {
  const size_t smem = (size_t)NBUF * THREADS * SS3_ROW * sizeof(double);
  CU(cudaFuncSetAttribute((ss3_main_kernel<D, THREADS, NBUF, MINB>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nc, batch); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  CU(cudaLaunchKernelEx(&cfg, ss3_main_kernel<D, THREADS, NBUF, MINB>, y, ystride, N, cst, W, R, nc + 1, alpha, part, fin));
  ctx->launches++;
  return GPAR_OK;
}

template <int D>
int lgssm_run_steady_long(gpar_ctx* ctx, SeqParams sp, int batch, int64_t N, const double* y, const LgssmOut& o, bool* used) {
  typedef SS2Layout<D> SL;
  *used = false;
  int64_t min_n = 32 * (int64_t)SS2_STEPS; int max_batch = 16;
  if (const char* e = getenv("GPAR_SS_LONG_MIN_N")) min_n = std::max<int64_t>(atoll(e), SS2_STEPS + 2 * SS2_WFIX);      // tuning knobs
  if (const char* e = getenv("GPAR_SS_LONG_MAX_BATCH")) max_batch = atoi(e);
  if (!(sp.reg_dt > 0.0) || batch > max_batch || N < min_n || o.mean || o.table || o.dlml || o.dalpha || o.dtable) return GPAR_OK;
  if (const char* e = getenv("GPAR_KF_STEADY")) { if (atoi(e) != 1 && atoi(e) != 3) return GPAR_OK; }   // 0: off, 2: two-pass only
  if (ctx->ss_skip > 0 || !ctx->ss_deferred_ok) return GPAR_OK;   // (ss_skip is decremented by the two-pass path's own check)
  CU(ctx->kal_f.reserve(((size_t)batch * (SL::SIZE + SS2_STEPS) + 8) * sizeof(double)));
  double* cst = ctx->kal_f.as<double>(); double* scratchS = cst + (size_t)batch * SL::SIZE;
  int* ticket = reinterpret_cast<int*>(cst + (size_t)batch * (SL::SIZE + SS2_STEPS) + 4);      // (the flag double sits at + 0)
  LAUNCH(ctx, ss2_setup_kernel<D>, (batch + 31) / 32, 32, 0, sp, batch, cst, ticket);
  // Fixed burn-in of SS2_WFIX steps (14 % re-reads): the grid layout then does not depend on the model, so no
  // host round trip separates the set-up from the main pass; a model that needs more (or whose covariance
  // has not settled) is FLAGGED by the kernels and the caller falls back after its own final synchronisation.
  const int W = SS2_WFIX;
  // variant of the non-head pass: 0 = ss2 (one window per block), 1..3 = ss3 (persistent, pipelined)
  // (measured, 10M steps: one sequence 61 us with 1 CTA of 256 threads per SM, 65 / 75 us with 3 / 2 CTAs of 128;
  //  eight sequences 200 us with 3 CTAs per SM, 216 / 224 us with 2 / 1 — the window-per-block pass took 72 / 340 us)
  int variant = batch <= 2 ? 3 : 2;
  if (const char* e = getenv("GPAR_SS3_VARIANT")) variant = atoi(e);
  const int payload = SS2_STEPS - W;
  int nseg, nc = 0; int64_t R = 0;
  if (variant == 0) nseg = (int)(1 + (N - SS2_STEPS + payload - 1) / payload);
  else {
    // CTAs per SM of the variant; every sequence gets the same number of CTAs, one slot per sequence is left to the head blocks
    const int per_sm = variant == 1 ? 2 : (variant == 2 ? 3 : 1);
    nc = std::max(1, (ctx->num_sms * per_sm - batch) / batch);
    R = ((N - SS2_STEPS + nc - 1) / nc + SS3_LS - 1) / SS3_LS * SS3_LS;
    if (R < 4 * W) { R = 4 * W; nc = (int)((N - SS2_STEPS + R - 1) / R); }      // keep the burn-in below a quarter of a CTA's range
    nseg = nc + 1;
  }
  CU(ctx->kal_b.reserve((size_t)batch * nseg * 2 * sizeof(double)));
  double* part = ctx->kal_b.as<double>();
  const size_t smem = (size_t)(SS2_STEPS + SS2_THREADS) * sizeof(double);
  CU(cudaFuncSetAttribute((ss2_main_kernel<D, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(cudaFuncSetAttribute((ss2_main_kernel<D, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ystride = o.ystride > 0 ? o.ystride : N;
  // the head blocks (sequential transient) on the side stream, underneath all the other blocks
  cudaStream_t main_stream = ctx->stream;
  CU(cudaEventRecord(ctx->ev_fork, main_stream));
  CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
  ctx->stream = ctx->stream2;
  double* flag_dev = ctx->kal_f.as<double>() + (size_t)batch * (SL::SIZE + SS2_STEPS);
  bool fused_finish = variant != 0, pdl = variant != 0;
  if (const char* e = getenv("GPAR_SS3_FUSED")) { fused_finish = fused_finish && (atoi(e) & 1); pdl = pdl && (atoi(e) & 2); }   // bit 0: ticket epilogue, bit 1: PDL
  SSFin fin{fused_finish ? ticket : nullptr, (nc + 1) * batch, batch, o.lml, o.sums, flag_dev};
  const SSFin nofin{nullptr, 0, 0, nullptr, nullptr, nullptr};
  int rc = [&]() -> int {
    LAUNCH(ctx, (ss2_main_kernel<D, true>), dim3(1, batch), SS2_THREADS, smem, y, ystride, N, sp, cst, W, nseg, o.alpha, part, scratchS, fin);
    return GPAR_OK;
  }();
  cudaEventRecord(ctx->ev_side, ctx->stream2);
  ctx->stream = main_stream;
  CHK(rc);
  if (!ctx->pinned) { CU(cudaMallocHost(&ctx->pinned, 4096)); ctx->pinned_cap = 4096; }
  if (variant == 0) {
    if (nseg > 1) LAUNCH(ctx, (ss2_main_kernel<D, false>), dim3(nseg - 1, batch), SS2_THREADS, smem, y, ystride, N, sp, cst, W, nseg, o.alpha, part, scratchS, nofin);
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
    LAUNCH(ctx, ss2_finish_kernel<D>, batch, 256, 0, part, nseg, cst, N, W, o.lml, o.sums);
    LAUNCH(ctx, ss2_flags_kernel<D>, 1, 32, 0, cst, batch, flag_dev);
  } else {
    if (variant == 1) CHK((ss3_launch<D, 128, 3, 2>(ctx, batch, N, y, ystride, cst, W, o.alpha, part, nc, R, fin, pdl)));
    else if (variant == 2) CHK((ss3_launch<D, 128, 2, 3>(ctx, batch, N, y, ystride, cst, W, o.alpha, part, nc, R, fin, pdl)));
    else CHK((ss3_launch<D, 256, 3, 1>(ctx, batch, N, y, ystride, cst, W, o.alpha, part, nc, R, fin, pdl)));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
    if (!fused_finish) LAUNCH(ctx, ss3_finish_kernel<D>, 1, 512, 0, part, nseg, cst, N, batch, o.lml, o.sums, flag_dev);
  }
  // flags (set-up: doubling converged / burn-in long enough; head block: transient settled) go to pinned host
  // memory WITHOUT a synchronisation: lgssm_steady_failed() reads them after the caller's final sync
  CU(cudaMemcpyAsync(ctx->pinned, flag_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->ss_pending = true;
  *used = true;
  return GPAR_OK;
}

constexpr int64_t SS_KTR = 2048;       // steps given to the general scan before the hand-over

// Returns GPAR_OK and *used = true when every sequence had converged at the hand-over (results are then
// in o.lml / o.alpha / o.sums); *used = false means "not applicable or not converged": run the general path.
template <int D>
int lgssm_run_steady(gpar_ctx* ctx, SeqParams sp, int batch, int64_t N, const double* t, const double* y, const LgssmOut& o, bool* used) {
  typedef SSLayout<D> SL;
  typedef AffineElem<D> AE;
  *used = false;
  if (!(sp.reg_dt > 0.0) || N < 4 * SS_KTR || o.mean || o.table || o.dlml || o.dalpha || o.dtable) return GPAR_OK;
  if (const char* e = getenv("GPAR_KF_STEADY")) { if (atoi(e) != 2 && atoi(e) != 3) return GPAR_OK; }   // 0: off, 1: single-pass only
  if (ctx->ss_skip > 0) { ctx->ss_skip--; return GPAR_OK; }
  const int64_t Ns = N - SS_KTR;
  // chunk length (multiple of 32): a block's time grows with L and the grid runs in ceil(blocks / resident) waves
  int L = 64;
  {
    const double resident = (double)ctx->num_sms * 5.0;       // 5 blocks/SM (registers; 33 KB shared tile per block)
    double best = 1e300;
    for (int cand = 64; cand <= 512; cand += 32) {
      const int64_t ncs = (Ns + cand - 1) / cand;
      const double blocks = (double)batch * (double)((ncs + SS_WARPS * 32 - 1) / (SS_WARPS * 32));
      const double cost = std::ceil(blocks / resident) * (cand + 40.0);   // + fixed per-block overhead
      if (cost < best) { best = cost; L = cand; }
    }
  }
  if (const char* e = getenv("GPAR_KF_SSL")) { int v = atoi(e); if (v >= 32 && v <= 1024 && v % 32 == 0) L = v; }   // tuning knob
  const int nC = (int)((Ns + L - 1) / L);
  // transient: the general scan on the first SS_KTR steps
  CU(ctx->kal_f.reserve(((size_t)batch * (D + NSYM<D> + SL::SIZE + 2 + 2 + 1) + 8) * sizeof(double)));
  double* fstate = ctx->kal_f.as<double>(); double* ssc = fstate + (size_t)batch * (D + NSYM<D>);
  double* sums_tr = ssc + (size_t)batch * SL::SIZE; double* sums_ss = sums_tr + 2 * (size_t)batch; double* flags = sums_ss + 2 * (size_t)batch;
  LgssmOut tr; tr.alpha = o.alpha; tr.sums = sums_tr; tr.fstate = fstate; tr.ystride = o.ystride > 0 ? o.ystride : N;
  CHK((lgssm_run_d<D, double>(ctx, sp, batch, SS_KTR, t, y, nullptr, tr)));
  LAUNCH(ctx, ss_setup_kernel<D>, (batch + 127) / 128, 128, 0, sp, batch, fstate, L, ssc);
  // steady part (kal_a is free again: the transient's scan levels are dead)
  LevelPlan ap = plan_levels(nC, AE::NFD, batch);
  const int nslice = std::max(1, std::min(64, (nC + 2047) / 2048));
  const size_t part_doubles = ((size_t)batch * nC + (size_t)batch * nslice) * 2;
  CU(ctx->kal_b.reserve((ap.doubles + part_doubles) * sizeof(double)));
  double* base = ctx->kal_b.as<double>();
  bind_levels(ap, base, AE::NFD, batch);
  double* part = base + ap.doubles; double* part2 = part + (size_t)batch * nC * 2;
  const Level none{nullptr, 0, 0};
  const Level a0 = ap.lv[0], a1 = ap.lv.size() > 1 ? ap.lv[1] : none;
  const int64_t ystride = tr.ystride;
  dim3 g1((a0.P + SS_WARPS * 32 - 1) / (SS_WARPS * 32), batch);
  LAUNCH(ctx, (ss_chunk_kernel<D, false>), g1, SS_WARPS * 32, 0, y, ystride, SS_KTR, N, L, nC, ssc, a0, a1, batch, (double*)nullptr, (double*)nullptr);
  CHK(run_scan<AE>(ctx, ap, batch));
  dim3 g2((nC + SS_WARPS * 32 - 1) / (SS_WARPS * 32), batch);
  LAUNCH(ctx, (ss_chunk_kernel<D, true>), g2, SS_WARPS * 32, 0, y, ystride, SS_KTR, N, L, nC, ssc, a0, a1, batch, o.alpha, part);
  LAUNCH(ctx, lml_partial_kernel<1>, dim3(nslice, batch), 256, 0, part, nC, nslice, part2);
  LAUNCH(ctx, lml_reduce_kernel<1>, batch, 64, 0, part2, nslice, N, (double*)nullptr, (double*)nullptr, sums_ss);
  LAUNCH(ctx, ss_finish_kernel<D>, (batch + 127) / 128, 128, 0, sums_tr, sums_ss, ssc, N, SS_KTR, batch, o.lml, o.sums, flags);
  std::vector<double> hf(batch);
  CU(cudaMemcpyAsync(hf.data(), flags, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b < batch; b++) if (hf[b] != 1.0) { ctx->ss_skip = 8; return GPAR_OK; }   // slow model: general path, and for the next calls too
  *used = true;
  return GPAR_OK;
}

int upload_params(gpar_ctx* ctx, const double* hl, const double* hs, const double* hn, int nparam, SeqParams* sp) {
  CU(ctx->kal_c.reserve((size_t)3 * nparam * sizeof(double)));
  double* dp = ctx->kal_c.as<double>();
  // one copy instead of three (each is ~5 us of host time on the latency path); from the caller's pinned staging when it
  // provides one (the fused small-problem sequence replays this copy inside a CUDA graph)
  std::vector<double> pack_local;
  double* pack = ctx->param_staging;
  if (!pack) { pack_local.resize((size_t)3 * nparam); pack = pack_local.data(); }
  for (int i = 0; i < nparam; i++) { pack[i] = hl[i]; pack[nparam + i] = hs[i]; pack[2 * (size_t)nparam + i] = hn[i]; }
  CU(cudaMemcpyAsync(dp, pack, (size_t)3 * nparam * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  *sp = SeqParams{dp, dp + nparam, dp + 2 * nparam, nparam, -1, -1, -1, 0.0};
  return GPAR_OK;
}

}  // namespace

// Device-level entry used by the ABI functions below and by the scaled-GPAR path.
// params: host arrays (nparam = 1 or batch) of positive (l, s, noise).
int lgssm_run(gpar_ctx* ctx, int kind, const double* hl, const double* hs, const double* hn, int nparam, int batch, int64_t N,
              const double* t, const double* y, const double* rvec, double* d_alpha, double* d_lml, double* d_mean, double* d_var,
              double* d_table, double* d_sums, double* d_table_fwd) {
  if (N < 1 || batch < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm: need at least one time step and one sequence");
  if (d_table_fwd && !d_mean) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm: a separate forward table belongs to a smoother run");
  if ((d_table || d_table_fwd) && batch != 1 && !(ctx->y_broadcast && !d_table_fwd)) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm: step table needs batch == 1");
  SeqParams sp;
  CHK(upload_params(ctx, hl, hs, hn, nparam, &sp));
  if (t == ctx->t.as<double>()) sp.reg_dt = ctx->t_reg_dt;
  LgssmOut o; o.alpha = d_alpha; o.lml = d_lml; o.mean = d_mean; o.var = d_var; o.table = d_table; o.sums = d_sums; o.table_fwd = d_table_fwd;
  o.ybroadcast = ctx->y_broadcast;
  // log-pdf only on short sequences: the one-pass path (constant transition, no exponentials) beats the steady-state
  // scheme, whose 2048-step transient is a large share of a 10k-step sequence (1024 x 10k: 0.25 -> 0.13 ms)
  bool short_logpdf = !d_alpha && !d_mean && !d_table && N < 32768;
  if (const char* e = getenv("GPAR_KF_ONEPASS")) short_logpdf = short_logpdf && atoi(e) != 0;
  if (!rvec && sp.reg_dt > 0.0 && !o.ybroadcast && !short_logpdf) {       // regular grid, scalar noise: steady-state path when every model converges early
    bool used = false;
    switch (kind) {      // long sequences: single-pass burn-in scheme
      case GPAR_MATERN12: CHK(lgssm_run_steady_long<1>(ctx, sp, batch, N, y, o, &used)); break;
      case GPAR_MATERN32: CHK(lgssm_run_steady_long<2>(ctx, sp, batch, N, y, o, &used)); break;
      case GPAR_MATERN52: CHK(lgssm_run_steady_long<3>(ctx, sp, batch, N, y, o, &used)); break;
      default: break;
    }
    if (used) return GPAR_OK;
    switch (kind) {
      case GPAR_MATERN12: CHK(lgssm_run_steady<1>(ctx, sp, batch, N, t, y, o, &used)); break;
      case GPAR_MATERN32: CHK(lgssm_run_steady<2>(ctx, sp, batch, N, t, y, o, &used)); break;
      case GPAR_MATERN52: CHK(lgssm_run_steady<3>(ctx, sp, batch, N, t, y, o, &used)); break;
      default: break;
    }
    if (used) return GPAR_OK;
  }
  switch (kind) {
    case GPAR_MATERN12: return lgssm_run_d<1, double>(ctx, sp, batch, N, t, y, rvec, o);
    case GPAR_MATERN32: return lgssm_run_d<2, double>(ctx, sp, batch, N, t, y, rvec, o);
    case GPAR_MATERN52: return lgssm_run_d<3, double>(ctx, sp, batch, N, t, y, rvec, o);
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "kernel code %d has no state-space form (use Matern12/32/52)", kind);
  }
}

// Forward-mode run: values plus two tangents in ONE pass.  dirs = tangent slot (0, 1 or -1) of
// (l, s, noise).  Outputs: d_lml (batch), d_dlml (batch x 2), d_sums (batch x 6: sum log S, its two
// tangents, sum alpha^2, its two tangents), optionally alpha, d alpha (2 x batch x N) and, for
// batch == 1, the step table and its tangent rows (layout: kf_chunk_filter_kernel).
int lgssm_run_tangent(gpar_ctx* ctx, int kind, const double* hl, const double* hs, const double* hn, int nparam, int batch, int64_t N,
                      const double* t, const double* y, const double* rvec, const int dirs[3],
                      double* d_alpha, double* d_lml, double* d_dlml, double* d_sums, double* d_dalpha, double* d_table, double* d_dtable) {
  if (N < 1 || batch < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm: need at least one time step and one sequence");
  if ((d_table || d_dtable) && batch != 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm: step table needs batch == 1");
  SeqParams sp;
  CHK(upload_params(ctx, hl, hs, hn, nparam, &sp));
  sp.dir_l = dirs[0]; sp.dir_s = dirs[1]; sp.dir_n = dirs[2];
  if (t == ctx->t.as<double>()) sp.reg_dt = ctx->t_reg_dt;
  LgssmOut o; o.alpha = d_alpha; o.lml = d_lml; o.dlml = d_dlml; o.sums = d_sums; o.dalpha = d_dalpha; o.table = d_table; o.dtable = d_dtable;
  switch (kind) {
    case GPAR_MATERN12: return lgssm_run_d<1, Dual<2>>(ctx, sp, batch, N, t, y, rvec, o);
    case GPAR_MATERN32: return lgssm_run_d<2, Dual<2>>(ctx, sp, batch, N, t, y, rvec, o);
    case GPAR_MATERN52: return lgssm_run_d<3, Dual<2>>(ctx, sp, batch, N, t, y, rvec, o);
    default: return gpar_fail(ctx, GPAR_ERR_INVALID, "kernel code %d has no state-space form (use Matern12/32/52)", kind);
  }
}

namespace {
// After the caller's final stream synchronisation: did the steady-state path flag a model it cannot handle?
bool lgssm_steady_failed(gpar_ctx* ctx) {
  if (!ctx->ss_pending) return false;
  ctx->ss_pending = false;
  if (*reinterpret_cast<const double*>(ctx->pinned) == 1.0) return false;
  ctx->ss_skip = 8;                 // general path now, and for the next calls
  return true;
}
int check_seq(gpar_ctx* ctx, const char* who) {
  if (ctx->Nt < 1) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: times not set (gpar_set_times)", who);
  if (ctx->ybatch < 1 || ctx->Ny != ctx->Nt) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: outputs length %lld != number of times %lld", who, (long long)ctx->Ny, (long long)ctx->Nt);
  if (ctx->has_rvec && ctx->Nr != ctx->Nt) return gpar_fail(ctx, GPAR_ERR_INVALID, "%s: noise vector length %lld != number of times %lld", who, (long long)ctx->Nr, (long long)ctx->Nt);
  return GPAR_OK;
}
}  // namespace

extern "C" {

int gpar_lgssm_logpdf(gpar_ctx* ctx, int kernel, const double* theta, int32_t batch_theta, double* lml) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!theta || !lml) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm_logpdf: theta and lml must not be NULL");
  CHK(check_seq(ctx, "lgssm_logpdf"));
  // one resident sequence, several parameter sets: hyper-parameter CANDIDATES (simplex vertices x restarts of
  // temporal_gp_inference.jl:82) evaluated in one pass, every candidate reading the same y
  const bool candidates = ctx->ybatch == 1 && batch_theta > 1;
  const int batch = candidates ? batch_theta : ctx->ybatch;
  if (batch_theta != 1 && batch_theta != batch) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm_logpdf: batch_theta=%d must be 1 or the outputs batch %d (or any number with ONE resident sequence)", batch_theta, batch);
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
  std::vector<double> hl(batch_theta), hs(batch_theta), hn(batch_theta);
  for (int b = 0; b < batch_theta; b++) { GpParams p = unpack_gp3(theta + 3 * b); hl[b] = p.l; hs[b] = p.s; hn[b] = p.noise; }
  CU(ctx->kal_d.reserve((size_t)batch * sizeof(double)));
  // one model, many sequences: covariance recursion once (smooth_shared.cu) — unless the whole batch is so small that the
  // fixed cost of that path (two transposes, a single-sequence table run) exceeds the one-pass path (1024 x 10k: 0.20 vs 0.15 ms)
  bool shared_path = batch_theta == 1 && batch >= 4 && (double)batch * (double)ctx->Nt > 1.5e7;
  if (const char* e = getenv("GPAR_FILTER_SHARED")) shared_path = batch_theta == 1 && batch >= 4 && atoi(e) != 0;
  for (int attempt = 0; attempt < 2; attempt++) {      // a second pass only when the steady-state path flagged a model it cannot handle
    int rc;
    if (shared_path) {
      rc = lgssm_filter_shared_seqmajor(ctx, kernel, hl[0], hs[0], hn[0], ctx->Nt, batch, ctx->t.as<double>(), ctx->y.as<double>(),
                                        ctx->has_rvec ? ctx->rvec.as<double>() : nullptr, nullptr, ctx->kal_d.as<double>());
    } else {
      ctx->ss_deferred_ok = true; ctx->y_broadcast = candidates;
      rc = lgssm_run(ctx, kernel, hl.data(), hs.data(), hn.data(), batch_theta, batch, ctx->Nt, ctx->t.as<double>(), ctx->y.as<double>(),
                     ctx->has_rvec ? ctx->rvec.as<double>() : nullptr, nullptr, ctx->kal_d.as<double>(), nullptr, nullptr, nullptr, nullptr);
      ctx->ss_deferred_ok = false; ctx->y_broadcast = false;
    }
    CHK(rc);
    timer.stop();
    CU(cudaMemcpyAsync(lml, ctx->kal_d.p, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (!lgssm_steady_failed(ctx)) break;
  }
  return GPAR_OK;
}

// logpdf and its gradient with respect to the raw parameters theta (NEW: the reference's optimisers
// are derivative-free).  One forward-mode pass carries two tangents through the whole scan; the third
// derivative follows from the scale identity  s dF/ds + sigma^2 dF/dsigma^2 = -1/2 (N - sum alpha^2)
// (Sigma(c s, c sigma^2) = c Sigma).  With a resident noise vector the tangents are (l, s) and the
// theta noise parameter has no effect (gradient 0).
int gpar_lgssm_logpdf_grad(gpar_ctx* ctx, int kernel, const double* theta, int32_t batch_theta, double* lml, double* grad) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!theta || !lml || !grad) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm_logpdf_grad: theta, lml and grad must not be NULL");
  CHK(check_seq(ctx, "lgssm_logpdf_grad"));
  const int batch = ctx->ybatch; const int64_t N = ctx->Nt;
  if (batch_theta != 1 && batch_theta != batch) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm_logpdf_grad: batch_theta=%d must be 1 or the outputs batch %d", batch_theta, batch);
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
  std::vector<GpParams> ps(batch_theta);
  std::vector<double> hl(batch_theta), hs(batch_theta), hn(batch_theta);
  for (int b = 0; b < batch_theta; b++) { ps[b] = unpack_gp3(theta + 3 * b); hl[b] = ps[b].l; hs[b] = ps[b].s; hn[b] = ps[b].noise; }
  const bool rv = ctx->has_rvec;
  const int dirs[3] = {0, rv ? 1 : -1, rv ? -1 : 1};
  CU(ctx->kal_d.reserve((size_t)batch * 9 * sizeof(double)));
  double* d_lml = ctx->kal_d.as<double>(); double* d_dlml = d_lml + batch; double* d_sums = d_dlml + 2 * (size_t)batch;
  CHK(lgssm_run_tangent(ctx, kernel, hl.data(), hs.data(), hn.data(), batch_theta, batch, N, ctx->t.as<double>(), ctx->y.as<double>(),
                        rv ? ctx->rvec.as<double>() : nullptr, dirs, nullptr, d_lml, d_dlml, d_sums, nullptr, nullptr, nullptr));
  timer.stop();
  std::vector<double> h((size_t)batch * 9);
  CU(cudaMemcpyAsync(h.data(), d_lml, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b < batch; b++) {
    const GpParams& p = ps[batch_theta == 1 ? 0 : b];
    lml[b] = h[b];
    const double d0 = h[batch + 2 * b], d1 = h[batch + 2 * b + 1], a2 = h[3 * (size_t)batch + 6 * b + 3];
    double dl = d0, ds, dn;
    if (rv) { ds = d1; dn = 0.0; }
    else { dn = d1; ds = (-0.5 * ((double)N - a2) - p.noise * dn) / p.s; }
    grad[3 * b + 0] = dl * p.dl; grad[3 * b + 1] = ds * p.ds_dv; grad[3 * b + 2] = dn * p.dn;
  }
  return GPAR_OK;
}

int gpar_lgssm_decorrelate(gpar_ctx* ctx, int kernel, const double theta[3], double* alpha, double* lml) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!theta || !alpha) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm_decorrelate: theta and alpha must not be NULL");
  CHK(check_seq(ctx, "lgssm_decorrelate"));
  const int batch = ctx->ybatch; const int64_t N = ctx->Nt;
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
  GpParams p = unpack_gp3(theta);
  CU(ctx->kal_d.reserve(((size_t)batch * N + batch) * sizeof(double)));
  double* d_alpha = ctx->kal_d.as<double>(); double* d_lml = d_alpha + (size_t)batch * N;
  bool shared_path = batch >= 4;                          // decorrelate always shares theta
  if (const char* e = getenv("GPAR_FILTER_SHARED")) shared_path = shared_path && atoi(e) != 0;
  for (int attempt = 0; attempt < 2; attempt++) {      // see gpar_lgssm_logpdf
    int rc;
    if (shared_path) {
      rc = lgssm_filter_shared_seqmajor(ctx, kernel, p.l, p.s, p.noise, N, batch, ctx->t.as<double>(), ctx->y.as<double>(),
                                        ctx->has_rvec ? ctx->rvec.as<double>() : nullptr, d_alpha, d_lml);
    } else {
      ctx->ss_deferred_ok = true;
      rc = lgssm_run(ctx, kernel, &p.l, &p.s, &p.noise, 1, batch, N, ctx->t.as<double>(), ctx->y.as<double>(),
                     ctx->has_rvec ? ctx->rvec.as<double>() : nullptr, d_alpha, d_lml, nullptr, nullptr, nullptr, nullptr);
      ctx->ss_deferred_ok = false;
    }
    CHK(rc);
    timer.stop();
    CU(cudaMemcpyAsync(alpha, d_alpha, (size_t)batch * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (lml) CU(cudaMemcpyAsync(lml, d_lml, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (!lgssm_steady_failed(ctx)) break;
  }
  return GPAR_OK;
}

int gpar_lgssm_smooth(gpar_ctx* ctx, int kernel, const double theta[3], double* mean, double* var, double* lml) {
  if (!ctx) return GPAR_ERR_INVALID;
  if (!theta || ((mean == nullptr) != (var == nullptr))) return gpar_fail(ctx, GPAR_ERR_INVALID, "lgssm_smooth: theta must not be NULL; mean and var are given together or both NULL");
  CHK(check_seq(ctx, "lgssm_smooth"));
  const int batch = ctx->ybatch; const int64_t N = ctx->Nt;
  CU(cudaSetDevice(ctx->device));
  CallTimer timer(ctx); ctx->phase_valid = false; gpar_drop_result(ctx);
  GpParams p = unpack_gp3(theta);
  CU(ctx->kal_d.reserve((2 * (size_t)batch * N + batch) * sizeof(double)));
  double* d_mean = ctx->kal_d.as<double>(); double* d_var = d_mean + (size_t)batch * N; double* d_lml = d_var + (size_t)batch * N;
  // several sequences always share theta here: covariances and gains once, then two affine passes per sequence
  bool shared_path = batch >= 4;
  if (const char* e = getenv("GPAR_SMOOTH_SHARED")) shared_path = atoi(e) != 0 && batch >= 2;      // tuning / testing knob
  if (shared_path) {
    CHK(lgssm_smooth_shared_seqmajor(ctx, kernel, p.l, p.s, p.noise, N, batch, ctx->t.as<double>(), ctx->y.as<double>(),
                                     ctx->has_rvec ? ctx->rvec.as<double>() : nullptr, d_mean, d_var, d_lml));
  } else {
    CHK(lgssm_run(ctx, kernel, &p.l, &p.s, &p.noise, 1, batch, N, ctx->t.as<double>(), ctx->y.as<double>(),
                  ctx->has_rvec ? ctx->rvec.as<double>() : nullptr, nullptr, d_lml, d_mean, d_var, nullptr, nullptr));
  }
  timer.stop();
  ctx->res_a = d_mean; ctx->res_b = d_var; ctx->res_len = (int64_t)batch * N;      // stays resident for gpar_take_test
  if (mean) {
    CU(cudaMemcpyAsync(mean, d_mean, (size_t)batch * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(var, d_var, (size_t)batch * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (lml) CU(cudaMemcpyAsync(lml, d_lml, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GPAR_OK;
}

}  // extern "C"
