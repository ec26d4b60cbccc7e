// lgssm_math.cuh — register-resident small-matrix algebra for the Matern LGSSM kernels.
// State dimension D = 1, 2, 3 (Matern-1/2, 3/2, 5/2).  Symmetric matrices are packed (upper,
// row-major): D=3 -> [00 01 02 11 12 22].
#pragma once
#include "common.cuh"
#include "dual.cuh"

template <int D> struct SymIdx;
template <> struct SymIdx<1> { __host__ __device__ static constexpr int at(int i, int j) { return 0; } };
template <> struct SymIdx<2> { __host__ __device__ static constexpr int at(int i, int j) { return i <= j ? (i == 0 ? j : 2) : (j == 0 ? i : 2); } };
template <> struct SymIdx<3> {
  __host__ __device__ static constexpr int at(int i, int j) {
    return i <= j ? (i == 0 ? j : (i == 1 ? 2 + j : 5)) : (j == 0 ? i : (j == 1 ? 2 + i : 5));
  }
};
template <int D> constexpr int NSYM = D * (D + 1) / 2;
#define SYM(S, i, j) S[SymIdx<D>::at(i, j)]

template <int D> __device__ __forceinline__ double lgssm_lambda() { return D == 1 ? 1.0 : (D == 2 ? 1.7320508075688772 : 2.23606797749979); }

// Stationary unit-variance covariance P_inf (packed) — TemporalGPs stationary_distribution.
template <int D, class F> __device__ __forceinline__ void lgssm_pinf(F* P) {
  if (D == 1) { P[0] = 1.0; }
  else if (D == 2) { P[0] = 1.0; P[1] = 0.0; P[2] = 3.0; }
  else { P[0] = 1.0; P[1] = 0.0; P[2] = -5.0 / 3.0; P[3] = 5.0 / 3.0; P[4] = 0.0; P[5] = 25.0; }
}

// A = exp(F a) = e^{-lam a} (I + N a + N^2 a^2/2),  N = F + lam I  (row-major D x D)
template <int D, class F> __device__ __forceinline__ void lgssm_transition(F a, F* A) {
  const double lam = lgssm_lambda<D>();
  const F e = exp(-lam * a);
  if (D == 1) { A[0] = e; }
  else if (D == 2) {
    // N = [lam 1; -lam^2 -lam]
    A[0] = e * (1.0 + lam * a); A[1] = e * a;
    A[2] = e * (-lam * lam * a); A[3] = e * (1.0 - lam * a);
  } else {
    // N = [lam 1 0; 0 lam 1; -lam^3 -3lam^2 -2lam],  N^2 = [lam^2 2lam 1; -lam^3 -2lam^2 -lam; lam^4 2lam^3 lam^2]
    const double l2 = lam * lam, l3 = l2 * lam, l4 = l2 * l2; const F h = 0.5 * a * a;
    A[0] = e * (1.0 + lam * a + l2 * h);  A[1] = e * (a + 2.0 * lam * h);            A[2] = e * h;
    A[3] = e * (-l3 * h);                 A[4] = e * (1.0 + lam * a - 2.0 * l2 * h);  A[5] = e * (a - lam * h);
    A[6] = e * (-l3 * a + l4 * h);        A[7] = e * (-3.0 * l2 * a + 2.0 * l3 * h);  A[8] = e * (1.0 - 2.0 * lam * a + l2 * h);
  }
}

// Branch-free exp(x) for x <= 0 (clamped at -700, where the result is ~1e-304 and every use of it rounds to zero) and
// branch-free 1/s for normal s > 0.  The library versions carry special-case branches, i.e. basic-block boundaries the
// instruction scheduler cannot move work across; inside the one-pass Kalman loop these two dependent chains are meant
// to overlap the matrix recursion.  Accuracy: ~1 ulp (12-term Taylor on |r| <= ln2/2: truncation 2.4e-16 relative; the
// reciprocal is MUFU.RCP64H refined by one cubic Newton step: measured <= 1 ulp over 4M arguments, tools/bin/rcp_test).
// (constants live in constant memory: as immediates every 64-bit coefficient costs two UMOV issue slots per use, and in
// the one-pass Kalman loop every non-FP64 instruction delays the FP64 pipe by a cycle)
__constant__ double kExpC[18] = {
    1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
    0.0, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
    2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02,
    1.6666666666666666e-01, 0.5, 1.0, -700.0};
__device__ __forceinline__ double exp_nonpos(double x) {
  x = fmax(x, kExpC[17]);
  const double t = fma(x, kExpC[0], kExpC[1]);               // round(x / ln 2) in the low word
  const int n = __double2loint(t);
  const double nf = t - kExpC[1];
  double r = fma(nf, kExpC[2], x);
  r = fma(nf, kExpC[3], r);
  double p = kExpC[5];                                       // 1/12!, then 1/11! ... 1/2!, 1, 1  (truncation <= 2.4e-16 relative)
#pragma unroll
  for (int i = 6; i <= 16; i++) p = fma(p, r, kExpC[i]);
  p = fma(p, r, kExpC[16]);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));      // p in [0.70, 1.42], n >= -1010: stays normal
}
__device__ __forceinline__ double rcp_pos(double s) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
  double e = fma(-s, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);           // seed error <= 9.9e-7 (measured, tools/rcp_test.cu) -> cubic step: <= 2.3e-16
}

template <int NT> __device__ __forceinline__ Dual<NT> exp_nonpos(const Dual<NT>& x) {       // (the clamp only acts where the value is ~0)
  Dual<NT> r; r.v = exp_nonpos(x.v);
#pragma unroll
  for (int i = 0; i < NT; i++) r.d[i] = r.v * x.d[i];
  return r;
}
template <int NT> __device__ __forceinline__ Dual<NT> rcp_pos(const Dual<NT>& s) {
  Dual<NT> r; r.v = rcp_pos(s.v);
  const double f = -r.v * r.v;
#pragma unroll
  for (int i = 0; i < NT; i++) r.d[i] = f * s.d[i];
  return r;
}
// x = m 2^e with m in [1, 2): x <- m (tangents scaled alike, so log x keeps its tangent), e added to eacc.  x > 0 normal.
__device__ __forceinline__ void peel_exponent(double& x, int& eacc) {
  const int hi = __double2hiint(x);
  eacc += (hi >> 20) - 1023;
  x = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
}
template <int NT> __device__ __forceinline__ void peel_exponent(Dual<NT>& x, int& eacc) {
  const int e = (__double2hiint(x.v) >> 20) - 1023;
  const double sc = __hiloint2double((1023 - e) << 20, 0);          // 2^-e  (|e| < 1000 by construction)
  eacc += e; x.v *= sc;
#pragma unroll
  for (int i = 0; i < NT; i++) x.d[i] *= sc;
}

// The same matrix from a and e = exp(-lam a) computed elsewhere (the one-pass kernel evaluates the exponential of the
// NEXT step while the current step's recursion runs: it does not depend on the state).
template <int D> __device__ __forceinline__ void lgssm_transition_from(double a, double e, double* A) {
  const double lam = lgssm_lambda<D>();
  if (D == 1) { A[0] = e; }
  else if (D == 2) {
    A[0] = e * (1.0 + lam * a); A[1] = e * a;
    A[2] = e * (-lam * lam * a); A[3] = e * (1.0 - lam * a);
  } else {
    const double l2 = lam * lam, l3 = l2 * lam, l4 = l2 * l2; const double h = 0.5 * a * a;
    A[0] = e * (1.0 + lam * a + l2 * h);  A[1] = e * (a + 2.0 * lam * h);            A[2] = e * h;
    A[3] = e * (-l3 * h);                 A[4] = e * (1.0 + lam * a - 2.0 * l2 * h);  A[5] = e * (a - lam * h);
    A[6] = e * (-l3 * a + l4 * h);        A[7] = e * (-3.0 * l2 * a + 2.0 * l3 * h);  A[8] = e * (1.0 - 2.0 * lam * a + l2 * h);
  }
}

// R = A S A^T  (S symmetric packed; result packed)
template <int D, class F> __device__ __forceinline__ void asat(const F* A, const F* S, F* R) {
  F T[D * D];
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) { F v = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) v = fma(A[i * D + k], SYM(S, k, j), v);
      T[i * D + j] = v; }
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = i; j < D; j++) { F v = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) v = fma(T[i * D + k], A[j * D + k], v);
      SYM(R, i, j) = v; }
}

// ---- Jordan coordinates ---------------------------------------------------------------------------------------------
// The Matern feedback matrix F is a companion matrix with the single eigenvalue mu = -lam; its Jordan chain is the
// confluent Vandermonde basis V = [[1,0,0],[mu,1,0],[mu^2,2mu,1]] (D = 3; leading blocks for D = 2, 1), so with
// z = V^-1 x:   exp(F a) = V e^{-lam a} U(a) V^-1,   U(a) = [[1, a, a^2/2], [0, 1, a], [0, 0, 1]],
// and — V being unit lower triangular with first row e_0' — the observation row stays H = e_0'.  In z the transition
// is a unit upper-triangular Toeplitz matrix times a scalar: U X costs 9 FMA instead of 27, the covariance congruence
// e^2 U S U' 12 FMA + 6 MUL instead of 45 FMA, and no matrix has to be formed from (a, e).  Everything a caller sees
// (log-pdf, innovations, first state component = x_0 = z_0) is invariant under the change of basis; FiltElem / SmoothElem
// combine generic matrices, so elements built in z scan with the same operators.
template <int D, class F> __device__ __forceinline__ void lgssm_pinf_jordan(F s, F* Pz) {     // s V^-1 P_inf V^-T (packed)
  const double lam = lgssm_lambda<D>();
  double P[NSYM<D>]; lgssm_pinf<D>(P);
  if (D == 1) { Pz[0] = s * P[0]; return; }
  double Vi[D * D];
#pragma unroll
  for (int i = 0; i < D * D; i++) Vi[i] = (i / D == i % D) ? 1.0 : 0.0;
  Vi[D] = lam;                                           // V^-1 = [[1,0,0],[lam,1,0],[lam^2,2 lam,1]]
  if (D == 3) { Vi[6] = lam * lam; Vi[7] = 2.0 * lam; }
  double R[NSYM<D>];
  asat<D>(Vi, P, R);
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) Pz[i] = s * R[i];
}
// T = U(a) X  (h = a^2/2; X, T row-major D x D; the last row is X's own)
template <int D, class F> __device__ __forceinline__ void jordan_rows(F a, F h, const F* X, F* T) {
#pragma unroll
  for (int j = 0; j < D; j++) {
    if (D == 1) T[j] = X[j];
    else if (D == 2) { T[j] = fma(a, X[D + j], X[j]); T[D + j] = X[D + j]; }
    else { T[j] = fma(h, X[2 * D + j], fma(a, X[D + j], X[j])); T[D + j] = fma(a, X[2 * D + j], X[D + j]); T[2 * D + j] = X[2 * D + j]; }
  }
}
template <int D, class F> __device__ __forceinline__ void jordan_vec(F a, F h, const F* x, F* u) {
  if (D == 1) u[0] = x[0];
  else if (D == 2) { u[0] = fma(a, x[1], x[0]); u[1] = x[1]; }
  else { u[0] = fma(h, x[2], fma(a, x[1], x[0])); u[1] = fma(a, x[2], x[1]); u[2] = x[2]; }
}
// R = e2 U(a) S U(a)'  (S, R symmetric packed)
template <int D, class F> __device__ __forceinline__ void jordan_congruence(F a, F h, F e2, const F* S, F* R) {
  if (D == 1) { R[0] = e2 * S[0]; }
  else if (D == 2) {
    const F t01 = fma(a, S[2], S[1]), t00 = fma(a, S[1], S[0]);
    R[0] = e2 * fma(a, t01, t00); R[1] = e2 * t01; R[2] = e2 * S[2];
  } else {      // packed [00 01 02 11 12 22]
    const F t02 = fma(h, S[5], fma(a, S[4], S[2]));
    const F t01 = fma(h, S[4], fma(a, S[3], S[1]));
    const F t00 = fma(h, S[2], fma(a, S[1], S[0]));
    const F t12 = fma(a, S[5], S[4]);
    const F t11 = fma(a, S[4], S[3]);
    R[0] = e2 * fma(h, t02, fma(a, t01, t00));
    R[1] = e2 * fma(a, t02, t01);
    R[2] = e2 * t02;
    R[3] = e2 * fma(a, t12, t11);
    R[4] = e2 * t12;
    R[5] = e2 * S[5];
  }
}

// Q = P0 - A P0 A^T   (P0 = s P_inf; for D = 3 the zeros of P_inf are skipped: 33 instead of 45 FMA)
template <int D, class F> __device__ __forceinline__ void lgssm_q(const F* A, const F* P0, F* Q) {
  F R[NSYM<D>];
  if (D == 3) {
    F T[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      T[i * 3 + 0] = fma(A[i * 3 + 2], P0[2], A[i * 3 + 0] * P0[0]);
      T[i * 3 + 1] = A[i * 3 + 1] * P0[3];
      T[i * 3 + 2] = fma(A[i * 3 + 2], P0[5], A[i * 3 + 0] * P0[2]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = i; j < 3; j++) { F v = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) v = fma(T[i * 3 + k], A[j * 3 + k], v);
        SYM(R, i, j) = v; }
  } else {
    asat<D>(A, P0, R);
  }
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) Q[i] = P0[i] - R[i];
}

// Predicted covariance with the process noise folded in:
//   A P A^T + Q,  Q = P0 - A P0 A^T   ==   P0 + A (P - P0) A^T      (one congruence instead of two)
template <int D, class F> __device__ __forceinline__ void predict_cov(const F* A, const F* P, const F* P0, F* Pp) {
  F Dl[NSYM<D>];
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) Dl[i] = P[i] - P0[i];
  asat<D>(A, Dl, Pp);
#pragma unroll
  for (int i = 0; i < NSYM<D>; i++) Pp[i] = Pp[i] + P0[i];
}

template <int D, class F> __device__ __forceinline__ void matvec(const F* A, const F* x, F* y) {
#pragma unroll
  for (int i = 0; i < D; i++) { F v = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) v = fma(A[i * D + k], x[k], v);
    y[i] = v; }
}
template <int D, class F> __device__ __forceinline__ void matmul(const F* A, const F* B, F* C) {
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) { F v = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) v = fma(A[i * D + k], B[k * D + j], v);
      C[i * D + j] = v; }
}
// C = A * S (S symmetric packed) -> full
template <int D, class F> __device__ __forceinline__ void matsym(const F* A, const F* S, F* C) {
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) { F v = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) v = fma(A[i * D + k], SYM(S, k, j), v);
      C[i * D + j] = v; }
}
// C = S * A (S symmetric packed) -> full
template <int D, class F> __device__ __forceinline__ void symmat(const F* S, const F* A, F* C) {
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) { F v = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) v = fma(SYM(S, i, k), A[k * D + j], v);
      C[i * D + j] = v; }
}
template <int D, class F> __device__ __forceinline__ void symvec(const F* S, const F* x, F* y) {
#pragma unroll
  for (int i = 0; i < D; i++) { F v = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) v = fma(SYM(S, i, k), x[k], v);
    y[i] = v; }
}

// general inverse (adjugate); returns via Minv
template <int D, class F> __device__ __forceinline__ void inv_general(const F* M, F* Minv) {
  if (D == 1) { Minv[0] = 1.0 / M[0]; }
  else if (D == 2) {
    F id = 1.0 / (M[0] * M[3] - M[1] * M[2]);
    Minv[0] = M[3] * id; Minv[1] = -M[1] * id; Minv[2] = -M[2] * id; Minv[3] = M[0] * id;
  } else {
    F c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    F id = 1.0 / (M[0] * c00 + M[1] * c01 + M[2] * c02);
    Minv[0] = c00 * id; Minv[3] = c01 * id; Minv[6] = c02 * id;
    Minv[1] = (M[2] * M[7] - M[1] * M[8]) * id; Minv[4] = (M[0] * M[8] - M[2] * M[6]) * id; Minv[7] = (M[1] * M[6] - M[0] * M[7]) * id;
    Minv[2] = (M[1] * M[5] - M[2] * M[4]) * id; Minv[5] = (M[2] * M[3] - M[0] * M[5]) * id; Minv[8] = (M[0] * M[4] - M[1] * M[3]) * id;
  }
}

// X = W * (S + eps I)^{-1} for symmetric positive definite S (packed), through the Cholesky
// factor as TemporalGPs' smooth does (U = cholesky(P + eps I).U; Gt = U \ (U' \ (A P))).
template <int D, class F> __device__ __forceinline__ void solve_spd_right(const F* W, const F* S, double eps, F* X) {
  F L[D * D];
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) {
      F v = SYM(S, i, j) + (i == j ? eps : 0.0);
#pragma unroll
      for (int q = 0; q < j; q++) v -= L[i * D + q] * L[j * D + q];
      L[i * D + j] = (i == j) ? sqrt(v) : v / L[j * D + j];
    }
  // row r of X solves (L L^T) x = W[r,:]^T
#pragma unroll
  for (int r = 0; r < D; r++) {
    F z[D];
#pragma unroll
    for (int i = 0; i < D; i++) { F v = W[r * D + i];
#pragma unroll
      for (int q = 0; q < i; q++) v -= L[i * D + q] * z[q];
      z[i] = v / L[i * D + i]; }
#pragma unroll
    for (int i = D - 1; i >= 0; i--) { F v = z[i];
#pragma unroll
      for (int q = i + 1; q < D; q++) v -= L[q * D + i] * X[r * D + q];
      X[r * D + i] = v / L[i * D + i]; }
  }
}

// C = S1 * S2 (both symmetric packed) -> full
template <int D, class F> __device__ __forceinline__ void symsym(const F* S1, const F* S2, F* C) {
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) { F v = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) v = fma(SYM(S1, i, k), SYM(S2, k, j), v);
      C[i * D + j] = v; }
}

// ---- filtering element (Sarkka & Garcia-Fernandez 2021, temporal parallelisation of Bayesian
//      smoothers): p(x_k | x_{k-1}, y) = N(A x + b, C),  p(y | x_{k-1}) ~ N_info(eta, J) ----------
template <int D, class F = double> struct FiltElem {
  static constexpr int NF = D * D + 2 * D + 2 * NSYM<D>;
  static constexpr int OB = D * D, OC = OB + D, OE = OC + NSYM<D>, OJ = OE + D;
  typedef F scalar_t;
  static constexpr int NFD = NF * Scalar<F>::NC;     // doubles per element (value + tangents)
  F v[NF];
  __device__ __forceinline__ void set_identity() {
    F *A = v, *b = v + OB, *C = v + OC, *eta = v + OE, *J = v + OJ;
#pragma unroll
    for (int i = 0; i < D * D; i++) A[i] = (i / D == i % D) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) { b[i] = 0.0; eta[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) { C[i] = 0.0; J[i] = 0.0; }
  }
  // result = e1 (earlier in time) then e2 (later in time)
  __device__ __forceinline__ static FiltElem combine(const FiltElem& x1, const FiltElem& x2) {
    FiltElem rr;
    struct V { const F *A, *b, *C, *eta, *J; };
    struct W { F *A, *b, *C, *eta, *J; };
    const V e1{x1.v, x1.v + OB, x1.v + OC, x1.v + OE, x1.v + OJ}, e2{x2.v, x2.v + OB, x2.v + OC, x2.v + OE, x2.v + OJ};
    const W r{rr.v, rr.v + OB, rr.v + OC, rr.v + OE, rr.v + OJ};
    F Mx[D * D], Mi[D * D], X[D * D], T[D * D], v[D], u[D];
    symsym<D>(e1.C, e2.J, Mx);
#pragma unroll
    for (int i = 0; i < D; i++) Mx[i * D + i] += 1.0;
    inv_general<D>(Mx, Mi);
    matmul<D>(e2.A, Mi, X);
    matmul<D>(X, e1.A, r.A);
    symvec<D>(e1.C, e2.eta, v);
#pragma unroll
    for (int i = 0; i < D; i++) v[i] += e1.b[i];
    matvec<D>(X, v, u);
#pragma unroll
    for (int i = 0; i < D; i++) r.b[i] = u[i] + e2.b[i];
    matsym<D>(X, e1.C, T);
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) { F a = SYM(e2.C, i, j);
#pragma unroll
        for (int k = 0; k < D; k++) a = fma(T[i * D + k], e2.A[j * D + k], a);
        SYM(r.C, i, j) = a; }
    // Y = A1^T (I + J2 C1)^{-1} = A1^T Mi^T
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = 0; j < D; j++) { F a = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) a = fma(e1.A[k * D + i], Mi[j * D + k], a);
        X[i * D + j] = a; }
    symvec<D>(e2.J, e1.b, v);
#pragma unroll
    for (int i = 0; i < D; i++) v[i] = e2.eta[i] - v[i];
    matvec<D>(X, v, u);
#pragma unroll
    for (int i = 0; i < D; i++) r.eta[i] = u[i] + e1.eta[i];
    matsym<D>(X, e2.J, T);
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
      for (int j = i; j < D; j++) { F a = SYM(e1.J, i, j);
#pragma unroll
        for (int k = 0; k < D; k++) a = fma(T[i * D + k], e1.A[k * D + j], a);
        SYM(r.J, i, j) = a; }
    return rr;
  }
  // scan order == time order
  __device__ __forceinline__ static FiltElem scan_combine(const FiltElem& first, const FiltElem& second) { return combine(first, second); }
};

// ---- smoothing element: m^s_k = E m^s_{k+1} + g,  P^s_k = E P^s_{k+1} E^T + L ---------------
template <int D, class F = double> struct SmoothElem {
  static constexpr int NF = D * D + D + NSYM<D>;
  static constexpr int OG = D * D, OL = OG + D;
  typedef F scalar_t;
  static constexpr int NFD = NF * Scalar<F>::NC;
  F v[NF];
  __device__ __forceinline__ void set_identity() {
    F *E = v, *g = v + OG, *L = v + OL;
#pragma unroll
    for (int i = 0; i < D * D; i++) E[i] = (i / D == i % D) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) g[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) L[i] = 0.0;
  }
  // e1 earlier in time, e2 later in time: (E1 E2, E1 g2 + g1, E1 L2 E1^T + L1)
  __device__ __forceinline__ static SmoothElem combine(const SmoothElem& x1, const SmoothElem& x2) {
    SmoothElem rr;
    struct V { const F *E, *g, *L; };
    struct W { F *E, *g, *L; };
    const V e1{x1.v, x1.v + OG, x1.v + OL}, e2{x2.v, x2.v + OG, x2.v + OL};
    const W r{rr.v, rr.v + OG, rr.v + OL};
    F u[D], R[NSYM<D>];
    matmul<D>(e1.E, e2.E, r.E);
    matvec<D>(e1.E, e2.g, u);
#pragma unroll
    for (int i = 0; i < D; i++) r.g[i] = u[i] + e1.g[i];
    asat<D>(e1.E, e2.L, R);
#pragma unroll
    for (int i = 0; i < NSYM<D>; i++) r.L[i] = R[i] + e1.L[i];
    return rr;
  }
  // the smoother scans over REVERSED time: `first` is later in time
  __device__ __forceinline__ static SmoothElem scan_combine(const SmoothElem& first, const SmoothElem& second) { return combine(second, first); }
};

// ---- affine map x -> M x + r (steady-state mean recursion over a chunk); scan order == time order ----
template <int D, class F = double> struct AffineElem {
  static constexpr int NF = D * D + D;
  static constexpr int OR = D * D;
  typedef F scalar_t;
  static constexpr int NFD = NF * Scalar<F>::NC;
  F v[NF];
  __device__ __forceinline__ void set_identity() {
#pragma unroll
    for (int i = 0; i < D * D; i++) v[i] = (i / D == i % D) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) v[OR + i] = 0.0;
  }
  // first (earlier) then second (later): (M2 M1, M2 r1 + r2)
  __device__ __forceinline__ static AffineElem scan_combine(const AffineElem& e1, const AffineElem& e2) {
    AffineElem r;
    F u[D];
    matmul<D>(e2.v, e1.v, r.v);
    matvec<D>(e2.v, e1.v + OR, u);
#pragma unroll
    for (int i = 0; i < D; i++) r.v[OR + i] = u[i] + e2.v[OR + i];
    return r;
  }
};
