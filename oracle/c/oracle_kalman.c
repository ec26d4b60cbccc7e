/* oracle_kalman.c — plain-C twin of oracle/lgssm.py (TEST INFRASTRUCTURE / CPU BASELINE ONLY).
 *
 * Sequential Kalman filter (`decorrelate` / `logpdf`) and RTS smoother (`smooth`) of the Matern
 * LGSSM, restating TemporalGPs.jl ~0.2 [un-vendored dependency] as the reference uses it:
 * src/gp/temporal_gp_inference.jl:15-39,78,109; src/gp/dtc.jl:101-117;
 * src/gp/gpar_scaled_inference.jl:105-117,163-183.  Fixed-size d x d loops stand in for the
 * StaticArrays code the reference runs on the CPU (a Python loop would understate it).
 * Never linked into, or called by, the product library.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LOG2PI 1.8378770664093454835606594728112
#define DMAX 3

static int kind_dim(int kind) { return kind == 1 ? 1 : kind == 2 ? 2 : kind == 3 ? 3 : 0; }

/* unit-variance stationary covariance P_inf and companion F (row-major) */
static void sde(int kind, double* F, double* P) {
  int d = kind_dim(kind);
  memset(F, 0, sizeof(double) * DMAX * DMAX); memset(P, 0, sizeof(double) * DMAX * DMAX);
  if (d == 1) { F[0] = -1.0; P[0] = 1.0; }
  else if (d == 2) { double lam = sqrt(3.0); F[0 * DMAX + 1] = 1.0; F[1 * DMAX + 0] = -lam * lam; F[1 * DMAX + 1] = -2.0 * lam; P[0] = 1.0; P[1 * DMAX + 1] = 3.0; }
  else { double lam = sqrt(5.0), kap = 5.0 / 3.0;
    F[0 * DMAX + 1] = 1.0; F[1 * DMAX + 2] = 1.0; F[2 * DMAX + 0] = -lam * lam * lam; F[2 * DMAX + 1] = -3.0 * lam * lam; F[2 * DMAX + 2] = -3.0 * lam;
    P[0] = 1.0; P[0 * DMAX + 2] = -kap; P[1 * DMAX + 1] = kap; P[2 * DMAX + 0] = -kap; P[2 * DMAX + 2] = 25.0; }
}

/* A = exp(F a) = e^{-lam a} sum_{j<d} (N a)^j / j!,  N = F + lam I (nilpotent of index d) */
void oracle_transition(int kind, double a, double* A) {
  int d = kind_dim(kind);
  double F[DMAX * DMAX], P[DMAX * DMAX], Nm[DMAX * DMAX], term[DMAX * DMAX], tmp[DMAX * DMAX];
  double lam = d == 1 ? 1.0 : d == 2 ? sqrt(3.0) : sqrt(5.0);
  sde(kind, F, P);
  for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { Nm[i * DMAX + j] = F[i * DMAX + j] + (i == j ? lam : 0.0); term[i * DMAX + j] = (i == j); A[i * DMAX + j] = (i == j); }
  for (int p = 1; p < d; p++) {
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { double s = 0; for (int k = 0; k < d; k++) s += term[i * DMAX + k] * Nm[k * DMAX + j]; tmp[i * DMAX + j] = s * (a / p); }
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { term[i * DMAX + j] = tmp[i * DMAX + j]; A[i * DMAX + j] += tmp[i * DMAX + j]; }
  }
  double e = exp(-lam * a);
  for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) A[i * DMAX + j] *= e;
}

/* One filter pass.  Stores (all nullable): alpha[n]; mf,Pf (filtered), mp,Pp (predicted) with row
 * stride d / d*d.  Returns lml. */
double oracle_kalman_filter(int kind, long n, const double* t, const double* y, const double* rvec, double noise,
                            double l, double s, double* alpha, double* mf, double* Pf, double* mp_out, double* Pp_out) {
  int d = kind_dim(kind);
  double F[DMAX * DMAX], Pinf[DMAX * DMAX], A[DMAX * DMAX], P[DMAX * DMAX], P0[DMAX * DMAX], m[DMAX] = {0, 0, 0};
  double AP[DMAX * DMAX], Pp[DMAX * DMAX], Q[DMAX * DMAX], mp[DMAX], B[DMAX];
  sde(kind, F, Pinf);
  for (int i = 0; i < DMAX * DMAX; i++) { P0[i] = s * Pinf[i]; P[i] = P0[i]; }
  double sum_logS = 0.0, sum_a2 = 0.0, tprev = n > 0 ? t[0] - 1.0 : 0.0;
  for (long k = 0; k < n; k++) {
    oracle_transition(kind, (t[k] - tprev) / l, A); tprev = t[k];
    /* Q = P0 - A P0 A^T */
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { double v = 0; for (int q = 0; q < d; q++) v += A[i * DMAX + q] * P0[q * DMAX + j]; AP[i * DMAX + j] = v; }
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { double v = 0; for (int q = 0; q < d; q++) v += AP[i * DMAX + q] * A[j * DMAX + q]; Q[i * DMAX + j] = P0[i * DMAX + j] - v; }
    /* predict */
    for (int i = 0; i < d; i++) { double v = 0; for (int q = 0; q < d; q++) v += A[i * DMAX + q] * m[q]; mp[i] = v; }
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { double v = 0; for (int q = 0; q < d; q++) v += A[i * DMAX + q] * P[q * DMAX + j]; AP[i * DMAX + j] = v; }
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { double v = 0; for (int q = 0; q < d; q++) v += AP[i * DMAX + q] * A[j * DMAX + q]; Pp[i * DMAX + j] = v + Q[i * DMAX + j]; }
    /* update (H = e_1^T) */
    double R = rvec ? rvec[k] : noise, S = Pp[0] + R, sq = sqrt(S), a = (y[k] - mp[0]) / sq;
    for (int i = 0; i < d; i++) B[i] = Pp[0 * DMAX + i] / sq;
    for (int i = 0; i < d; i++) m[i] = mp[i] + B[i] * a;
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) P[i * DMAX + j] = Pp[i * DMAX + j] - B[i] * B[j];
    sum_logS += log(S); sum_a2 += a * a;
    if (alpha) alpha[k] = a;
    if (mf) for (int i = 0; i < d; i++) { mf[k * d + i] = m[i]; mp_out[k * d + i] = mp[i]; for (int j = 0; j < d; j++) { Pf[(k * d + i) * d + j] = P[i * DMAX + j]; Pp_out[(k * d + i) * d + j] = Pp[i * DMAX + j]; } }
  }
  return -0.5 * ((double)n * LOG2PI + sum_logS + sum_a2);
}

/* RTS smoother: mean_k = m^s_k[1], var_k = P^s_k[1,1]; G^T = (P^-_{k+1} + 1e-12 I)^{-1} A_{k+1} P_k
 * through a Cholesky, as TemporalGPs `smooth`.  Returns lml. */
double oracle_kalman_smooth(int kind, long n, const double* t, const double* y, const double* rvec, double noise,
                            double l, double s, double* mean, double* var) {
  int d = kind_dim(kind);
  if (n == 0) return 0.0;
  double* mf = malloc(sizeof(double) * n * d); double* mp = malloc(sizeof(double) * n * d);
  double* Pf = malloc(sizeof(double) * n * d * d); double* Pp = malloc(sizeof(double) * n * d * d);
  double lml = oracle_kalman_filter(kind, n, t, y, rvec, noise, l, s, NULL, mf, Pf, mp, Pp);
  double ms[DMAX], Ps[DMAX * DMAX], A[DMAX * DMAX], L[DMAX * DMAX], Gt[DMAX * DMAX], W[DMAX * DMAX], T1[DMAX * DMAX];
  for (int i = 0; i < d; i++) { ms[i] = mf[(n - 1) * d + i]; for (int j = 0; j < d; j++) Ps[i * DMAX + j] = Pf[((n - 1) * d + i) * d + j]; }
  mean[n - 1] = ms[0]; var[n - 1] = Ps[0];
  for (long k = n - 2; k >= 0; k--) {
    oracle_transition(kind, (t[k + 1] - t[k]) / l, A);
    /* L L^T = Pp[k+1] + eps I */
    for (int i = 0; i < d; i++) for (int j = 0; j <= i; j++) {
      double v = Pp[((k + 1) * d + i) * d + j] + (i == j ? 1e-12 : 0.0);
      for (int q = 0; q < j; q++) v -= L[i * DMAX + q] * L[j * DMAX + q];
      L[i * DMAX + j] = (i == j) ? sqrt(v) : v / L[j * DMAX + j];
    }
    /* W = A P_k ; Gt = L^-T L^-1 W */
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { double v = 0; for (int q = 0; q < d; q++) v += A[i * DMAX + q] * Pf[(k * d + q) * d + j]; W[i * DMAX + j] = v; }
    for (int j = 0; j < d; j++) {
      for (int i = 0; i < d; i++) { double v = W[i * DMAX + j]; for (int q = 0; q < i; q++) v -= L[i * DMAX + q] * T1[q * DMAX + j]; T1[i * DMAX + j] = v / L[i * DMAX + i]; }
      for (int i = d - 1; i >= 0; i--) { double v = T1[i * DMAX + j]; for (int q = i + 1; q < d; q++) v -= L[q * DMAX + i] * Gt[q * DMAX + j]; Gt[i * DMAX + j] = v / L[i * DMAX + i]; }
    }
    /* ms_k = mf_k + Gt^T (ms_{k+1} - mp_{k+1});  Ps_k = Pf_k + Gt^T (Ps_{k+1} - Pp_{k+1}) Gt */
    double dm[DMAX], dP[DMAX * DMAX], nm[DMAX], nP[DMAX * DMAX];
    for (int i = 0; i < d; i++) { dm[i] = ms[i] - mp[(k + 1) * d + i]; for (int j = 0; j < d; j++) dP[i * DMAX + j] = Ps[i * DMAX + j] - Pp[((k + 1) * d + i) * d + j]; }
    for (int i = 0; i < d; i++) { double v = mf[k * d + i]; for (int q = 0; q < d; q++) v += Gt[q * DMAX + i] * dm[q]; nm[i] = v; }
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { double v = 0; for (int q = 0; q < d; q++) v += Gt[q * DMAX + i] * dP[q * DMAX + j]; T1[i * DMAX + j] = v; }
    for (int i = 0; i < d; i++) for (int j = 0; j < d; j++) { double v = Pf[(k * d + i) * d + j]; for (int q = 0; q < d; q++) v += T1[i * DMAX + q] * Gt[q * DMAX + j]; nP[i * DMAX + j] = v; }
    for (int i = 0; i < d; i++) { ms[i] = nm[i]; for (int j = 0; j < d; j++) Ps[i * DMAX + j] = nP[i * DMAX + j]; }
    mean[k] = ms[0]; var[k] = Ps[0];
  }
  free(mf); free(mp); free(Pf); free(Pp);
  return lml;
}

/* `batch` independent sequences (y + b*n), per-sequence parameters (l[b], s[b], noise[b]); OpenMP
 * over sequences.  alpha nullable. */
void oracle_kalman_filter_batch(int kind, long n, long batch, const double* t, const double* y, const double* rvec,
                                const double* noise, const double* l, const double* s, long nparam, double* alpha, double* lml) {
#pragma omp parallel for schedule(dynamic)
  for (long b = 0; b < batch; b++) {
    long pb = nparam == 1 ? 0 : b;
    lml[b] = oracle_kalman_filter(kind, n, t, y + b * n, rvec, noise[pb], l[pb], s[pb], alpha ? alpha + b * n : NULL, NULL, NULL, NULL, NULL);
  }
}
void oracle_kalman_smooth_batch(int kind, long n, long batch, const double* t, const double* y, const double* rvec,
                                double noise, double l, double s, double* mean, double* var, double* lml) {
#pragma omp parallel for schedule(dynamic)
  for (long b = 0; b < batch; b++) lml[b] = oracle_kalman_smooth(kind, n, t, y + b * n, rvec, noise, l, s, mean + b * n, var + b * n);
}

/* The reference's literal beta loop (src/gp/dtc.jl:108-117): one full filter per column of Cfu
 * (column-major n x m), each re-running the covariance recursion.  OpenMP over columns is the
 * "all host threads" variant; set OMP_NUM_THREADS=1 for the reference's serial behaviour. */
void oracle_decorrelate_columns(int kind, long n, long m, const double* t, const double* C, const double* rvec, double noise,
                                double l, double s, double* beta) {
#pragma omp parallel for schedule(dynamic)
  for (long c = 0; c < m; c++) oracle_kalman_filter(kind, n, t, C + c * n, rvec, noise, l, s, beta + c * n, NULL, NULL, NULL, NULL);
}
