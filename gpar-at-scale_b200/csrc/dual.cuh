// dual.cuh — forward-mode dual numbers (value + NT tangents) for the likelihood gradients through the
// Kalman scan.  The reference has no gradients (every optimiser is Nelder-Mead); the
// LGSSM kernels are templated on the scalar type and instantiated with double (values only — the
// code generated is unchanged) and with Dual<NT> (values and d/d(l, s, sigma^2) in one pass).
#pragma once
#include <cmath>

template <int NT> struct Dual {
  double v;
  double d[NT];
  __host__ __device__ __forceinline__ Dual() {}
  __host__ __device__ __forceinline__ Dual(double x) : v(x) {
#pragma unroll
    for (int i = 0; i < NT; i++) d[i] = 0.0;
  }
  __host__ __device__ __forceinline__ Dual& operator+=(const Dual& o) { v += o.v;
#pragma unroll
    for (int i = 0; i < NT; i++) d[i] += o.d[i];
    return *this; }
  __host__ __device__ __forceinline__ Dual& operator-=(const Dual& o) { v -= o.v;
#pragma unroll
    for (int i = 0; i < NT; i++) d[i] -= o.d[i];
    return *this; }
  __host__ __device__ __forceinline__ Dual& operator*=(const Dual& o) { *this = *this * o; return *this; }
  __host__ __device__ __forceinline__ Dual& operator+=(double o) { v += o; return *this; }

  friend __host__ __device__ __forceinline__ Dual operator-(const Dual& a) { Dual r; r.v = -a.v;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = -a.d[i];
    return r; }
  friend __host__ __device__ __forceinline__ Dual operator+(const Dual& a, const Dual& b) { Dual r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = a.d[i] + b.d[i];
    return r; }
  friend __host__ __device__ __forceinline__ Dual operator+(const Dual& a, double b) { Dual r = a; r.v += b; return r; }
  friend __host__ __device__ __forceinline__ Dual operator+(double a, const Dual& b) { Dual r = b; r.v += a; return r; }
  friend __host__ __device__ __forceinline__ Dual operator-(const Dual& a, const Dual& b) { Dual r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = a.d[i] - b.d[i];
    return r; }
  friend __host__ __device__ __forceinline__ Dual operator-(const Dual& a, double b) { Dual r = a; r.v -= b; return r; }
  friend __host__ __device__ __forceinline__ Dual operator-(double a, const Dual& b) { Dual r; r.v = a - b.v;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = -b.d[i];
    return r; }
  friend __host__ __device__ __forceinline__ Dual operator*(const Dual& a, const Dual& b) { Dual r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = fma(a.v, b.d[i], a.d[i] * b.v);
    return r; }
  friend __host__ __device__ __forceinline__ Dual operator*(const Dual& a, double b) { Dual r; r.v = a.v * b;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = a.d[i] * b;
    return r; }
  friend __host__ __device__ __forceinline__ Dual operator*(double a, const Dual& b) { return b * a; }
  friend __host__ __device__ __forceinline__ Dual operator/(const Dual& a, const Dual& b) {
    const double ib = 1.0 / b.v;
    Dual r; r.v = a.v * ib;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib;
    return r; }
  friend __host__ __device__ __forceinline__ Dual operator/(double a, const Dual& b) {
    const double ib = 1.0 / b.v;
    Dual r; r.v = a * ib;
    const double f = -r.v * ib;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = f * b.d[i];
    return r; }
  friend __host__ __device__ __forceinline__ Dual operator/(const Dual& a, double b) { return a * (1.0 / b); }

  // fma(a, b, c) = a b + c
  friend __host__ __device__ __forceinline__ Dual fma(const Dual& a, const Dual& b, const Dual& c) { Dual r; r.v = fma(a.v, b.v, c.v);
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = fma(a.v, b.d[i], fma(a.d[i], b.v, c.d[i]));
    return r; }
  friend __host__ __device__ __forceinline__ Dual fma(double a, const Dual& b, const Dual& c) { Dual r; r.v = fma(a, b.v, c.v);
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = fma(a, b.d[i], c.d[i]);
    return r; }
  friend __host__ __device__ __forceinline__ Dual fma(const Dual& a, double b, const Dual& c) { return fma(b, a, c); }

  friend __host__ __device__ __forceinline__ Dual exp(const Dual& a) { Dual r; r.v = ::exp(a.v);
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = r.v * a.d[i];
    return r; }
  friend __host__ __device__ __forceinline__ Dual log(const Dual& a) { Dual r; r.v = ::log(a.v);
    const double ia = 1.0 / a.v;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = a.d[i] * ia;
    return r; }
  friend __host__ __device__ __forceinline__ Dual sqrt(const Dual& a) { Dual r; r.v = ::sqrt(a.v);
    const double f = 0.5 / r.v;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = a.d[i] * f;
    return r; }
  friend __device__ __forceinline__ Dual rsqrt(const Dual& a) { Dual r; r.v = ::rsqrt(a.v);
    const double f = -0.5 * r.v / a.v;
#pragma unroll
    for (int i = 0; i < NT; i++) r.d[i] = a.d[i] * f;
    return r; }
};

// number of doubles of a scalar and component access (component 0 = value)
template <class F> struct Scalar;
template <> struct Scalar<double> {
  static constexpr int NC = 1;
  __host__ __device__ __forceinline__ static double& comp(double& x, int) { return x; }
  __host__ __device__ __forceinline__ static double comp(const double& x, int) { return x; }
};
template <int NT> struct Scalar<Dual<NT>> {
  static constexpr int NC = 1 + NT;
  __host__ __device__ __forceinline__ static double& comp(Dual<NT>& x, int c) { return c == 0 ? x.v : x.d[c > 0 ? c - 1 : 0]; }
  __host__ __device__ __forceinline__ static double comp(const Dual<NT>& x, int c) { return c == 0 ? x.v : x.d[c > 0 ? c - 1 : 0]; }
};
__host__ __device__ __forceinline__ double value_of(double x) { return x; }
template <int NT> __host__ __device__ __forceinline__ double value_of(const Dual<NT>& x) { return x.v; }
