// optim_host.h — the host optimisers that gpar_group_fit runs on every device thread (pure C++, no CUDA: unit-tested on
// the CPU by tests/test_optim_host.py through tools/optim_check.cpp, like syrk_plan.h).
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <vector>

// The Nelder-Mead of gpar-at-scale_b200/neldermead.py (the restatement of Optim.jl's defaults: AffineSimplexer
// a = 0.025, b = 0.5; adaptive parameters; g_tol = 1e-8 on sqrt(var(f) n/(n+1)); the centroid is also tried at the
// end), operation for operation, so that both host layers walk the same simplices.
template <class F>
void nelder_mead(F f, const double* x0, int n, int iterations, double g_tol, double* xbest, double* fbest, int* calls_out) {
  const int m = n + 1;
  const double alpha = 1.0, beta = 1.0 + 2.0 / n, gamma = 0.75 - 1.0 / (2.0 * n), delta = 1.0 - 1.0 / n;
  std::vector<std::vector<double>> sx(m, std::vector<double>(x0, x0 + n));
  for (int i = 0; i < n; i++) sx[i + 1][i] = 1.5 * x0[i] + 0.025;
  std::vector<double> fv(m);
  for (int i = 0; i < m; i++) fv[i] = f(sx[i].data());
  int calls = m, it = 0;
  std::vector<double> cen(n), xr(n), xe(n), xc(n);
  while (it < iterations) {
    std::vector<int> order(m);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return fv[a] < fv[b]; });
    { std::vector<std::vector<double>> s2(m); std::vector<double> f2(m);
      for (int i = 0; i < m; i++) { s2[i] = sx[order[i]]; f2[i] = fv[order[i]]; }
      sx.swap(s2); fv.swap(f2); }
    double mean = 0.0; for (double v : fv) mean += v; mean /= m;
    double var = 0.0; for (double v : fv) var += (v - mean) * (v - mean); var /= (m - 1);
    if (std::sqrt(var * ((double)n / m)) <= g_tol) break;
    it++;
    for (int j = 0; j < n; j++) { double s = 0.0; for (int i = 0; i < m - 1; i++) s += sx[i][j]; cen[j] = s / (m - 1); }
    for (int j = 0; j < n; j++) xr[j] = cen[j] + alpha * (cen[j] - sx[m - 1][j]);
    const double fr = f(xr.data()); calls++;
    if (fr < fv[0]) {
      for (int j = 0; j < n; j++) xe[j] = cen[j] + beta * (xr[j] - cen[j]);
      const double fe = f(xe.data()); calls++;
      if (fe < fr) { sx[m - 1] = xe; fv[m - 1] = fe; } else { sx[m - 1] = xr; fv[m - 1] = fr; }
    } else if (fr < fv[m - 2]) {
      sx[m - 1] = xr; fv[m - 1] = fr;
    } else {
      bool ok; double fc;
      if (fr < fv[m - 1]) { for (int j = 0; j < n; j++) xc[j] = cen[j] + gamma * (xr[j] - cen[j]); fc = f(xc.data()); calls++; ok = fc <= fr; }
      else { for (int j = 0; j < n; j++) xc[j] = cen[j] - gamma * (xr[j] - cen[j]); fc = f(xc.data()); calls++; ok = fc < fv[m - 1]; }
      if (ok) { sx[m - 1] = xc; fv[m - 1] = fc; }
      else for (int i = 1; i < m; i++) { for (int j = 0; j < n; j++) sx[i][j] = sx[0][j] + delta * (sx[i][j] - sx[0][j]); fv[i] = f(sx[i].data()); calls++; }
    }
  }
  int best = 0;
  for (int i = 1; i < m; i++) if (fv[i] < fv[best]) best = i;
  for (int j = 0; j < n; j++) { double s = 0.0; for (int i = 0; i < m; i++) s += sx[i][j]; cen[j] = s / m; }
  const double fcen = f(cen.data()); calls++;
  if (fcen < fv[best]) { std::copy(cen.begin(), cen.end(), xbest); *fbest = fcen; }
  else { std::copy(sx[best].begin(), sx[best].end(), xbest); *fbest = fv[best]; }
  *calls_out = calls;
}

// The L-BFGS of gpar-at-scale_b200/lbfgs.py (two-loop recursion, memory 10, backtracking line search with a
// safeguarded quadratic step on the Armijo condition, curvature-guarded updates), on the library's analytic gradients.
template <class FG>
void lbfgs(FG fg, const double* x0, int n, int iterations, double g_tol, double f_reltol, double* xbest, double* fbest, int* calls_out) {
  typedef std::vector<double> V;
  auto dot = [&](const V& a, const V& b) { double s = 0.0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; };
  const int memory = 10;
  V x(x0, x0 + n), g(n), xn(n), gn(n), q(n), d(n);
  double f = fg(x.data(), g.data());
  int calls = 1, it = 0;
  std::vector<V> S, Y;
  if (std::isfinite(f)) {
    const double step0 = 1.0 / std::max(std::sqrt(dot(g, g)), 1.0);
    while (it < iterations) {
      double gmax = 0.0; for (double v : g) gmax = std::max(gmax, std::fabs(v));
      if (gmax <= g_tol) break;
      it++;
      q = g;
      V alphas;
      for (int k = (int)S.size() - 1; k >= 0; k--) { const double a = dot(S[k], q) / dot(Y[k], S[k]); alphas.push_back(a); for (int i = 0; i < n; i++) q[i] -= a * Y[k][i]; }
      if (!S.empty()) { const double sc = dot(S.back(), Y.back()) / dot(Y.back(), Y.back()); for (double& v : q) v *= sc; }
      for (size_t k = 0; k < S.size(); k++) { const double a = alphas[S.size() - 1 - k], b = dot(Y[k], q) / dot(Y[k], S[k]); for (int i = 0; i < n; i++) q[i] += (a - b) * S[k][i]; }
      for (int i = 0; i < n; i++) d[i] = -q[i];
      double dg = dot(d, g);
      if (dg >= 0) { S.clear(); Y.clear(); for (int i = 0; i < n; i++) d[i] = -g[i]; dg = dot(d, g); }
      double step = S.empty() ? step0 : 1.0, fn = std::numeric_limits<double>::infinity();
      bool found = false;
      for (int ls = 0; ls < 25; ls++) {
        for (int i = 0; i < n; i++) xn[i] = x[i] + step * d[i];
        fn = fg(xn.data(), gn.data()); calls++;
        if (std::isfinite(fn) && fn <= f + 1e-4 * step * dg) { found = true; break; }
        if (std::isfinite(fn)) { const double nw = -dg * step * step / (2.0 * (fn - f - dg * step)); step = std::min(std::max(nw, 0.1 * step), 0.5 * step); }
        else step *= 0.25;
      }
      if (!found) break;
      V sv(n), yv(n);
      for (int i = 0; i < n; i++) { sv[i] = xn[i] - x[i]; yv[i] = gn[i] - g[i]; }
      if (dot(sv, yv) > 1e-12 * std::sqrt(dot(sv, sv)) * std::sqrt(dot(yv, yv))) {
        S.push_back(sv); Y.push_back(yv);
        if ((int)S.size() > memory) { S.erase(S.begin()); Y.erase(Y.begin()); }
      }
      const double df = f - fn;
      x = xn; f = fn; g = gn;
      if (df <= f_reltol * std::fabs(f)) break;
    }
  }
  std::copy(x.begin(), x.end(), xbest); *fbest = f; *calls_out = calls;
}

