// Host harness of csrc/optim_host.h for the CPU test-suite: runs the C++ Nelder-Mead / L-BFGS twins on analytic test
// functions and prints "<minimum> <calls> <x...>" with full precision, to be compared with the Python mirror.
//   optim_check <nm|lbfgs> <rosenbrock|quadratic5|wall> <iterations> x0...
#include "../gpar-at-scale_b200/csrc/optim_host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

static double fval(const std::string& name, const double* x, int n, double* g) {
  if (name == "rosenbrock") {
    const double a = 1.0 - x[0], b = x[1] - x[0] * x[0];
    if (g) { g[0] = -2.0 * a - 400.0 * x[0] * b; g[1] = 200.0 * b; }
    return a * a + 100.0 * b * b;
  }
  if (name == "quadratic5") {
    double f = 0.0;
    for (int i = 0; i < n; i++) { const double d = x[i] - 0.1 * (i + 1); f += (i + 1) * d * d; if (g) g[i] = 2.0 * (i + 1) * d; }
    return f;
  }
  // "wall": +inf outside x0 > -0.5 (a failed Cholesky in the real objective), a bowl inside
  if (x[0] <= -0.5) { if (g) for (int i = 0; i < n; i++) g[i] = 0.0; return std::numeric_limits<double>::infinity(); }
  double f = 0.0;
  for (int i = 0; i < n; i++) { f += (x[i] + 0.4) * (x[i] + 0.4); if (g) g[i] = 2.0 * (x[i] + 0.4); }
  return f;
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const std::string alg = argv[1], fn = argv[2];
  const int it = atoi(argv[3]), n = argc - 4;
  std::vector<double> x0(n), xb(n);
  for (int i = 0; i < n; i++) x0[i] = atof(argv[4 + i]);
  double fb = 0.0; int calls = 0;
  if (alg == "nm") nelder_mead([&](const double* x) { return fval(fn, x, n, nullptr); }, x0.data(), n, it, 1e-8, xb.data(), &fb, &calls);
  else lbfgs([&](const double* x, double* g) { return fval(fn, x, n, g); }, x0.data(), n, it, 1e-6, 1e-10, xb.data(), &fb, &calls);
  printf("%.17g %d", fb, calls);
  for (int i = 0; i < n; i++) printf(" %.17g", xb[i]);
  printf("\n");
  return 0;
}
