// example_nmpso.cpp — nlsolver::NelderMeadPSO through the drop-in header: the reference's single-solver call
// (example.cpp uses the same constructor) and the batch form this engine adds, one whole solver per start point.
//
//   g++ -std=c++17 -O2 -Iinclude examples/example_nmpso.cpp -Lnlsolver_b200 -lnls_b200
//       -Wl,-rpath,$PWD/nlsolver_b200 -o examples/example_nmpso && examples/example_nmpso
#include <cstdio>

#include "nlsolver_b200.hpp"

int main() {
  nlsolver::test_functions::Rosenbrock<double> f;
  nlsolver::rng::xorshift<double> gen;
  nlsolver::NelderMeadPSO<decltype(f), decltype(gen), double> nm(f, gen);
  std::vector<double> x = {2, 5};
  std::cout << "NelderMeadPSO, one solver:" << std::endl;
  nm.minimize(x).print();
  for (double v : x) std::printf("%.17g,", v);
  std::printf("\n");

  const size_t n = 512, d = 8;
  std::vector<std::vector<double>> starts(n, std::vector<double>(d, 1.5));
  for (size_t c = 0; c < n; c++) starts[c][c % d] += 0.001 * double(c + 1);
  auto res = nm.minimize_batch(starts);
  unsigned long long iters = 0, calls = 0;
  size_t best = 0;
  double best_f = 0;
  for (size_t c = 0; c < n; c++) {
    const auto sum = res[c].get_summary();       // (function calls, iterations, f value, ...) as in nlsolver.h:2078-2082
    calls += std::get<0>(sum);
    iters += std::get<1>(sum);
    if (c == 0 || std::get<2>(sum) < best_f) { best = c; best_f = std::get<2>(sum); }
  }
  std::printf("batch of %zu solvers: iterations %llu calls %llu best solver %zu f %.17g\n", n, iters, calls, best, best_f);
  for (double v : starts[best]) std::printf("%.17g,", v);
  std::printf("\n");
  return 0;
}
