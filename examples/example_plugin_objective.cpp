// example_plugin_objective.cpp — a user objective in the Callable slot: the functor lives in an objective plugin
// (examples/objectives/styblinski_tang.cu, built with nvcc) and is loaded at run time.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -Inlsolver_b200/csrc -Iinclude
//        examples/objectives/styblinski_tang.cu -o examples/objectives/libstyblinski_tang.so
//   g++ -std=c++17 -O2 -Iinclude examples/example_plugin_objective.cpp -Lnlsolver_b200 -lnls_b200
//       -Wl,-rpath,$PWD/nlsolver_b200 -o examples/example_plugin_objective
//   ./examples/example_plugin_objective examples/objectives/libstyblinski_tang.so
#include "nlsolver_b200.hpp"

int main(int argc, char **argv) {
  if (argc < 2) { std::cerr << "usage: " << argv[0] << " <objective plugin .so>\n"; return 2; }
  auto prob = nlsolver::b200::load_objective(argv[1]);
  nlsolver::rng::xorshift<double> gen;
  nlsolver::DE<nlsolver::b200::PluginObjective, nlsolver::rng::xorshift<double>> solver(prob, gen, 0.9, 0.5, 1e-6, 4096,
                                                                                         500, 100);
  std::vector<double> x(8, 4.0);
  auto res = solver.minimize(x);
  res.print();
  print_vector(x);   // every coordinate near -2.903534, f near 8 * -39.16617
  return 0;
}
