// example_multi_gpu.cpp — the reference's call sites (README.md:94-110, example.cpp:197-215) on several GPUs of one box
// from plain C++: nlsolver::b200::devices(n) is the only added line.  PSO shards ONE swarm over the devices (same result
// as on one GPU); DE runs one island of pop_size agents per device with ring migration over NVLink.
//
//   g++ -std=c++17 -O2 -Iinclude examples/example_multi_gpu.cpp -Lnlsolver_b200 -lnls_b200
//       -Wl,-rpath,$PWD/nlsolver_b200 -o examples/example_multi_gpu && examples/example_multi_gpu 8
#include <cstdlib>

#include "nlsolver_b200.hpp"

using nlsolver::DESolver;
using nlsolver::PSOSolver;
using nlsolver::rng::xorshift;
using Rastrigin = nlsolver::test_functions::Rastrigin<double>;
using Ackley = nlsolver::test_functions::Ackley<double>;

int main(int argc, char **argv) {
  const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 1;
  nlsolver::b200::devices(n_gpus, /*migrate_every=*/10, /*migrants=*/64);
  xorshift<double> gen;

  // accelerated PSO on 64-D Ackley: 2^20 particles sharded over the devices, one min-loc exchange per generation
  Ackley ackley;
  PSOSolver<Ackley, xorshift<double>, double, nlsolver::PSOType::Accelerated> pso(ackley, gen, 0.8, 1.8, 1.8, 1 << 20, 50);
  std::vector<double> x(64, 32.768);
  auto pso_res = pso.minimize(x);
  std::cout << "PSO-accelerated, Ackley d=64, 2^20 particles on " << n_gpus << " GPU(s):" << std::endl;
  pso_res.print();

  // DE on 100-D Rastrigin: one island of 2^16 agents per device
  Rastrigin rastrigin;
  DESolver<Rastrigin, xorshift<double>> de(rastrigin, gen, 0.9, 0.3, 0.0, 1 << 16, 60, 1000);
  std::vector<double> y(100, 10.24);
  auto de_res = de.minimize(y);
  std::cout << "DE islands, Rastrigin d=100, 2^16 agents per GPU on " << n_gpus << " GPU(s):" << std::endl;
  de_res.print();
  return 0;
}
