// example_de_pso.cpp — the DE / PSO / SANN sections of the reference's example.cpp (example.cpp:159-223) and the README
// snippet (README.md:94-110), compiled against the drop-in header.  Apart from the include and the objective type
// (a device functor tag instead of a host functor) the call sites are the reference's.
//
//   g++ -std=c++17 -O2 -Iinclude examples/example_de_pso.cpp -Lnlsolver_b200 -lnls_b200
//       -Wl,-rpath,$PWD/nlsolver_b200 -o examples/example_de_pso   (one line)
#include "nlsolver_b200.hpp"

using nlsolver::DE;
using nlsolver::DESolver;
using nlsolver::PSO;
using nlsolver::SANN;
using nlsolver::rng::xorshift;
using nlsolver::rng::xoshiro;
// the reference example defines its own Rosenbrock functor (example.cpp:41-48); this is its device twin
using Rosenbrock = nlsolver::test_functions::RosenbrockExample<double>;

template <typename T>
void run_solver(T &solver, std::vector<double> init = {2, 5}) {
  auto de_res = solver.minimize(init);
  de_res.print();
  print_vector(init);
}
template <typename T>
void run_solver(T &solver, std::vector<double> lower, std::vector<double> upper, std::vector<double> init = {2, 5}) {
  auto de_res = solver.minimize(init, lower, upper);
  de_res.print();
  print_vector(init);
}

int main() {
  Rosenbrock prob;
  using DEStrat = nlsolver::RecombinationStrategy;
  std::cout << "Differential evolution with xorshift: " << std::endl;
  xorshift<double> gen;
  auto de_solver = DE<Rosenbrock, xorshift<double>, double, DEStrat::best>(prob, gen);
  run_solver(de_solver, {2, 7});

  std::cout << "README snippet (DESolver, random recombination): " << std::endl;
  gen.reset();
  auto readme_solver = DESolver<Rosenbrock, xorshift<double>, double>(prob, gen);
  std::vector<double> de_init = {5, 7};
  auto de_res = readme_solver.minimize(de_init);
  de_res.print();
  std::cout << de_init[0] << "," << de_init[1] << std::endl;

  std::cout << "Particle Swarm Optimization: " << std::endl;
  using nlsolver::PSOType;
  gen.reset();
  auto pso_solver = PSO<Rosenbrock, xorshift<double>, double>(prob, gen);
  run_solver(pso_solver, {3, 3});
  std::cout << "Particle Swarm Optimization (and bounds): " << std::endl;
  run_solver(pso_solver, {-1, -1}, {1, 1}, {0, 0});
  std::cout << "Accelerated Particle Swarm Optimization: " << std::endl;
  gen.reset();
  auto apso_solver = PSO<Rosenbrock, xorshift<double>, double, PSOType::Accelerated>(prob, gen);
  run_solver(apso_solver, {3, 3});

  std::cout << "Simulated Annealing with xoshiro: " << std::endl;   // example.cpp:216-223
  xoshiro<double> xos_gen;
  auto sann_solver = SANN<Rosenbrock, xoshiro<double>, double>(prob, xos_gen);
  run_solver(sann_solver, {5, 5});
  // what the reference cannot do: 4096 chains from the same start, the best of them is reported
  std::cout << "Simulated Annealing, 4096 chains: " << std::endl;
  std::vector<double> sann_init = {5, 5};
  auto sann_res = sann_solver.minimize_multistart(sann_init, 4096);
  sann_res.print();
  print_vector(sann_init);

  // a larger problem than the reference example can afford: Rastrigin, d = 100, 64k agents
  std::cout << "DE on Rastrigin d=100, population 65536: " << std::endl;
  nlsolver::test_functions::Rastrigin<double> rastrigin;
  gen.reset();
  auto big = DE<decltype(rastrigin), xorshift<double>, double>(rastrigin, gen, 0.9, 0.5, 1e-3, 65536, 200);
  std::vector<double> x(100, 10.24);
  auto res = big.minimize(x);
  res.print();
  std::cout << "host re-evaluation of the returned point: " << rastrigin(x) << std::endl;
  return 0;
}
