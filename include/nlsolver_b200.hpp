// nlsolver_b200.hpp — drop-in C++17 header for the DE / PSO part of JSzitas/nlsolver, backed by the B200 engine.
//
// Same names, template parameter lists, constructor defaults and method signatures as the reference
// (nlsolver.h:2377-2477 DE, 2496-2742 PSO, 2744-2815 SANN, 2054-2097 solver_status, 1263-1288 / 1343-1381 the generators), so call
// sites like example.cpp:184-215 or README.md:94-110 compile unchanged apart from the objective type: `Callable`
// must be one of the device functor tags below (test_functions.h:51-92 names), because the objective runs on the
// GPU.  Everything here is plain host C++ over the C ABI in nls_b200.h — no CUDA headers, link with -lnls_b200.
//
// What differs from the reference, by design:
//   * the population loop runs on the GPU; there is no CPU fallback (an error from the C ABI becomes
//     std::runtime_error);
//   * the user's generator is advanced by exactly TWO draws per minimize()/maximize() call — they seed the device
//     draw tape (DESIGN.md "RNG tape"); callers that share one generator across solvers or reset() it between runs
//     (test_functions.h:434-470, example.cpp:209) keep deterministic behaviour;
//   * vanilla PSO with n_particles > dim uses swarm_best_position[j] in the social term; the reference indexes it
//     with the particle number there and reads out of bounds (nlsolver.h:2674).
#ifndef NLSOLVER_B200_HPP_
#define NLSOLVER_B200_HPP_

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

#include "nls_b200.h"

// nlsolver.h:49-55
template <typename T>
void print_vector(T x) {
  for (auto &val : x) std::cout << val << ",";
  std::cout << "\n";
}

namespace nlsolver {

// ------------------------------------------------------------------------------------------------ generators
namespace rng {
namespace detail {
inline uint64_t splitmix_out(uint64_t z) {   // output function of splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
template <typename T>
inline T to_unit(uint64_t u) { return static_cast<T>(u / static_cast<T>(18446744073709551615U)); }
}  // namespace detail

// nlsolver.h:1263-1288
template <typename scalar_t = float>
struct splitmix {
  explicit splitmix() : s(12374563468ull) {}
  scalar_t yield() { return detail::to_unit<scalar_t>(yield_init()); }
  scalar_t operator()() { return yield(); }
  uint64_t yield_init() { return detail::splitmix_out(s += 0x9E3779B97F4A7C15ull); }
  void set_state(uint64_t seed) { s = seed; }
  std::vector<scalar_t> get_state() const { return {static_cast<scalar_t>(s)}; }

 private:
  uint64_t s;
};

// nlsolver.h:1343-1381 — xorshift128+ (23 / 18 / 5), default-seeded from splitmix
template <typename scalar_t = float>
struct xorshift {
  xorshift() { reset(); }
  scalar_t yield() {
    uint64_t t = x[0];
    const uint64_t s = x[1];
    x[0] = s;
    t ^= t << 23;
    t ^= t >> 18;
    t ^= s ^ (s >> 5);
    x[1] = t;
    return detail::to_unit<scalar_t>(t + s);
  }
  scalar_t operator()() { return yield(); }
  void reset() {
    splitmix<scalar_t> gn;
    x[0] = gn.yield_init();
    x[1] = x[0] >> 32;
  }
  void set_state(uint64_t y, uint64_t z) { x[0] = y; x[1] = z; }
  std::vector<scalar_t> get_state() const { return {static_cast<scalar_t>(x[0]), static_cast<scalar_t>(x[1])}; }

 private:
  uint64_t x[2]{};
};

// nlsolver.h:1289-1341 — xoshiro-style generator.  Two quirks of the reference are kept so that streams match: s[2] is
// seeded from splitmix::yield() (a value in [0,1] truncated to an integer, i.e. 0) and the rotation is by 45 of 64.
template <typename scalar_t = float>
struct xoshiro {
  xoshiro() { reset(); }
  scalar_t yield() {
    const uint64_t result = s[0] + s[3];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = (s[3] << 45) | (s[3] >> 19);
    return detail::to_unit<scalar_t>(result);
  }
  scalar_t operator()() { return yield(); }
  void reset() {
    splitmix<scalar_t> gn;
    s[0] = gn.yield_init();
    s[1] = s[0] >> 32;
    s[2] = static_cast<uint64_t>(gn.yield());
    s[3] = s[2] >> 32;
  }
  void set_state(uint64_t x, uint64_t y, uint64_t z, uint64_t t) { s[0] = x; s[1] = y; s[2] = z; s[3] = t; }
  std::vector<scalar_t> get_state() const {
    return {static_cast<scalar_t>(s[0]), static_cast<scalar_t>(s[1]), static_cast<scalar_t>(s[2]),
            static_cast<scalar_t>(s[3])};
  }

 private:
  uint64_t s[4];
};

// nlsolver.h:1179-1225 — van der Corput / Halton sequence in base b (quasi-random)
template <typename scalar_t = float>
struct halton {
  explicit halton(const scalar_t base = 2) : b(base), y(1), n(0), d(1), x(1) {}
  scalar_t yield() {
    x = d - n;
    if (x == 1) {
      n = 1;
      d *= b;
    } else {
      y = d;
      while (x <= y) {
        y /= b;
        n = (b + 1) * y - x;
      }
    }
    return static_cast<scalar_t>(n / d);
  }
  scalar_t operator()() { return yield(); }
  void reset() { b = 2; y = 1; n = 0; d = 1; x = 1; }
  std::vector<scalar_t> get_state() const { return {b, y, n, d, x}; }
  void set_state(scalar_t b_, scalar_t y_, scalar_t n_, scalar_t d_, scalar_t x_) { b = b_; y = y_; n = n_; d = d_; x = x_; }

 private:
  scalar_t b, y, n, d, x;
};

// nlsolver.h:1228-1261 — additive recurrence z <- frac(z + alpha), alpha = 0.618034
template <typename scalar_t = float>
struct recurrent {
  recurrent() : recurrent(static_cast<scalar_t>(0.5)) {}
  explicit recurrent(scalar_t seed) : alpha_(0.618034), seed_(seed), z_(alpha_ + seed_) { wrap(); }
  scalar_t yield() {
    z_ += alpha_;
    wrap();
    return z_;
  }
  scalar_t operator()() { return yield(); }
  void reset() { alpha_ = 0.618034; seed_ = 0.5; z_ = 0; }
  std::vector<scalar_t> get_state() const { return {alpha_, z_}; }
  void set_state(scalar_t alpha = 0.618034, scalar_t z = 0) { alpha_ = alpha; z_ = z; }

 private:
  void wrap() { z_ -= static_cast<scalar_t>(static_cast<uint64_t>(z_)); }
  scalar_t alpha_, seed_, z_;
};
}  // namespace rng

// ------------------------------------------------------------------------------------------------ objectives
// Device functor tags.  operator() evaluates the same N-D form on the host, in the device's summation order, so a
// caller can re-evaluate the returned point; the solvers never call it.
namespace test_functions {
namespace detail {
template <typename T, typename Term>
T lane_sum(size_t first_j, size_t d, T lane0_init, Term term) {
  constexpr size_t V = 16 / sizeof(T);
  T acc[32] = {};
  acc[0] = lane0_init;
  for (size_t j = first_j; j < d; j++) { T &a = acc[(j / V) % 32]; a = a + term(j); }
  for (int off = 16; off >= 1; off >>= 1) {
    T nxt[32];
    for (int l = 0; l < 32; l++) nxt[l] = acc[l] + acc[l ^ off];
    for (int l = 0; l < 32; l++) acc[l] = nxt[l];
  }
  return acc[0];
}
}  // namespace detail

template <typename T> struct Sphere {                      // test_functions.h:51-57
  static constexpr int nls_objective = NLS_SPHERE;
  static constexpr size_t input_size() { return 2; }
  T operator()(const std::vector<T> &x) const {
    return detail::lane_sum<T>(0, x.size(), T(0), [&](size_t j) { return x[j] * x[j]; });
  }
  std::vector<T> minimum() const { return {0.0, 0.0}; }
};
template <typename T> struct Rosenbrock {                  // test_functions.h:59-68
  static constexpr int nls_objective = NLS_ROSENBROCK;
  static constexpr size_t input_size() { return 2; }
  T operator()(const std::vector<T> &x) const {
    return detail::lane_sum<T>(1, x.size(), T(0), [&](size_t j) {
      const T p = x[j - 1] * x[j - 1] - x[j], q = x[j - 1] - 1;
      return static_cast<T>(100.0) * (p * p) + q * q;
    });
  }
  std::vector<T> minimum() const { return {1.0, 1.0}; }
};
template <typename T> struct Rastrigin {                   // test_functions.h:70-79
  static constexpr int nls_objective = NLS_RASTRIGIN;
  static constexpr size_t input_size() { return 2; }
  T operator()(const std::vector<T> &x) const {
    const T two_pi = static_cast<T>(2 * 3.14159265358979323846);
    return detail::lane_sum<T>(0, x.size(), static_cast<T>(10) * static_cast<T>(x.size()),
                               [&](size_t j) { return x[j] * x[j] - static_cast<T>(10) * std::cos(two_pi * x[j]); });
  }
  std::vector<T> minimum() const { return {0.0, 0.0}; }
};
template <typename T> struct Ackley {                      // test_functions.h:81-92
  static constexpr int nls_objective = NLS_ACKLEY;
  static constexpr size_t input_size() { return 2; }
  T operator()(const std::vector<T> &x) const {
    const T two_pi = static_cast<T>(2 * 3.14159265358979323846);
    const T sq = detail::lane_sum<T>(0, x.size(), T(0), [&](size_t j) { return x[j] * x[j]; });
    const T cs = detail::lane_sum<T>(0, x.size(), T(0), [&](size_t j) { return std::cos(two_pi * x[j]); });
    const T inv_d = static_cast<T>(1.0) / static_cast<T>(x.size());
    const T a = static_cast<T>(-20) * std::exp(static_cast<T>(-0.2) * std::sqrt(inv_d * sq));
    const T b = -std::exp(inv_d * cs);
    return a + b + static_cast<T>(std::exp(1.0)) + static_cast<T>(20);
  }
  std::vector<T> minimum() const { return {0.0, 0.0}; }
};
// ---- the other problems of the reference's test driver (test_functions.h:94-318, 485-524).  Closed forms of fixed
// dimension (2, Shekel 4) except StyblinskiTang; operator() is the host evaluation in T (pow(v,2) written as v*v).
#define NLS_B200_TAG(NAME, ID, DIM, MIN, ...)                                    \
  template <typename T> struct NAME {                                            \
    static constexpr int nls_objective = ID;                                     \
    static constexpr size_t input_size() { return DIM; }                         \
    T operator()(const std::vector<T> &x) const __VA_ARGS__                      \
    std::vector<T> minimum() const { return MIN; }                               \
  };
#define NLS_B200_MIN(...) std::vector<T>{__VA_ARGS__}
NLS_B200_TAG(Beale, NLS_BEALE, 2, NLS_B200_MIN(3.0, 0.5), {
  const T xy = x[0] * x[1], a = T(1.5) - x[0] + xy, b = T(2.25) - x[0] + xy * x[1], c = T(2.625) - x[0] + xy * x[1] * x[1];
  return a * a + b * b + c * c;
})
NLS_B200_TAG(Goldstein_Price, NLS_GOLDSTEIN_PRICE, 2, NLS_B200_MIN(0.0, -1.0), {
  const T x0 = x[0], x1 = x[1], s1 = x0 + x1 + 1, s2 = 2 * x0 - 3 * x1;
  const T a = 1 + (s1 * s1) * (19 - 14 * x0 + 3 * x0 * x0 - 14 * x1 + 6 * x0 * x1 + 3 * x1 * x1);
  const T b = 30 + (s2 * s2) * (18 - 32 * x0 + 12 * x0 * x0 + 48 * x1 - 36 * x0 * x1 + 27 * x1 * x1);
  return a * b;
})
NLS_B200_TAG(ThreeHumpCamel, NLS_THREE_HUMP_CAMEL, 2, NLS_B200_MIN(0.0, 0.0), {
  const T x2 = x[0] * x[0], x4 = x2 * x2;
  return 2 * x[0] * x[0] - T(1.05) * x4 + x4 * x2 / 6 + x[0] * x[1] + x[1] * x[1];
})
NLS_B200_TAG(McCormick, NLS_MCCORMICK, 2, NLS_B200_MIN(-0.54719, -1.54719), {
  const T dlt = x[0] - x[1];
  return std::sin(x[0] + x[1]) + dlt * dlt - T(1.5) * x[0] + T(2.5) * x[1] + 1;
})
NLS_B200_TAG(SchafferN2, NLS_SCHAFFER_N2, 2, NLS_B200_MIN(0.0, 0.0), {
  const T sn = std::sin(x[0] * x[0] - x[1] * x[1]), dn = 1 + T(0.001) * (x[0] * x[0] + x[1] * x[1]);
  return T(0.5) + (sn * sn - T(0.5)) / (dn * dn);
})
NLS_B200_TAG(StyblinskiTang, NLS_STYBLINSKI_TANG, 2, NLS_B200_MIN(-2.903534, -2.903534), {
  return detail::lane_sum<T>(0, x.size(), T(0), [&](size_t j) { const T x2 = x[j] * x[j]; return x2 * x2 - 16 * x2 + 5 * x[j]; }) / T(2.0);
})
NLS_B200_TAG(Shekel, NLS_SHEKEL, 4, NLS_B200_MIN(4.0, 4.0, 4.0, 4.0), {
  const T a[40] = {4, 4, 4, 4, 1, 1, 1, 1, 8, 8, 8, 8, 6, 6, 6, 6, 3, 7, 3, 7,
                   2, 9, 2, 9, 5, 5, 3, 3, 8, 1, 8, 1, 6, 2, 6, 2, 7, T(3.6), 7, T(3.2)};
  const T c[10] = {T(0.1), T(0.2), T(0.2), T(0.4), T(0.4), T(0.6), T(0.3), T(0.7), T(0.5), T(0.5)};
  T sum = 0;
  for (int i = 0; i < 10; i++) {
    T inner = 0;
    for (int j = 0; j < 4; j++) { const T dlt = x[j] - a[i * 4 + j]; inner += dlt * dlt; }
    sum += T(1.0) / (inner + c[i]);
  }
  return -sum;
})
NLS_B200_TAG(Booth, NLS_BOOTH, 2, NLS_B200_MIN(1.0, 3.0), {
  const T a = x[0] + 2 * x[1] - 7, b = 2 * x[0] + x[1] - 5;
  return a * a + b * b;
})
NLS_B200_TAG(BukinN6, NLS_BUKIN_N6, 2, NLS_B200_MIN(-10.0, 1.0), {
  return 100 * std::sqrt(std::abs(x[1] - T(0.01) * x[0] * x[0])) + T(0.01) * std::abs(x[0] + 10);
})
NLS_B200_TAG(Matyas, NLS_MATYAS, 2, NLS_B200_MIN(0.0, 0.0), {
  return T(0.26) * (x[0] * x[0] + x[1] * x[1]) - T(0.48) * x[0] * x[1];
})
NLS_B200_TAG(LeviN13, NLS_LEVI_N13, 2, NLS_B200_MIN(1.0, 1.0), {
  const T pi3 = T(3 * 3.14159265358979323846), pi2 = T(2 * 3.14159265358979323846);
  const T s0 = std::sin(pi3 * x[0]), s1 = std::sin(pi3 * x[1]), s2 = std::sin(pi2 * x[1]), a = x[0] - 1, b = x[1] - 1;
  return s0 * s0 + (a * a) * (1 + s1 * s1) + (b * b) * (1 + s2 * s2);
})
#undef NLS_B200_TAG
#undef NLS_B200_MIN

// the Rosenbrock variant of example.cpp:41-48 and README.md:83-90
template <typename T> struct RosenbrockExample {
  static constexpr int nls_objective = NLS_ROSENBROCK_EX;
  static constexpr size_t input_size() { return 2; }
  T operator()(const std::vector<T> &x) const {
    return detail::lane_sum<T>(1, x.size(), T(0), [&](size_t j) {
      const T t1 = 1 - x[j - 1], t2 = x[j] - x[j - 1] * x[j - 1];
      return t1 * t1 + static_cast<T>(100) * t2 * t2;
    });
  }
  std::vector<T> minimum() const { return {1.0, 1.0}; }
};
}  // namespace test_functions

namespace b200 {
template <typename C, typename = void> struct objective_of : std::integral_constant<int, -1> {};
template <typename C>
struct objective_of<C, std::void_t<decltype(C::nls_objective)>> : std::integral_constant<int, C::nls_objective> {};
template <typename T> constexpr int dtype_of() {
  static_assert(std::is_same<T, double>::value || std::is_same<T, float>::value, "scalar_t must be float or double");
  return std::is_same<T, double>::value ? NLS_F64 : NLS_F32;
}
inline void check(int rc) {
  if (rc != NLS_OK) throw std::runtime_error(std::string("nls_b200: ") + nls_last_error());
}
// A user objective built as an objective plugin (nlsolver_b200/csrc/objective_plugin.cuh) fills the Callable slot
// through this type:   auto prob = nlsolver::b200::load_objective("libmy_objective.so");
//                      nlsolver::DE<nlsolver::b200::PluginObjective, xorshift<double>> solver(prob, gen);
struct PluginObjective { int id; };
inline PluginObjective load_objective(const std::string &plugin_path) {
  int32_t id = -1;
  check(nls_load_objective(plugin_path.c_str(), &id));
  return PluginObjective{id};
}
template <typename C> constexpr bool is_objective() {
  return objective_of<C>::value >= 0 || std::is_same<C, PluginObjective>::value;
}
template <typename C> int objective_id(const C &f) {
  if constexpr (std::is_same<C, PluginObjective>::value) return f.id;
  else return objective_of<C>::value;
}
// one context per process, created on first use on device $NLS_B200_DEVICE (default 0)
inline nls_ctx *default_context() {
  struct Holder {
    nls_ctx *ctx = nullptr;
    Holder() {
      const char *dev = std::getenv("NLS_B200_DEVICE");
      check(nls_ctx_create(dev ? std::atoi(dev) : 0, nullptr, &ctx));
    }
    ~Holder() { nls_ctx_destroy(ctx); }
  };
  static Holder h;
  return h.ctx;
}
// ---- several GPUs from one process -------------------------------------------------------------------------------
// nlsolver::b200::devices(n) makes every solve that follows use n GPUs (devices 0 .. n - 1) of the box:
//   PSO  — ONE swarm of n_particles sharded over the devices; the result is identical to the single-GPU solve;
//   DE   — n islands of pop_size agents each (a DE population does not shard: every agent gathers three random rows),
//          the `migrants` best rows moving around the ring every `migrate_every` generations; x receives the best
//          island's best agent.  With n = 1 (the default) a solve is exactly the reference's single population.
struct group_options {
  int n_devices = 1;
  uint64_t migrate_every = 10, migrants = 64;
};
inline group_options &options() {
  static group_options o;
  return o;
}
inline void devices(int n, uint64_t migrate_every = 10, uint64_t migrants = 64) {
  options().n_devices = n < 1 ? 1 : n;
  options().migrate_every = migrate_every;
  options().migrants = migrants;
}
inline nls_group *default_group() {
  struct Holder {
    nls_group *g = nullptr;
    int n = 0;
    ~Holder() { nls_group_destroy(g); }
  };
  static Holder h;
  if (h.n != options().n_devices) {
    nls_group_destroy(h.g);
    h.g = nullptr;
    h.n = 0;
    check(nls_group_create(options().n_devices, nullptr, &h.g));
    h.n = options().n_devices;
  }
  return h.g;
}
// two draws of the user's generator -> 64-bit tape seed (hi word first)
template <typename RNG>
uint64_t seed_from(RNG &generator) {
  auto word = [&]() {
    const double v = static_cast<double>(generator()) * 4294967296.0;
    return v >= 4294967295.0 ? 0xFFFFFFFFull : (v <= 0.0 ? 0ull : static_cast<uint64_t>(v));
  };
  const uint64_t hi = word();
  const uint64_t lo = word();
  return (hi << 32) | lo;
}
}  // namespace b200

// ------------------------------------------------------------------------------------------------ solver_status
// nlsolver.h:2054-2097
template <typename scalar_t = double>
struct solver_status {
  solver_status(const scalar_t f_val, const size_t iter_used, const size_t f_calls_used,
                const size_t grad_evals_used = 0ul, const size_t hess_evals_used = 0ul)
      : f_value(f_val), iteration(iter_used), function_calls_used(f_calls_used),
        gradient_evals_used(grad_evals_used), hessian_evals_used(hess_evals_used) {}
  void print() const {
    std::cout << "Function calls used: " << function_calls_used << std::endl;
    std::cout << "Algorithm iterations used: " << iteration << std::endl;
    if (gradient_evals_used > 0) std::cout << "Gradient evaluations used: " << gradient_evals_used << std::endl;
    if (hessian_evals_used > 0) std::cout << "Hessian evaluations used: " << hessian_evals_used << std::endl;
    std::cout << "With final function value of " << f_value << std::endl;
  }
  std::tuple<size_t, size_t, scalar_t, size_t, size_t> get_summary() const {
    return std::make_tuple(function_calls_used, iteration, f_value, gradient_evals_used, hessian_evals_used);
  }
  void add(const solver_status<scalar_t> &more) {
    const auto o = more.get_summary();
    function_calls_used += std::get<0>(o);
    iteration += std::get<1>(o);
    f_value = std::get<2>(o);
    gradient_evals_used += std::get<3>(o);
    hessian_evals_used += std::get<4>(o);
  }

 private:
  scalar_t f_value;
  size_t iteration, function_calls_used, gradient_evals_used, hessian_evals_used;
};

// ------------------------------------------------------------------------------------------------ DE
enum RecombinationStrategy { best, random };   // nlsolver.h:2377

// nlsolver.h:2379-2477
template <typename Callable, typename RNG, typename scalar_t = double,
          RecombinationStrategy RecombinationType = random>
class DE {
  static_assert(b200::is_objective<Callable>(),
                "nlsolver_b200: Callable must be a device objective tag (nlsolver::test_functions::Sphere, "
                "Rosenbrock, Rastrigin, Ackley, RosenbrockExample) or a loaded nlsolver::b200::PluginObjective; "
                "host functors cannot run on the GPU");

 public:
  DE(Callable &f, RNG &generator, const scalar_t crossover_prob = 0.9, const scalar_t differential_weight = 0.8,
     const scalar_t eps = 10e-4, const size_t pop_size = 50, const size_t max_iter = 1000,
     const size_t best_val_no_change = 50)
      : f(f), generator(generator), crossover_prob(crossover_prob), differential_weight(differential_weight),
        eps(eps), pop_size(pop_size), max_iter(max_iter), best_value_no_change(best_val_no_change) {}
  solver_status<scalar_t> minimize(std::vector<scalar_t> &x) { return solve(x, true); }
  solver_status<scalar_t> maximize(std::vector<scalar_t> &x) { return solve(x, false); }

 private:
  solver_status<scalar_t> solve(std::vector<scalar_t> &x, bool minimize) {
    nls_de_cfg cfg{};
    cfg.dtype = b200::dtype_of<scalar_t>();
    cfg.objective = b200::objective_id(f);
    cfg.strategy = RecombinationType == random ? NLS_DE_RANDOM : NLS_DE_BEST;
    cfg.minimize = minimize ? 1 : 0;
    cfg.pop_size = pop_size;
    cfg.dim = x.size();
    cfg.crossover_prob = crossover_prob;
    cfg.differential_weight = differential_weight;
    cfg.eps = eps;
    cfg.max_iter = max_iter;
    cfg.best_val_no_change = best_value_no_change;
    cfg.seed = b200::seed_from(generator);
    std::vector<scalar_t> best_row(x.size());
    nls_status st{};
    if (b200::options().n_devices > 1)
      b200::check(nls_de_solve_islands(b200::default_group(), &cfg, x.data(), b200::options().migrate_every,
                                       b200::options().migrants, best_row.data(), &st));
    else
      b200::check(nls_de_solve(b200::default_context(), &cfg, x.data(), best_row.data(), &st));
    x = best_row;                                              // nlsolver.h:2444
    return solver_status<scalar_t>(static_cast<scalar_t>(st.f_value), st.iterations, st.function_calls);
  }
  Callable &f;
  RNG &generator;
  const scalar_t crossover_prob, differential_weight, eps;
  const size_t pop_size, max_iter, best_value_no_change;
};

// ------------------------------------------------------------------------------------------------ PSO
enum PSOType { Vanilla, Accelerated };   // nlsolver.h:2496

// nlsolver.h:2498-2742
template <typename Callable, typename RNG, typename scalar_t = double, PSOType Type = Vanilla>
class PSO {
  static_assert(b200::is_objective<Callable>(),
                "nlsolver_b200: Callable must be a device objective tag (see nlsolver::test_functions) or a loaded "
                "nlsolver::b200::PluginObjective");

 public:
  PSO(Callable &f, RNG &generator, const scalar_t inertia = 0.8, const scalar_t cognitive_coef = 1.8,
      const scalar_t social_coef = 1.8, const size_t n_particles = 10, const size_t max_iter = 5000,
      const size_t best_val_no_change = 50, const scalar_t eps = 10e-4)
      : generator(generator), f(f), inertia(inertia), cognitive_coef(cognitive_coef), social_coef(social_coef),
        n_particles(n_particles), max_iter(max_iter), best_val_no_change(best_val_no_change), eps(eps) {}
  // unbounded: lower = -|x|, upper = |x| seed the swarm, positions are never clamped (nlsolver.h:2553-2575)
  solver_status<scalar_t> minimize(std::vector<scalar_t> &x) { return unbounded(x, true); }
  solver_status<scalar_t> maximize(std::vector<scalar_t> &x) { return unbounded(x, false); }
  // bounded: positions are clamped to [lower, upper] every generation (nlsolver.h:2577-2589)
  solver_status<scalar_t> minimize(std::vector<scalar_t> &x, const std::vector<scalar_t> &lower,
                                   const std::vector<scalar_t> &upper) { return solve(x, lower, upper, true, true); }
  solver_status<scalar_t> maximize(std::vector<scalar_t> &x, const std::vector<scalar_t> &lower,
                                   const std::vector<scalar_t> &upper) { return solve(x, lower, upper, false, true); }

 private:
  solver_status<scalar_t> unbounded(std::vector<scalar_t> &x, bool minimize) {
    std::vector<scalar_t> lower(x.size()), upper(x.size());
    for (size_t i = 0; i < x.size(); i++) { upper[i] = std::abs(x[i]); lower[i] = -upper[i]; }
    return solve(x, lower, upper, minimize, false);
  }
  solver_status<scalar_t> solve(std::vector<scalar_t> &x, const std::vector<scalar_t> &lower,
                                const std::vector<scalar_t> &upper, bool minimize, bool constrained) {
    if (lower.size() != upper.size()) throw std::invalid_argument("nlsolver_b200: lower / upper sizes differ");
    nls_pso_cfg cfg{};
    cfg.dtype = b200::dtype_of<scalar_t>();
    cfg.objective = b200::objective_id(f);
    cfg.pso_type = Type == Vanilla ? NLS_PSO_VANILLA : NLS_PSO_ACCELERATED;
    cfg.minimize = minimize ? 1 : 0;
    cfg.n_particles = n_particles;
    cfg.dim = lower.size();
    cfg.inertia = inertia;
    cfg.cognitive_coef = cognitive_coef;
    cfg.social_coef = social_coef;
    cfg.eps = eps;
    cfg.max_iter = max_iter;
    cfg.best_val_no_change = best_val_no_change;
    cfg.constrained = constrained ? 1 : 0;
    cfg.flags = (Type == Vanilla && n_particles > lower.size()) ? NLS_FLAG_SOCIAL_INDEX_J : 0u;
    cfg.seed = b200::seed_from(generator);
    std::vector<scalar_t> best_row(lower.size());
    nls_status st{};
    if (b200::options().n_devices > 1 && n_particles >= size_t(b200::options().n_devices))
      b200::check(nls_pso_solve_sharded(b200::default_group(), &cfg, lower.data(), upper.data(), best_row.data(), &st));
    else
      b200::check(nls_pso_solve(b200::default_context(), &cfg, lower.data(), upper.data(), best_row.data(), &st));
    if (st.best_valid) x = best_row;
    else x.clear();   // the reference assigns a never-filled swarm_best_position (nlsolver.h:2601)
    return solver_status<scalar_t>(static_cast<scalar_t>(st.f_value), st.iterations, st.function_calls);
  }
  RNG &generator;
  Callable &f;
  const scalar_t inertia, cognitive_coef, social_coef;
  const size_t n_particles, max_iter, best_val_no_change;
  const scalar_t eps;
};

// ------------------------------------------------------------------------------------------------ SANN
// nlsolver.h:2744-2815.  minimize(x) / maximize(x) run ONE chain from x, as the reference does.  The batch calls are
// what this engine adds (the reference has no batching concept): one independent chain per start point — or n_chains
// chains from one point — each the reference's loop on its own draw stream, all resident on the GPU.
template <typename Callable, typename RNG, typename scalar_t = double>
class SANN {
  static_assert(b200::is_objective<Callable>(),
                "nlsolver_b200: Callable must be a device objective tag (see nlsolver::test_functions) or a loaded "
                "nlsolver::b200::PluginObjective");

 public:
  SANN(Callable &f, RNG &generator, const size_t max_iter = 5000, const size_t temperature_iter = 10,
       const scalar_t temperature_max = 10.0)
      : generator(generator), f(f), f_evals(0), max_iter(max_iter), temperature_iter(temperature_iter),
        temperature_max(temperature_max) {}
  solver_status<scalar_t> minimize(std::vector<scalar_t> &x) { return solve(x, true); }
  solver_status<scalar_t> maximize(std::vector<scalar_t> &x) { return solve(x, false); }
  // one chain per row of xs; every row is overwritten with its chain's best point
  std::vector<solver_status<scalar_t>> minimize_batch(std::vector<std::vector<scalar_t>> &xs) { return batch(xs, true); }
  std::vector<solver_status<scalar_t>> maximize_batch(std::vector<std::vector<scalar_t>> &xs) { return batch(xs, false); }
  // n_chains chains from the same start x; x receives the best chain's point, the return value its status
  solver_status<scalar_t> minimize_multistart(std::vector<scalar_t> &x, const size_t n_chains) { return multistart(x, n_chains, true); }
  solver_status<scalar_t> maximize_multistart(std::vector<scalar_t> &x, const size_t n_chains) { return multistart(x, n_chains, false); }

 private:
  nls_sann_cfg config(size_t n_chains, size_t dim, bool minimize) {
    nls_sann_cfg cfg{};
    cfg.dtype = b200::dtype_of<scalar_t>();
    cfg.objective = b200::objective_id(f);
    cfg.minimize = minimize ? 1 : 0;
    cfg.n_chains = n_chains;
    cfg.dim = dim;
    cfg.max_iter = max_iter;
    cfg.temperature_iter = temperature_iter;
    cfg.temperature_max = temperature_max;
    cfg.seed = b200::seed_from(generator);
    return cfg;
  }
  solver_status<scalar_t> multistart(std::vector<scalar_t> &x, size_t n_chains, bool minimize) {
    const nls_sann_cfg cfg = config(n_chains, x.size(), minimize);
    std::vector<scalar_t> best_row(x.size());
    nls_status st{};
    b200::check(nls_sann_solve(b200::default_context(), &cfg, x.data(), 1, best_row.data(), &st));
    x = best_row;
    f_evals += st.function_calls;                              // a member that is never reset (nlsolver.h:2751, 2783)
    return solver_status<scalar_t>(static_cast<scalar_t>(st.f_value), st.iterations, f_evals);
  }
  solver_status<scalar_t> solve(std::vector<scalar_t> &x, bool minimize) { return multistart(x, 1, minimize); }
  std::vector<solver_status<scalar_t>> batch(std::vector<std::vector<scalar_t>> &xs, bool minimize) {
    std::vector<solver_status<scalar_t>> out;
    if (xs.empty()) return out;
    const size_t n = xs.size(), d = xs[0].size();
    std::vector<scalar_t> flat(n * d), fbest(n);
    for (size_t c = 0; c < n; c++) {
      if (xs[c].size() != d) throw std::invalid_argument("nlsolver_b200: start points of different sizes");
      for (size_t j = 0; j < d; j++) flat[c * d + j] = xs[c][j];
    }
    const nls_sann_cfg cfg = config(n, d, minimize);
    nls_sann *h = nullptr;
    b200::check(nls_sann_create(b200::default_context(), &cfg, flat.data(), n, &h));
    nls_status st{};
    int rc = nls_sann_step(h, ~0ull);
    if (rc == NLS_OK) rc = nls_sann_sync(h, &st);
    if (rc == NLS_OK) rc = nls_sann_read_chains(h, flat.data(), fbest.data(), nullptr, nullptr, nullptr);
    nls_sann_destroy(h);
    b200::check(rc);
    f_evals += st.function_calls;
    for (size_t c = 0; c < n; c++) {
      for (size_t j = 0; j < d; j++) xs[c][j] = flat[c * d + j];
      out.emplace_back(fbest[c], st.iterations, st.function_calls / n);
    }
    return out;
  }
  RNG &generator;
  Callable &f;
  size_t f_evals;
  const size_t max_iter, temperature_iter;
  const scalar_t temperature_max;
};

// ------------------------------------------------------------------------------------------------ NelderMeadPSO
// nlsolver.h:3546-3920.  minimize(x) / maximize(x) run ONE solver from x, as the reference does; the batch calls are
// what this engine adds (one independent solver per start point, all resident on the GPU, one warp each).  The
// reference's bounded overloads clamp with lower[i] / upper[i] indexed by the particle loop counter (nlsolver.h:3859),
// which is out of bounds for every particle; they are declared here and throw.
template <typename Callable, typename RNG, typename scalar_t = double>
class NelderMeadPSO {
  static_assert(b200::is_objective<Callable>(),
                "nlsolver_b200: Callable must be a device objective tag (see nlsolver::test_functions)");

 public:
  NelderMeadPSO(Callable &f, RNG &generator, const scalar_t alpha = 1, const scalar_t gamma = 2,
                const scalar_t rho = 0.5, const scalar_t sigma = 0.5, const scalar_t inertia = 0.8,
                const scalar_t cognitive_coef = 1.8, const scalar_t social_coef = 1.8, const scalar_t eps = 1e-6,
                const size_t max_iter = 1000, const size_t no_change_best_iter = 20)
      : generator(generator), f(f), alpha(alpha), gamma(gamma), rho(rho), sigma(sigma), inertia(inertia),
        cognitive_coef(cognitive_coef), social_coef(social_coef), eps(eps), max_iter(max_iter),
        no_change_best_iter(no_change_best_iter) {}
  solver_status<scalar_t> minimize(std::vector<scalar_t> &x) { return solve(x, true); }
  solver_status<scalar_t> maximize(std::vector<scalar_t> &x) { return solve(x, false); }
  solver_status<scalar_t> minimize(std::vector<scalar_t> &, const std::vector<scalar_t> &, const std::vector<scalar_t> &) {
    throw std::invalid_argument("nlsolver_b200: the bounded NelderMeadPSO overloads read lower[i] / upper[i] out of "
                                "bounds in the reference (nlsolver.h:3859) and have no defined result");
  }
  solver_status<scalar_t> maximize(std::vector<scalar_t> &x, const std::vector<scalar_t> &l, const std::vector<scalar_t> &u) {
    return minimize(x, l, u);
  }
  // one solver per row of xs; every row is overwritten with its solver's best point
  std::vector<solver_status<scalar_t>> minimize_batch(std::vector<std::vector<scalar_t>> &xs) { return batch(xs, true); }
  std::vector<solver_status<scalar_t>> maximize_batch(std::vector<std::vector<scalar_t>> &xs) { return batch(xs, false); }

 private:
  nls_nmpso_cfg config(size_t n_solvers, size_t dim, bool minimize) {
    nls_nmpso_cfg cfg{};
    cfg.dtype = b200::dtype_of<scalar_t>();
    cfg.objective = b200::objective_id(f);
    cfg.minimize = minimize ? 1 : 0;
    cfg.n_solvers = n_solvers;
    cfg.dim = dim;
    cfg.alpha = alpha; cfg.gamma = gamma; cfg.rho = rho; cfg.sigma = sigma; cfg.inertia = inertia;
    cfg.cognitive_coef = cognitive_coef; cfg.social_coef = social_coef; cfg.eps = eps;
    cfg.max_iter = max_iter; cfg.no_change_best_iter = no_change_best_iter;
    cfg.seed = b200::seed_from(generator);
    return cfg;
  }
  solver_status<scalar_t> solve(std::vector<scalar_t> &x, bool minimize) {
    if (x.size() < 2) {   // nlsolver.h:3619-3629
      std::cout << "You are trying to optimize a one dimensional function "
                << "you should probably be using vanilla NelderMead (or vanilla PSO)"
                << " - our implementation does not support this in the "
                   "NelderMead-PSO hybrid."
                << std::endl;
      return solver_status<scalar_t>(999999, 0, 0);
    }
    std::vector<std::vector<scalar_t>> one{x};
    auto res = batch(one, minimize);
    x = one[0];
    return res[0];
  }
  std::vector<solver_status<scalar_t>> batch(std::vector<std::vector<scalar_t>> &xs, bool minimize) {
    std::vector<solver_status<scalar_t>> out;
    if (xs.empty()) return out;
    const size_t n = xs.size(), d = xs[0].size();
    std::vector<scalar_t> flat(n * d), fbest(n);
    std::vector<uint64_t> iters(n), calls(n);
    for (size_t c = 0; c < n; c++) {
      if (xs[c].size() != d) throw std::invalid_argument("nlsolver_b200: start points of different sizes");
      for (size_t j = 0; j < d; j++) flat[c * d + j] = xs[c][j];
    }
    const nls_nmpso_cfg cfg = config(n, d, minimize);
    nls_status st{};
    b200::check(nls_nmpso_solve(b200::default_context(), &cfg, flat.data(), n, flat.data(), fbest.data(), iters.data(),
                                calls.data(), &st));
    for (size_t c = 0; c < n; c++) {
      for (size_t j = 0; j < d; j++) xs[c][j] = flat[c * d + j];
      out.emplace_back(fbest[c], iters[c], calls[c]);
    }
    return out;
  }
  RNG &generator;
  Callable &f;
  const scalar_t alpha, gamma, rho, sigma, inertia, cognitive_coef, social_coef, eps;
  const size_t max_iter, no_change_best_iter;
};

// the names README.md:80,99 uses
template <typename Callable, typename RNG, typename scalar_t = double,
          RecombinationStrategy RecombinationType = random>
using DESolver = DE<Callable, RNG, scalar_t, RecombinationType>;
template <typename Callable, typename RNG, typename scalar_t = double, PSOType Type = Vanilla>
using PSOSolver = PSO<Callable, RNG, scalar_t, Type>;

}  // namespace nlsolver
#endif  // NLSOLVER_B200_HPP_
