// objective_plugin.cuh — build your own device objective for the DE / PSO engine (the reference
// accepts any functor as `Callable`, README.md:127-136; on the GPU the functor has to be device code).
//
// A plugin is one .cu file compiled with nvcc into a shared library; the engine loads it with
// nls_load_objective(path, &id) and `id` is then valid wherever an objective id is (nls_de_cfg.objective, ...).
//
//     #include "objective_plugin.cuh"                       // -I <repo>/nlsolver_b200/csrc -I <repo>/include
//     template <class T> struct StyblinskiTang {            // f(x) = 0.5 * sum(x^4 - 16 x^2 + 5 x)
//       static constexpr bool pairwise = false;             // true: term() also gets x[j-1], terms start at j = 1
//       static __device__ T lane0_seed(unsigned d) { return T(0); }
//       static __device__ T term(T x, T x_prev, unsigned j, unsigned d) { const T x2 = x * x; return x2 * x2 - T(16) * x2 + T(5) * x; }
//       static __device__ T finish(T sum, unsigned d) { return T(0.5) * sum; }
//     };
//     NLS_EXPORT_OBJECTIVE(StyblinskiTang)
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC -I... my_objective.cu -o libmy_objective.so
//
// Closed forms over a short vector (e.g. the reference's 2-D test problems, test_functions.h:94-318) use the other form:
//
//     template <class T> struct Beale {                     // test_functions.h:94-105
//       static constexpr unsigned full_dim = 2;             // 1..8; the solver must then be run with dim == full_dim
//       static __device__ T full(const T (&x)[2]) { ... return f; }
//     };
//
// The kernels (donor selection, crossover, repair, reductions, PSO moves, annealing chains) are the engine's own templates instantiated
// for the functor; the sum runs in the canonical lane order (objectives.cuh), so results are reproducible.
#pragma once
#define NLS_PLUGIN_BUILD 1
#include "de_impl.cuh"
#include "pso_impl.cuh"
#include "sann_impl.cuh"

#define NLS_PLUGIN_ABI 5
struct nls_objective_plugin {
  int abi;
  unsigned full_dim;   // 0: separable / pairwise sum of any dimension; D > 0: closed form, the solver's dim must be D
  // the launchers below take the state structs BY VALUE: a plugin built against other headers must not load
  unsigned de_state_bytes, pso_state_bytes, sann_state_bytes, _pad;
  const nls::DEOps *de_f64, *de_f32;
  const nls::PSOOps *pso_f64, *pso_f32;
  const nls::SANNOps *sann_f64, *sann_f32;
};

#define NLS_EXPORT_OBJECTIVE(NAME)                                                                       \
  namespace nls {                                                                                        \
  template <class T> struct CustomObjective : NAME<T> {};                                                \
  NLS_DEFINE_DE_OPS(double, plugin_de_f64)                                                               \
  NLS_DEFINE_DE_OPS(float, plugin_de_f32)                                                                \
  NLS_DEFINE_PSO_OPS(double, plugin_pso_f64)                                                             \
  NLS_DEFINE_PSO_OPS(float, plugin_pso_f32)                                                              \
  NLS_DEFINE_SANN_OPS(double, plugin_sann_f64)                                                           \
  NLS_DEFINE_SANN_OPS(float, plugin_sann_f32)                                                            \
  }                                                                                                      \
  extern "C" __attribute__((visibility("default"))) const nls_objective_plugin *nls_objective_plugin_v1() { \
    static const nls_objective_plugin p = {NLS_PLUGIN_ABI, nls::plugin_full_dim<NAME<double>>::value,    \
                                           unsigned(sizeof(nls::DEState)), unsigned(sizeof(nls::PSOState)), \
                                           unsigned(sizeof(nls::SANNState)), 0u,                           \
                                           nls::plugin_de_f64(), nls::plugin_de_f32(),                   \
                                           nls::plugin_pso_f64(), nls::plugin_pso_f32(),                 \
                                           nls::plugin_sann_f64(), nls::plugin_sann_f32()};              \
    return &p;                                                                                           \
  }
