// objectives.cuh — device objective functors: one warp evaluates one agent.
//
// N-D forms of test_functions.h:51-92 (Sphere :55, Rosenbrock :63-66, Rastrigin :74-77, Ackley :85-90) and of the
// example.cpp:41-48 Rosenbrock.  Each lane owns the coordinates j with (j / V) % 32 == lane and adds its terms in
// increasing j; the 32 lane sums are combined by an xor-butterfly.  That order is the canonical summation order the
// CPU oracle mirrors (oracle/popsolve_oracle.cpp `Lanes`), and each form is arranged so that at d = 2 it performs the
// reference's own operations in the reference's order.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace nls {

enum { OBJ_SPHERE = 0, OBJ_ROSENBROCK = 1, OBJ_RASTRIGIN = 2, OBJ_ACKLEY = 3, OBJ_ROSENBROCK_EX = 4, OBJ_COUNT = 5,
       OBJ_CUSTOM = 100 };

// User-supplied objective of an objective plugin (objective_plugin.cuh): defined only in the plugin's translation unit.
//   static constexpr bool pairwise;                      term also receives x[j-1] (terms start at j = 1)
//   static __device__ T lane0_seed(unsigned d);          constant the sum starts from (0 for a plain sum)
//   static __device__ T term(T x, T x_prev, unsigned j, unsigned d);
//   static __device__ T finish(T sum, unsigned d);
// f(x) = finish(lane0_seed + sum_j term(x[j], x[j-1], j, d), d), summed in the canonical lane order.
// Alternatively, for small fixed dimension (closed forms like the reference's 2-D test problems):
//   static constexpr unsigned full_dim = D;              1 <= D <= 8, the solver must be run with dim == D
//   static __device__ T full(const T (&x)[D]);           f(x) from the whole vector
template <class T> struct CustomObjective;
constexpr unsigned kMaxFullDim = 8;
template <class C, class = void> struct plugin_full_dim : std::integral_constant<unsigned, 0> {};
template <class C> struct plugin_full_dim<C, std::void_t<decltype(C::full_dim)>> : std::integral_constant<unsigned, C::full_dim> {};

// W = lanes that cooperate on one agent (32, or a smaller power of two when d <= W * V so that one step covers the
// row); `lane` arguments are lane indices INSIDE the group.
template <class T, int OBJ, int W = 32>
struct Objective {
  static constexpr int V = Vec<T>::V;
  static constexpr unsigned custom_full_dim() {
    if constexpr (OBJ == OBJ_CUSTOM) return plugin_full_dim<CustomObjective<T>>::value;
    else return 0;
  }
  static constexpr unsigned kFullDim = custom_full_dim();  // > 0: closed form over the whole (short) vector
  static constexpr bool custom_pairwise() {
    if constexpr (OBJ == OBJ_CUSTOM && kFullDim == 0) return CustomObjective<T>::pairwise;
    else return false;
  }
  static constexpr bool kPairwise = (OBJ == OBJ_ROSENBROCK || OBJ == OBJ_ROSENBROCK_EX) || custom_pairwise();
  typedef Ar<T> A;
  T a, b, carry;

  __device__ __forceinline__ void begin(int lane, u32 d) {
    // Rastrigin's leading `2*10` (10*d in N-D) seeds lane 0's accumulator so that d = 2 gives (20 + t0) + t1
    a = (OBJ == OBJ_RASTRIGIN && lane == 0) ? A::mul(T(10), T(d)) : T(0);
    if constexpr (OBJ == OBJ_CUSTOM && kFullDim == 0) a = lane == 0 ? CustomObjective<T>::lane0_seed(d) : T(0);
    b = T(0);
    carry = T(0);
  }

  // x[q] is coordinate j0 + q of the agent; coordinates >= d are padding and contribute nothing.
  // Must be called by all 32 lanes (the pairwise forms shuffle).
  __device__ __forceinline__ void step(const T (&x)[V], u32 j0, u32 d, int lane) {
    if constexpr (kFullDim > 0) {
      // gather the whole vector into every lane of the group (coordinate k sits in lane k / V, slot k % V); the
      // group's first lane evaluates the closed form, the others contribute 0 to the butterfly
      static_assert(kFullDim <= kMaxFullDim && kFullDim <= W * V, "full_dim must fit one step of the smallest lane group");
      T xs[kFullDim];
#pragma unroll
      for (unsigned k = 0; k < kFullDim; k++) xs[k] = __shfl_sync(kFull, x[k % V], k / V, W);
      if (lane == 0 && j0 == 0 && d == kFullDim) a = CustomObjective<T>::full(xs);
    } else {
    T left = T(0);
    if (kPairwise) {
      left = __shfl_up_sync(kFull, x[V - 1], 1, W);       // x[j0 - 1] lives in the previous lane ...
      if (lane == 0) left = carry;                        // ... or in the last lane of the previous step
      carry = __shfl_sync(kFull, x[V - 1], W - 1, W);
    }
#pragma unroll
    for (int q = 0; q < V; q++) {
      const u32 j = j0 + q;
      const T xj = x[q];
      if constexpr (OBJ == OBJ_CUSTOM) {
        const T xl = (q == 0) ? left : x[q == 0 ? 0 : q - 1];
        if ((!kPairwise || j >= 1) && j < d) a = A::add(a, CustomObjective<T>::term(xj, xl, j, d));
      } else if (kPairwise) {
        const T xl = (q == 0) ? left : x[q == 0 ? 0 : q - 1];
        if (j >= 1 && j < d) {
          if (OBJ == OBJ_ROSENBROCK) {                     // 100*pow(x0*x0 - x1, 2) + pow(x0 - 1, 2)
            const T p = A::sub(A::mul(xl, xl), xj), r = A::sub(xl, T(1));
            a = A::add(a, A::add(A::mul(T(100), A::mul(p, p)), A::mul(r, r)));
          } else {                                         // t1*t1 + 100*t2*t2, t1 = 1 - x0, t2 = x1 - x0*x0
            const T t1 = A::sub(T(1), xl), t2 = A::sub(xj, A::mul(xl, xl));
            a = A::add(a, A::add(A::mul(t1, t1), A::mul(A::mul(T(100), t2), t2)));
          }
        }
      } else if (j < d) {
        if (OBJ == OBJ_SPHERE) {
          a = A::add(a, A::mul(xj, xj));
        } else if (OBJ == OBJ_RASTRIGIN) {                 // x*x - 10*cos(2*pi*x)
          a = A::add(a, A::sub(A::mul(xj, xj), A::mul(T(10), cos2pi<T>(xj))));
        } else if (OBJ == OBJ_ACKLEY) {
          a = A::add(a, A::mul(xj, xj));
          b = A::add(b, cos2pi<T>(xj));
        }
      }
    }
    }   // separable / pairwise forms
  }

  // every lane returns the objective value
  __device__ __forceinline__ T finish(u32 d) {
    a = warp_butterfly_add<T, W>(a);
    if constexpr (OBJ == OBJ_CUSTOM && kFullDim > 0) return a;
    else if constexpr (OBJ == OBJ_CUSTOM) return CustomObjective<T>::finish(a, d);
    if (OBJ == OBJ_ACKLEY) {
      b = warp_butterfly_add<T, W>(b);
      const T inv_d = T(1.0) / T(d);
      const T ra = A::mul(T(-20), t_exp<T>(A::mul(T(-0.2), t_sqrt<T>(A::mul(inv_d, a)))));
      const T rb = -t_exp<T>(A::mul(inv_d, b));
      return A::add(A::add(A::add(ra, rb), T(2.718281828459045235360287)), T(20));
    }
    return a;
  }
};

}  // namespace nls
