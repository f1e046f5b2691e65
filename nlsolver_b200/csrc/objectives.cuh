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

enum { OBJ_SPHERE = 0, OBJ_ROSENBROCK = 1, OBJ_RASTRIGIN = 2, OBJ_ACKLEY = 3, OBJ_ROSENBROCK_EX = 4,
       // the other problems the reference's test driver runs DE / PSO on (test_functions.h:94-318, 485-524)
       OBJ_BEALE = 5, OBJ_GOLDSTEIN_PRICE = 6, OBJ_THREE_HUMP_CAMEL = 7, OBJ_MCCORMICK = 8, OBJ_SCHAFFER_N2 = 9,
       OBJ_STYBLINSKI_TANG = 10, OBJ_SHEKEL = 11, OBJ_BOOTH = 12, OBJ_BUKIN_N6 = 13, OBJ_MATYAS = 14, OBJ_LEVI_N13 = 15,
       OBJ_COUNT = 16, OBJ_CUSTOM = 100 };

// Closed forms over a short vector: the reference's fixed-dimension test problems.  Dimension of objective `obj`
// (0: not a closed form — a sum of any dimension).
__host__ __device__ constexpr unsigned closed_form_dim(int obj) {
  return obj == OBJ_SHEKEL ? 4u
         : (obj == OBJ_BEALE || obj == OBJ_GOLDSTEIN_PRICE || obj == OBJ_THREE_HUMP_CAMEL || obj == OBJ_MCCORMICK ||
            obj == OBJ_SCHAFFER_N2 || obj == OBJ_BOOTH || obj == OBJ_BUKIN_N6 || obj == OBJ_MATYAS || obj == OBJ_LEVI_N13)
               ? 2u : 0u;
}

template <class T> __device__ __forceinline__ T t_sin(T x);
template <> __device__ __forceinline__ double t_sin<double>(double x) { return sin(x); }
template <> __device__ __forceinline__ float t_sin<float>(float x) { return sinf(x); }

// f(x) for the closed forms, contraction-free; pow(v, 2) of the reference is v*v, pow(v, 4) = (v*v)*(v*v),
// pow(v, 6) = ((v*v)*(v*v))*(v*v) (the oracle restates them the same way).
template <class T, int OBJ, unsigned D>
__device__ __forceinline__ T closed_form(const T (&x)[D]) {
  typedef Ar<T> A;
  auto sq = [](T v) { return Ar<T>::mul(v, v); };
  if constexpr (OBJ == OBJ_BEALE) {                     // test_functions.h:98-102
    const T xy = A::mul(x[0], x[1]), xyy = A::mul(xy, x[1]), xyyy = A::mul(xyy, x[1]);
    return A::add(A::add(sq(A::add(A::sub(T(1.5), x[0]), xy)), sq(A::add(A::sub(T(2.25), x[0]), xyy))),
                  sq(A::add(A::sub(T(2.625), x[0]), xyyy)));
  } else if constexpr (OBJ == OBJ_GOLDSTEIN_PRICE) {    // :109-117
    const T x0 = x[0], x1 = x[1];
    T p = A::sub(T(19), A::mul(T(14), x0));
    p = A::add(p, A::mul(A::mul(T(3), x0), x0));
    p = A::sub(p, A::mul(T(14), x1));
    p = A::add(p, A::mul(A::mul(T(6), x0), x1));
    p = A::add(p, A::mul(A::mul(T(3), x1), x1));
    const T a = A::add(T(1), A::mul(sq(A::add(A::add(x0, x1), T(1))), p));
    T q = A::sub(T(18), A::mul(T(32), x0));
    q = A::add(q, A::mul(A::mul(T(12), x0), x0));
    q = A::add(q, A::mul(T(48), x1));
    q = A::sub(q, A::mul(A::mul(T(36), x0), x1));
    q = A::add(q, A::mul(A::mul(T(27), x1), x1));
    const T b = A::add(T(30), A::mul(sq(A::sub(A::mul(T(2), x0), A::mul(T(3), x1))), q));
    return A::mul(a, b);
  } else if constexpr (OBJ == OBJ_THREE_HUMP_CAMEL) {   // :146-148
    const T x2 = sq(x[0]), x4 = A::mul(x2, x2), x6 = A::mul(x4, x2);
    T r = A::sub(A::mul(A::mul(T(2), x[0]), x[0]), A::mul(T(1.05), x4));
    r = A::add(r, x6 / T(6));
    r = A::add(r, A::mul(x[0], x[1]));
    return A::add(r, sq(x[1]));
  } else if constexpr (OBJ == OBJ_MCCORMICK) {          // :209-211
    T r = A::add(t_sin<T>(A::add(x[0], x[1])), sq(A::sub(x[0], x[1])));
    r = A::sub(r, A::mul(T(1.5), x[0]));
    r = A::add(r, A::mul(T(2.5), x[1]));
    return A::add(r, T(1));
  } else if constexpr (OBJ == OBJ_SCHAFFER_N2) {        // :219-221
    const T s2 = sq(t_sin<T>(A::sub(sq(x[0]), sq(x[1]))));
    const T den = sq(A::add(T(1), A::mul(T(0.001), A::add(sq(x[0]), sq(x[1])))));
    return A::add(T(0.5), A::sub(s2, T(0.5)) / den);
  } else if constexpr (OBJ == OBJ_SHEKEL) {             // :258-276, 4-D, ten wells
    const T a[40] = {4, 4, 4, 4, 1, 1, 1, 1, 8, 8, 8, 8, 6, 6, 6, 6, 3, 7, 3, 7,
                     2, 9, 2, 9, 5, 5, 3, 3, 8, 1, 8, 1, 6, 2, 6, 2, 7, T(3.6), 7, T(3.2)};
    const T c[10] = {T(0.1), T(0.2), T(0.2), T(0.4), T(0.4), T(0.6), T(0.3), T(0.7), T(0.5), T(0.5)};
    T sum = T(0);
#pragma unroll
    for (int i = 0; i < 10; i++) {
      T inner = T(0);
#pragma unroll
      for (int j = 0; j < 4; j++) inner = A::add(inner, sq(A::sub(x[j], a[i * 4 + j])));
      sum = A::add(sum, T(1.0) / A::add(inner, c[i]));
    }
    return -sum;
  } else if constexpr (OBJ == OBJ_BOOTH) {              // :283-285
    return A::add(sq(A::sub(A::add(x[0], A::mul(T(2), x[1])), T(7))), sq(A::sub(A::add(A::mul(T(2), x[0]), x[1]), T(5))));
  } else if constexpr (OBJ == OBJ_BUKIN_N6) {           // :292-295
    const T r = t_sqrt<T>(fabs(A::sub(x[1], A::mul(A::mul(T(0.01), x[0]), x[0]))));
    return A::add(A::mul(T(100), r), A::mul(T(0.01), fabs(A::add(x[0], T(10)))));
  } else if constexpr (OBJ == OBJ_MATYAS) {             // :302-304
    return A::sub(A::mul(T(0.26), A::add(sq(x[0]), sq(x[1]))), A::mul(A::mul(T(0.48), x[0]), x[1]));
  } else if constexpr (OBJ == OBJ_LEVI_N13) {           // :311-317
    const T pi3 = T(3 * 3.14159265358979323846), pi2 = T(2 * 3.14159265358979323846);
    const T s0 = sq(t_sin<T>(A::mul(pi3, x[0])));
    const T t1 = A::mul(sq(A::sub(x[0], T(1))), A::add(T(1), sq(t_sin<T>(A::mul(pi3, x[1])))));
    const T t2 = A::mul(sq(A::sub(x[1], T(1))), A::add(T(1), sq(t_sin<T>(A::mul(pi2, x[1])))));
    return A::add(A::add(s0, t1), t2);
  } else {
    return T(0);
  }
}

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
// S = accumulator slots per lane.  The canonical order has 32 accumulators, term j going to accumulator (j / V) % 32.
// A group of W < 32 lanes that sweeps a LONGER row (d > W * V) keeps S = 32 / W of them per lane: sweep k of the row
// feeds slot k % S, i.e. canonical accumulator slot * W + lane, and finish() runs the upper butterfly stages (offsets
// 16 ... W) between slots before the shuffles — the same pairs are added in the same stage order, so the result is
// bit-identical to the 32-lane evaluation.  S = 1 is the plain case (W = 32, or one sweep).
template <class T, int OBJ, int W = 32, int S = 1>
struct Objective {
  static constexpr int V = Vec<T>::V;
  static constexpr unsigned custom_full_dim() {
    if constexpr (OBJ == OBJ_CUSTOM) return plugin_full_dim<CustomObjective<T>>::value;
    else return closed_form_dim(OBJ);
  }
  static constexpr unsigned kFullDim = custom_full_dim();  // > 0: closed form over the whole (short) vector
  static constexpr bool custom_pairwise() {
    if constexpr (OBJ == OBJ_CUSTOM && kFullDim == 0) return CustomObjective<T>::pairwise;
    else return false;
  }
  static constexpr bool kPairwise = (OBJ == OBJ_ROSENBROCK || OBJ == OBJ_ROSENBROCK_EX) || custom_pairwise();
  typedef Ar<T> A;
  static_assert(S == 1 || S * W == 32, "S slots of W lanes must tile the 32 canonical accumulators");
  T acc_a[S], acc_b[S], carry;

  __device__ __forceinline__ void begin(int lane, u32 d) {
#pragma unroll
    for (int k = 0; k < S; k++) { acc_a[k] = T(0); acc_b[k] = T(0); }
    // Rastrigin's leading `2*10` (10*d in N-D) seeds lane 0's accumulator so that d = 2 gives (20 + t0) + t1
    acc_a[0] = (OBJ == OBJ_RASTRIGIN && lane == 0) ? A::mul(T(10), T(d)) : T(0);
    if constexpr (OBJ == OBJ_CUSTOM && kFullDim == 0) acc_a[0] = lane == 0 ? CustomObjective<T>::lane0_seed(d) : T(0);
    carry = T(0);
  }

  // x[q] is coordinate j0 + q of the agent; coordinates >= d are padding and contribute nothing.
  // Must be called by all 32 lanes (the pairwise forms shuffle).  `slot` = sweep number % S (a constant after unrolling).
  __device__ __forceinline__ void step(const T (&x)[V], u32 j0, u32 d, int lane, int slot = 0) {
    T &a = acc_a[slot];
    T &b = acc_b[slot];
    if constexpr (kFullDim > 0) {
      // gather the whole vector into every lane of the group (coordinate k sits in lane k / V, slot k % V); the
      // group's first lane evaluates the closed form, the others contribute 0 to the butterfly
      static_assert(kFullDim <= kMaxFullDim && kFullDim <= W * V, "full_dim must fit one step of the smallest lane group");
      T xs[kFullDim];
#pragma unroll
      for (unsigned k = 0; k < kFullDim; k++) xs[k] = __shfl_sync(kFull, x[k % V], k / V, W);
      if (lane == 0 && j0 == 0 && d == kFullDim) {
        if constexpr (OBJ == OBJ_CUSTOM) a = CustomObjective<T>::full(xs);
        else a = closed_form<T, OBJ, kFullDim>(xs);
      }
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
      } else if (OBJ == OBJ_ACKLEY) {
        // branch-free (a padding coordinate is evaluated and dropped): a jump around the cos polynomial would end the
        // basic block, and the terms of the lane's V coordinates would be evaluated one dependent chain after the other
        const T sq = A::mul(xj, xj), c = cos2pi<T>(xj);
        const bool valid = j < d;
        a = valid ? A::add(a, sq) : a;
        b = valid ? A::add(b, c) : b;
      } else if (j < d) {
        if (OBJ == OBJ_SPHERE) {
          a = A::add(a, A::mul(xj, xj));
        } else if (OBJ == OBJ_RASTRIGIN) {                 // x*x - 10*cos(2*pi*x)
          a = A::add(a, A::sub(A::mul(xj, xj), A::mul(T(10), cos2pi<T>(xj))));
        } else if (OBJ == OBJ_STYBLINSKI_TANG) {         // pow(x,4) - 16*pow(x,2) + 5*x, test_functions.h:246-252
          const T x2 = A::mul(xj, xj);
          a = A::add(a, A::add(A::sub(A::mul(x2, x2), A::mul(T(16), x2)), A::mul(T(5), xj)));
        }
      }
    }
    }   // separable / pairwise forms
  }

  // every lane returns the objective value
  __device__ __forceinline__ T finish(u32 d) {
    // butterfly stages with offsets >= W pair accumulators of the same lane: slot ^ (offset / W)
#pragma unroll
    for (int off = S / 2; off >= 1; off >>= 1) {
      T ta[S], tb[S];
#pragma unroll
      for (int k = 0; k < S; k++) { ta[k] = A::add(acc_a[k], acc_a[k ^ off]); tb[k] = A::add(acc_b[k], acc_b[k ^ off]); }
#pragma unroll
      for (int k = 0; k < S; k++) { acc_a[k] = ta[k]; acc_b[k] = tb[k]; }
    }
    T a = warp_butterfly_add<T, W>(acc_a[0]);
    T b = acc_b[0];
    if constexpr (kFullDim > 0) {
      return a;                                            // closed forms: lane 0's value, the others added 0
    } else if constexpr (OBJ == OBJ_CUSTOM) {
      return CustomObjective<T>::finish(a, d);
    } else if constexpr (OBJ == OBJ_STYBLINSKI_TANG) {
      return a / T(2.0);
    } else if constexpr (OBJ == OBJ_ACKLEY) {
      b = warp_butterfly_add<T, W>(b);
      const T inv_d = T(1.0) / T(d);
      const T ra = A::mul(T(-20), t_exp<T>(A::mul(T(-0.2), t_sqrt<T>(A::mul(inv_d, a)))));
      const T rb = -t_exp<T>(A::mul(inv_d, b));
      return A::add(A::add(A::add(ra, rb), T(2.718281828459045235360287)), T(20));
    } else {
      return a;
    }
  }
};

}  // namespace nls
