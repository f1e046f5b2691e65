// Objective plugin example: the Styblinski–Tang function (test_functions.h, StyblinskiTang — one of the reference's
// other benchmark objectives), N-D:  f(x) = 0.5 * sum_j (x_j^4 - 16 x_j^2 + 5 x_j),  minimum at x_j = -2.903534.
#include "objective_plugin.cuh"

template <class T>
struct StyblinskiTang {
  static constexpr bool pairwise = false;
  static __device__ T lane0_seed(unsigned) { return T(0); }
  static __device__ T term(T x, T, unsigned, unsigned) {
    const T x2 = x * x;
    return x2 * x2 - T(16) * x2 + T(5) * x;
  }
  static __device__ T finish(T sum, unsigned) { return T(0.5) * sum; }
};
NLS_EXPORT_OBJECTIVE(StyblinskiTang)
