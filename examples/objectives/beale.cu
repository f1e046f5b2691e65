// Objective plugin example, closed form over a short vector: the Beale function (test_functions.h:94-105),
//   f(x, y) = (1.5 - x + x y)^2 + (2.25 - x + x y^2)^2 + (2.625 - x + x y^3)^2,   minimum f(3, 0.5) = 0.
#include "objective_plugin.cuh"

template <class T>
struct Beale {
  static constexpr unsigned full_dim = 2;
  static __device__ T full(const T (&x)[2]) {
    const T a = T(1.5) - x[0] + x[0] * x[1];
    const T b = T(2.25) - x[0] + x[0] * x[1] * x[1];
    const T c = T(2.625) - x[0] + x[0] * x[1] * x[1] * x[1];
    return a * a + b * b + c * c;
  }
};
NLS_EXPORT_OBJECTIVE(Beale)
