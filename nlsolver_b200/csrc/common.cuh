// common.cuh — device primitives shared by the DE and PSO kernels (sm_100a).
//
//  * the draw tape: one splitmix64 stream per (generation tag, global agent id); random access in the draw index,
//    so the 32 lanes of a warp can each pull "their" coordinate's draw without a sequential generator
//    (replaces the shared sequential RNG& of the reference, nlsolver.h:2383-2384 / 1343-1361);
//  * u64 -> [0,1] exactly as the reference generators convert (nlsolver.h:1358-1359);
//  * contraction-free arithmetic (the reference test build has no FMA contraction);
//  * 128-bit row loads / stores that stay coherent at L2 (rows written earlier in a cooperative kernel are re-read).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nls {

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

constexpr u64 kGolden = 0x9E3779B97F4A7C15ull;
constexpr unsigned kFull = 0xffffffffu;

// splitmix64 output function (same constants as nlsolver.h:1267-1270)
__host__ __device__ __forceinline__ u64 mix64(u64 z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 tape_gen_key(u64 seed, u64 gen) { return mix64(seed + kGolden * (gen + 1)); }
__host__ __device__ __forceinline__ u64 tape_key(u64 gen_key, u64 agent) { return mix64(gen_key ^ agent); }
__host__ __device__ __forceinline__ u64 tape_state(u64 key, u64 k) { return key + kGolden * (k + 1); }
__host__ __device__ __forceinline__ u64 tape_draw(u64 key, u64 k) { return mix64(tape_state(key, k)); }

// T(u) / T(2^64 - 1): T(2^64 - 1) rounds to 2^64 in both float and double, and dividing by 2^64 is exact
template <class T> __device__ __forceinline__ T unit(u64 u);
template <> __device__ __forceinline__ double unit<double>(u64 u) { return __ull2double_rn(u) * 0x1p-64; }
template <> __device__ __forceinline__ float unit<float>(u64 u) { return __ull2float_rn(u) * 0x1p-64f; }

// contraction-free +, -, * (never fused into FMA by nvcc)
template <class T> struct Ar;
template <> struct Ar<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
};
template <> struct Ar<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
};

// generate_index (nlsolver.h:2325-2329): size_t(u * max) with the product in T; a draw of exactly 1.0 would index
// out of bounds in the reference — defined here as max - 1.
template <class T> __device__ __forceinline__ u64 index_from(u64 raw, u64 max_);
template <> __device__ __forceinline__ u64 index_from<double>(u64 raw, u64 max_) {
  const u64 v = __double2ull_rz(__dmul_rn(unit<double>(raw), __ull2double_rn(max_)));
  return v >= max_ ? max_ - 1 : v;
}
template <> __device__ __forceinline__ u64 index_from<float>(u64 raw, u64 max_) {
  const u64 v = __float2ull_rz(__fmul_rn(unit<float>(raw), __ull2float_rn(max_)));
  return v >= max_ ? max_ - 1 : v;
}

// generate_indices (nlsolver.h:2331-2355): draw until three proposals differ from `fixed` and from each other.
template <class T>
__device__ __forceinline__ void de_select_donors(u64 key, u64 P, u64 fixed, u64 &r1, u64 &r2, u64 &r3, u32 &rej) {
  u64 k = 0;
  rej = 0;
  for (;;) { r1 = index_from<T>(tape_draw(key, k++), P); if (r1 != fixed) break; rej++; }
  for (;;) { r2 = index_from<T>(tape_draw(key, k++), P); if (r2 != fixed && r2 != r1) break; rej++; }
  for (;;) { r3 = index_from<T>(tape_draw(key, k++), P); if (r3 != fixed && r3 != r1 && r3 != r2) break; rej++; }
}

// 16-byte vectors: V coordinates per lane per step
template <class T> struct Vec;
template <> struct Vec<double> { typedef double2 type; static constexpr int V = 2; };
template <> struct Vec<float> { typedef float4 type; static constexpr int V = 4; };

__device__ __forceinline__ void ld_row(const double *p, double (&x)[2]) {
  const double2 v = __ldcg(reinterpret_cast<const double2 *>(p));
  x[0] = v.x; x[1] = v.y;
}
__device__ __forceinline__ void ld_row(const float *p, float (&x)[4]) {
  const float4 v = __ldcg(reinterpret_cast<const float4 *>(p));
  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
// read-only rows shared by every agent / particle of a launch (swarm best, bounds, DE best row): keep them in L1
__device__ __forceinline__ void ld_row_shared(const double *p, double (&x)[2]) {
  const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
  x[0] = v.x; x[1] = v.y;
}
__device__ __forceinline__ void ld_row_shared(const float *p, float (&x)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
__device__ __forceinline__ void st_row(double *p, const double (&x)[2]) {
  __stcg(reinterpret_cast<double2 *>(p), make_double2(x[0], x[1]));
}
__device__ __forceinline__ void st_row(float *p, const float (&x)[4]) {
  __stcg(reinterpret_cast<float4 *>(p), make_float4(x[0], x[1], x[2], x[3]));
}

// xor-butterfly sum: every lane ends with the same value (a + b is commutative, so both partners agree bit-wise)
// W < 32: the butterfly runs inside aligned groups of W lanes.  When only the first W accumulators of the canonical
// 32 are non-zero (d <= W * V) this is bit-identical to the full butterfly, whose upper stages only add +0.
template <class T, int W = 32> __device__ __forceinline__ T warp_butterfly_add(T a) {
#pragma unroll
  for (int off = W / 2; off >= 1; off >>= 1) a = Ar<T>::add(a, __shfl_xor_sync(kFull, a, off));
  return a;
}

// cos(2*pi*x) for the Rastrigin / Ackley terms.  fp64: exact reduction r = x - rint(x) (|x| < 2^51), fold |r| > 1/4
// onto 1/2 - |r| with a sign flip, then a degree-8 polynomial in r^2 (Chebyshev fit of cos(2*pi*sqrt(t)) on
// [0, 1/16], absolute error < 2.5e-16).  This evaluates the same function as the reference's cos(2*M_PI*x)
// (test_functions.h:75, 88) without libdevice's table loads and Payne-Hanek path; the two differ by the rounding of
// 2*M_PI*x in the reference expression (<= 4e-15 per term for |x| <= 5.12), far inside the 1e-12 tolerance.
// Coefficients live in the constant bank: a DFMA can read a 64-bit constant-bank operand directly, whereas literal
// doubles are rebuilt with two UMOVs per use (17 % of the accelerated-PSO instruction stream before this change).
static __constant__ double kCos2piCoef[8] = {
    0x1.1678f9078a9b3p-2, -0x1.b6957b54dd389p+0, 0x1.f9d254582ac30p+2, -0x1.a6d1efc8c38bep+4,
    0x1.e1f506813a321p+5, -0x1.55d3c7e3bfbf5p+6, 0x1.03c1f081b5992p+6, -0x1.3bd3cc9be45dbp+4};
template <class T> __device__ __forceinline__ T cos2pi(T x);
template <> __device__ __forceinline__ double cos2pi<double>(double x) {
#ifdef NLS_LIBM_COS
  return cos(__dmul_rn(2 * 3.14159265358979323846, x));
#else
  const double magic = 6755399441055744.0;                 // 1.5 * 2^52: (x + magic) - magic == rint(x)
  const double r = x - (__dadd_rn(x, magic) - magic);      // exact, in [-1/2, 1/2]
  const bool fold = fabs(r) > 0.25;
  // (only a^2 is used, so the unfolded branch keeps r's sign: no |r| has to be materialised for the select; the final
  //  sign flip is an integer xor on the high word — the kernels that call this are bound by their FP64 operation count)
  const double a = fold ? 0.5 - fabs(r) : r;               // exact
  const double t = a * a;
  double p = kCos2piCoef[0];
#pragma unroll
  for (int k = 1; k < 8; k++) p = fma(p, t, kCos2piCoef[k]);
  p = fma(p, t, 1.0);
  return __hiloint2double(__double2hiint(p) ^ (fold ? static_cast<int>(0x80000000u) : 0), __double2loint(p));
#endif
}
template <> __device__ __forceinline__ float cos2pi<float>(float x) {
  return cosf(__fmul_rn(static_cast<float>(2 * 3.14159265358979323846), x));
}

template <class T> __device__ __forceinline__ T t_cos(T x);
template <> __device__ __forceinline__ double t_cos<double>(double x) { return cos(x); }
template <> __device__ __forceinline__ float t_cos<float>(float x) { return cosf(x); }
template <class T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <class T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }

}  // namespace nls
