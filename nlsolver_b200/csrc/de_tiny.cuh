// de_tiny.cuh — DE::minimize / maximize for TINY problems in one launch of one CTA, everything in shared memory.
//
// The reference's own shapes are tiny: DE defaults to a population of 50 (nlsolver.h:2391-2394), its example, README
// snippet and all fifteen problems of its test driver are 2-D (Shekel 4-D) — BASELINE.json configs[0].  A generation is
// then a few thousand instructions, and what a solve costs is latency: kernel launches, grid barriers, round trips to
// L2 for every dependent access.  For pop_size <= 1024 and dim <= 8 the whole solve (init_agents, scoring, every
// generation with the exact in-place repair, best scan, stop rules — DE::solve, nlsolver.h:2413-2476) therefore runs in
// ONE kernel of ONE block: one thread per agent, both row buffers and every per-agent array in shared memory,
// __syncthreads() where the multi-kernel path has kernel boundaries or grid barriers, and the result written straight
// into mapped host memory.  Same draw tape, same arithmetic, same summation order as the general kernels (the objective
// of a row of <= 8 coordinates touches at most four of the 32 canonical accumulators, which one thread can carry), so
// the results are bit-identical to them — and to the oracle.
#pragma once
#include <math_constants.h>

#include "common.cuh"
#include "objectives.cuh"
#include "reduce.cuh"
#include "state.h"

namespace nls {

constexpr int kTinyMaxDim = 8;
constexpr int kTinyMaxPop = 1024;

struct DETinyArgs {
  unsigned int P, d;
  unsigned long long seed, offset, cr_le, max_iter, vnc_limit;
  int cr_none, strategy;
  double F, fm, eps;
  double x0[kTinyMaxDim];
  void *result;                 // mapped host memory: DETinyResult
};
struct DETinyResult {
  double f_value, std_err;
  unsigned long long iterations, best_id, vnc, accepted, reruns, rounds;
  int stop_reason, _pad;
  double x[kTinyMaxDim];        // agents[best_id], widened
};

// f(x) of a row of d <= 8 coordinates by ONE thread, in the canonical summation order (objectives.cuh): term j goes to
// accumulator (j / V) % 32 — accumulators 0..3 at most here — and the 32 accumulators are combined by the butterfly
// 16, 8, 4, 2, 1.  The stages 16, 8, 4 only add the +0 of empty accumulators (x + 0 == x except that -0 becomes +0,
// and doing it once or three times is the same), then (a0 + a2) + (a1 + a3).
template <class T>
__device__ __forceinline__ T tiny_butterfly(const T (&a)[4]) {
  typedef Ar<T> A;
  const T z = T(0);
  const T b0 = A::add(a[0], z), b1 = A::add(a[1], z), b2 = A::add(a[2], z), b3 = A::add(a[3], z);
  return A::add(A::add(b0, b2), A::add(b1, b3));
}
template <class T, int OBJ>
__device__ __forceinline__ T tiny_objective(const T (&x)[kTinyMaxDim], u32 d) {
  typedef Ar<T> A;
  constexpr int V = Vec<T>::V;
  if constexpr (closed_form_dim(OBJ) > 0) {
    constexpr unsigned D = closed_form_dim(OBJ);
    T xs[D];
#pragma unroll
    for (unsigned k = 0; k < D; k++) xs[k] = x[k];
    const T a[4] = {closed_form<T, OBJ, D>(xs), T(0), T(0), T(0)};    // lane 0's value, the other lanes add 0
    return tiny_butterfly<T>(a);
  } else {
    T a[4] = {T(0), T(0), T(0), T(0)}, b[4] = {T(0), T(0), T(0), T(0)};
    if (OBJ == OBJ_RASTRIGIN) a[0] = A::mul(T(10), T(d));
#pragma unroll
    for (int j = 0; j < kTinyMaxDim; j++) {
      if (u32(j) >= d) break;
      const int k = j / V;
      const T xj = x[j];
      if constexpr (OBJ == OBJ_ROSENBROCK || OBJ == OBJ_ROSENBROCK_EX) {
        if (j >= 1) {
          const T xl = x[j >= 1 ? j - 1 : 0];
          if (OBJ == OBJ_ROSENBROCK) {
            const T p = A::sub(A::mul(xl, xl), xj), r = A::sub(xl, T(1));
            a[k] = A::add(a[k], A::add(A::mul(T(100), A::mul(p, p)), A::mul(r, r)));
          } else {
            const T t1 = A::sub(T(1), xl), t2 = A::sub(xj, A::mul(xl, xl));
            a[k] = A::add(a[k], A::add(A::mul(t1, t1), A::mul(A::mul(T(100), t2), t2)));
          }
        }
      } else if constexpr (OBJ == OBJ_SPHERE) {
        a[k] = A::add(a[k], A::mul(xj, xj));
      } else if constexpr (OBJ == OBJ_RASTRIGIN) {
        a[k] = A::add(a[k], A::sub(A::mul(xj, xj), A::mul(T(10), cos2pi<T>(xj))));
      } else if constexpr (OBJ == OBJ_ACKLEY) {
        a[k] = A::add(a[k], A::mul(xj, xj));
        b[k] = A::add(b[k], cos2pi<T>(xj));
      } else if constexpr (OBJ == OBJ_STYBLINSKI_TANG) {
        const T x2 = A::mul(xj, xj);
        a[k] = A::add(a[k], A::add(A::sub(A::mul(x2, x2), A::mul(T(16), x2)), A::mul(T(5), xj)));
      }
    }
    const T sa = tiny_butterfly<T>(a);
    if constexpr (OBJ == OBJ_STYBLINSKI_TANG) {
      return sa / T(2.0);
    } else if constexpr (OBJ == OBJ_ACKLEY) {
      const T sb = tiny_butterfly<T>(b);
      const T inv_d = T(1.0) / T(d);
      const T ra = A::mul(T(-20), t_exp<T>(A::mul(T(-0.2), t_sqrt<T>(A::mul(inv_d, sa)))));
      const T rb = -t_exp<T>(A::mul(inv_d, sb));
      return A::add(A::add(A::add(ra, rb), T(2.718281828459045235360287)), T(20));
    } else {
      return sa;
    }
  }
}

// block-wide min-loc + moments over one value per thread (threads >= n hold +inf / no sample); result in every thread
struct TinyReduceSmem {
  MinLoc ml[32];
  Moments mo[32];
};
__device__ __forceinline__ void tiny_reduce(MinLoc &ml, Moments &mo, TinyReduceSmem &sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    MinLoc o; o.v = __shfl_down_sync(kFull, ml.v, off); o.i = __shfl_down_sync(kFull, ml.i, off);
    Moments p; p.n = __shfl_down_sync(kFull, mo.n, off); p.mean = __shfl_down_sync(kFull, mo.mean, off);
    p.m2 = __shfl_down_sync(kFull, mo.m2, off);
    ml = minloc_merge(ml, o); mo = moments_merge(mo, p);
  }
  if (lane == 0) { sm.ml[w] = ml; sm.mo[w] = mo; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < n_warps; k++) { ml = minloc_merge(ml, sm.ml[k]); mo = moments_merge(mo, sm.mo[k]); }
    sm.ml[0] = ml; sm.mo[0] = mo;
  }
  __syncthreads();
  ml = sm.ml[0]; mo = sm.mo[0];
  __syncthreads();
}

// std_err exactly as the reference computes it (nlsolver.h:2037-2052) over a shared-memory array
template <class T>
__device__ T tiny_sequential_std_err(const T *x, u32 n) {
  T mean_val = T(0), result = T(0);
  for (u32 i = 0; i < n; i++) mean_val = Ar<T>::add(mean_val, x[i]);
  mean_val = mean_val / static_cast<T>(n);
  for (u32 i = 0; i < n; i++) {
    const double dlt = static_cast<double>(Ar<T>::sub(x[i], mean_val));
    result = static_cast<T>(__dadd_rn(static_cast<double>(result), __dmul_rn(dlt, dlt)));
  }
  result = result / static_cast<T>(n - 1);
  return static_cast<T>(sqrt(static_cast<double>(result)));
}

template <class T, int OBJ>
__global__ void __launch_bounds__(kTinyMaxPop, 1) de_tiny_solve_kernel(DETinyArgs a) {
  typedef Ar<T> A;
  extern __shared__ __align__(16) unsigned char tiny_smem[];
  const u32 P = a.P, d = a.d, i = threadIdx.x;
  const bool own = i < P;
  // shared-memory layout: rows[2][P][8] | score[P] | tscore[P] | fin[P] (u16) | where[P] | acc[P]
  // (the donors, the forced coordinate and the draw key of an agent stay in its thread's registers for the generation)
  T *rows = reinterpret_cast<T *>(tiny_smem);
  T *score = rows + size_t(2) * P * kTinyMaxDim;
  T *tscore = score + P;
  uint16_t *fin = reinterpret_cast<uint16_t *>(tscore + P);
  u8 *where = reinterpret_cast<u8 *>(fin + P);
  u8 *acc = where + P;
  __shared__ TinyReduceSmem red;
  __shared__ unsigned long long s_iter, s_best, s_vnc;
  __shared__ int s_stop;
  __shared__ unsigned int s_accepted;
  const T fm = static_cast<T>(a.fm), F = static_cast<T>(a.F);
  const bool random_mode = a.strategy != 0, cr_any = !a.cr_none;
  auto row_of = [&](u32 buf, u32 r) { return rows + (size_t(buf) * P + r) * kTinyMaxDim; };

  // ---- init_agents + initial scoring (nlsolver.h:2302-2323, 2423-2425)
  if (own) {
    const u64 key = tape_key(tape_gen_key(a.seed, 0), a.offset + i);
    T x[kTinyMaxDim];
#pragma unroll
    for (int j = 0; j < kTinyMaxDim; j++) {
      x[j] = T(0);
      if (u32(j) < d)
        x[j] = static_cast<T>(__dmul_rn(__dsub_rn(static_cast<double>(unit<T>(tape_draw(key, j))), 0.5),
                                        static_cast<double>(static_cast<T>(a.x0[j]))));
      row_of(0, i)[j] = x[j];
    }
    score[i] = A::mul(fm, tiny_objective<T, OBJ>(x, d));
    where[i] = 0; acc[i] = 0;
  }
  if (i == 0) { s_iter = 0; s_best = 0; s_vnc = 0; s_stop = 0; s_accepted = 0; }
  __syncthreads();
  unsigned long long reruns = 0, rounds = 0;     // (thread 0's copies are reported)
  int reason = 0;
  double last_se = 0.0;
  for (;;) {
    // ---- best scan, counters, stop rules (nlsolver.h:2430-2447)
    {
      MinLoc ml; ml.v = own ? static_cast<double>(score[i]) : CUDART_INF; ml.i = own ? i : ~0ull;
      Moments mo; mo.n = own ? 1.0 : 0.0; mo.mean = own ? static_cast<double>(score[i]) : 0.0; mo.m2 = 0.0;
      tiny_reduce(ml, mo, red);
      if (i == 0) {
        bool not_updated = true;
        if (ml.v < static_cast<double>(score[s_best])) { s_best = ml.i; not_updated = false; }
        s_vnc = not_updated ? s_vnc + 1 : 0;
        if (s_iter >= a.max_iter) reason = 1;
        else if (s_vnc >= a.vnc_limit) reason = 2;
        else {
          T se = static_cast<T>(sqrt(mo.m2 / (mo.n - 1.0)));
          const double e = static_cast<double>(static_cast<T>(a.eps));
          if (e > 0.0 && fabs(static_cast<double>(se) - e) <= std_err_window<T>(mo.n) * e) se = tiny_sequential_std_err<T>(score, P);
          last_se = static_cast<double>(se);
          if (se < static_cast<T>(a.eps)) reason = 3;
        }
        s_stop = reason;
      }
      __syncthreads();
      if (s_stop) break;
    }
    // ---- speculative pass: every agent against the pre-generation rows (loop body nlsolver.h:2449-2471)
    const u64 gen_key = tape_gen_key(a.seed, s_iter + 1);
    const u32 best_id = u32(s_best);
    u64 key = 0;
    u32 r1 = 0, r2 = 0, r3 = 0, dim = 0, rej = 0;
    auto evaluate = [&](bool resolved) -> bool {
      // rows of ids[0..3]; a lower donor (r < i) whose trial currently counts as accepted contributes its new row
      const u32 r0 = random_mode ? i : best_id;
      const u32 rr[4] = {r0, r1, r2, r3};
      const T *p[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        u32 w = where[rr[q]];
        if (resolved && rr[q] < i && acc[rr[q]]) w ^= 1u;
        p[q] = row_of(w, rr[q]);
      }
      const u64 sbase = tape_state(key, 4 + rej);
      T t[kTinyMaxDim];
#pragma unroll
      for (int j = 0; j < kTinyMaxDim; j++) {
        t[j] = T(0);
        if (u32(j) < d) {
          const bool mut = (cr_any && mix64(sbase + kGolden * j) <= a.cr_le) || (u32(j) == dim);
          t[j] = mut ? A::add(p[1][j], A::mul(F, A::sub(p[2][j], p[3][j]))) : p[0][j];
        }
      }
      const T sc = A::mul(fm, tiny_objective<T, OBJ>(t, d));
      const bool ok = sc < score[i];                       // strict <, NaN never accepted (nlsolver.h:2466)
      tscore[i] = sc;
      if (ok) {
        T *dst = row_of(where[i] ^ 1u, i);
#pragma unroll
        for (int j = 0; j < kTinyMaxDim; j++) dst[j] = t[j];
      }
      return ok;
    };
    if (own) {
      key = tape_key(gen_key, a.offset + i);
      u64 q1, q2, q3;
      de_select_donors<T>(key, P, random_mode ? i : best_id, q1, q2, q3, rej);
      r1 = u32(q1); r2 = u32(q2); r3 = u32(q3);
      dim = static_cast<u32>(index_from<T>(tape_draw(key, 3 + rej), d));
      const bool ok = evaluate(false);
      acc[i] = ok;
      fin[i] = ok ? uint16_t(0) : uint16_t(0xFFFFu);
    }
    const int any_ok = __syncthreads_or(own && acc[i]);
    // ---- exact in-place semantics: the fixed-point iteration of de_repair_kernel, block barriers for grid barriers
    if (any_ok) {
      for (u32 k = 1;; k++) {
        bool hit = false;
        if (own) {
          const u32 don[4] = {r1, r2, r3, best_id};
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const bool low = don[q] < i && (q < 3 || !random_mode);
            hit |= low && fin[don[q]] == uint16_t(k - 1);
          }
        }
        __syncthreads();                                   // every stamp of iteration k - 1 has been read
        bool changed = false;
        if (hit) {
          const bool old = acc[i] != 0;
          const bool now = evaluate(true);
          acc[i] = now;
          changed = old || now;
          if (changed) fin[i] = uint16_t(k);
        }
        if (i == 0) { rounds++; }
        const u32 n_hit = __syncthreads_count(hit);
        if (i == 0) reruns += n_hit;
        if (!__syncthreads_or(changed)) break;
      }
    }
    // ---- commit (nlsolver.h:2466-2471 for the accepted, :2474)
    if (own && acc[i]) { score[i] = tscore[i]; where[i] ^= 1u; atomicAdd(&s_accepted, 1u); }
    if (i == 0) s_iter += 1;
    __syncthreads();
  }
  // ---- x = agents[best_id]; solver_status(scores[best_id], iter, function_calls_used)   (nlsolver.h:2444-2446)
  if (i == 0) {
    DETinyResult *out = static_cast<DETinyResult *>(a.result);
    const u32 b = u32(s_best);
    out->f_value = static_cast<double>(score[b]);
    out->std_err = last_se;
    out->iterations = s_iter; out->best_id = b; out->vnc = s_vnc;
    out->accepted = s_accepted; out->reruns = reruns; out->rounds = rounds;
    out->stop_reason = reason; out->_pad = 0;
    const T *row = row_of(where[b], b);
    for (int j = 0; j < kTinyMaxDim; j++) out->x[j] = u32(j) < d ? static_cast<double>(row[j]) : 0.0;
    __threadfence_system();
  }
}

inline size_t de_tiny_smem_bytes(size_t P, size_t elem) {
  return (2 * P * kTinyMaxDim + 2 * P) * elem + P * (sizeof(uint16_t) + 2) + 16;
}

}  // namespace nls

namespace nls {
#ifndef NLS_PLUGIN_BUILD
template <class T, int O>
cudaError_t de_tiny_launch_o(const DETinyArgs &a, cudaStream_t st) {
  auto kernel = de_tiny_solve_kernel<T, O>;
  const size_t smem = de_tiny_smem_bytes(a.P, sizeof(T));
  static size_t allowed_of[64] = {};                       // function attributes are per DEVICE
  int dev = 0;
  cudaGetDevice(&dev);
  size_t &allowed = allowed_of[dev % 64];
  if (smem > allowed) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    allowed = smem;
  }
  const unsigned threads = (a.P + 31u) / 32u * 32u;
  kernel<<<1, threads, smem, st>>>(a);
  return cudaGetLastError();
}
template <class T>
cudaError_t de_tiny_launch(int objective, const DETinyArgs &a, cudaStream_t st) {
  switch (objective) {
#define NLS_TINY_CASE(O) case O: return de_tiny_launch_o<T, O>(a, st);
    NLS_TINY_CASE(OBJ_SPHERE) NLS_TINY_CASE(OBJ_ROSENBROCK) NLS_TINY_CASE(OBJ_RASTRIGIN) NLS_TINY_CASE(OBJ_ACKLEY)
    NLS_TINY_CASE(OBJ_ROSENBROCK_EX) NLS_TINY_CASE(OBJ_BEALE) NLS_TINY_CASE(OBJ_GOLDSTEIN_PRICE)
    NLS_TINY_CASE(OBJ_THREE_HUMP_CAMEL) NLS_TINY_CASE(OBJ_MCCORMICK) NLS_TINY_CASE(OBJ_SCHAFFER_N2)
    NLS_TINY_CASE(OBJ_STYBLINSKI_TANG) NLS_TINY_CASE(OBJ_SHEKEL) NLS_TINY_CASE(OBJ_BOOTH) NLS_TINY_CASE(OBJ_BUKIN_N6)
    NLS_TINY_CASE(OBJ_MATYAS) NLS_TINY_CASE(OBJ_LEVI_N13)
#undef NLS_TINY_CASE
    default: return cudaErrorInvalidValue;
  }
}
#endif
}  // namespace nls
