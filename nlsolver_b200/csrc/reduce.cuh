// reduce.cuh — the population reduction both solvers end a generation with (K3 / K7 in SURVEY.md §2.1):
// a min-loc (lower value wins, lower index on ties — the outcome of the reference's sequential strict-< scans,
// nlsolver.h:2432-2437 and 2723-2729) fused with the count / mean / M2 moments that std_err (nlsolver.h:2037-2052)
// needs, block level first, then grid level: the last block to finish (ticket) combines the per-block partials in
// index order, so the result does not depend on scheduling.
#pragma once
#include <math_constants.h>

#include "common.cuh"
#include "state.h"

namespace nls {

constexpr int kBlock = 256;
constexpr int kWarpsPerBlock = kBlock / 32;

struct MinLoc { double v; u64 i; };
__device__ __forceinline__ MinLoc minloc_merge(MinLoc a, MinLoc b) {
  return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
// Chan / Golub / LeVeque pairwise combination of (n, mean, M2)
__host__ __device__ __forceinline__ Moments moments_merge(Moments a, Moments b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Moments r;
  r.n = a.n + b.n;
  const double delta = b.mean - a.mean;
  r.mean = a.mean + delta * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + delta * delta * (a.n * b.n / r.n);
  return r;
}

// block-wide combine; the result is valid in thread 0
__device__ __forceinline__ void block_reduce(MinLoc &ml, Moments &mo, MinLoc *sm_ml, Moments *sm_mo) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    MinLoc o; o.v = __shfl_down_sync(kFull, ml.v, off); o.i = __shfl_down_sync(kFull, ml.i, off);
    Moments p; p.n = __shfl_down_sync(kFull, mo.n, off); p.mean = __shfl_down_sync(kFull, mo.mean, off);
    p.m2 = __shfl_down_sync(kFull, mo.m2, off);
    ml = minloc_merge(ml, o); mo = moments_merge(mo, p);
  }
  if (lane == 0) { sm_ml[w] = ml; sm_mo[w] = mo; }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 1; k < kWarpsPerBlock; k++) { ml = minloc_merge(ml, sm_ml[k]); mo = moments_merge(mo, sm_mo[k]); }
  __syncthreads();
}

// item(i, for_min, for_moments) yields the two values of element i and must only LOAD; store(i) runs after the loads
// of a batch of kReduceBatch elements were issued (side effects of committing element i); post() runs once per thread
// after its sweep.  Batching keeps kReduceBatch independent loads in flight per thread: the sweep is a chain of dependent
// Welford updates, and with one element per iteration each of them waited for its own load.
// Returns true in every thread of the last block to finish, with the grid-wide result in (ml, mo).
constexpr int kReduceBatch = 4;
template <class PerItem, class Store, class PostSweep>
__device__ __forceinline__ bool population_reduce(u64 n, double *part_min, unsigned long long *part_idx,
                                                  Moments *part_mom, unsigned int *ticket, PerItem item, Store store,
                                                  PostSweep post, MinLoc &ml, Moments &mo) {
  __shared__ MinLoc sm_ml[kWarpsPerBlock];
  __shared__ Moments sm_mo[kWarpsPerBlock];
  __shared__ bool is_last;
  ml.v = CUDART_INF; ml.i = ~0ull;
  // The thread's own elements are accumulated as shifted sums — count, sum (x - K), sum (x - K)^2 with K the thread's
  // first element — and turned into (n, mean, M2) once: a Welford update per element is a double-precision division per
  // element, which made this sweep FP64-bound (49 us for 2^22 agents, 6 % of a d = 64 generation).  A thread sees a
  // few dozen elements of one population, so the shift keeps the cancellation in M2 = S2 - S1^2 / n harmless; across
  // threads, warps and blocks the moments are merged pairwise (Chan) as before.
  double cnt = 0.0, shift = 0.0, s1 = 0.0, s2 = 0.0;
  const u64 stride = u64(gridDim.x) * kBlock;
  for (u64 i0 = u64(blockIdx.x) * kBlock + threadIdx.x; i0 < n; i0 += stride * kReduceBatch) {
    double v[kReduceBatch], w[kReduceBatch];
#pragma unroll
    for (int u = 0; u < kReduceBatch; u++) {
      const u64 i = i0 + u * stride;
      v[u] = CUDART_INF; w[u] = 0.0;
      if (i < n) item(i, v[u], w[u]);
    }
#pragma unroll
    for (int u = 0; u < kReduceBatch; u++) {
      const u64 i = i0 + u * stride;
      if (i < n) {
        store(i);
        if (v[u] < ml.v) { ml.v = v[u]; ml.i = i; }
        if (cnt == 0.0) shift = w[u];
        const double dlt = w[u] - shift;
        cnt += 1.0;
        s1 += dlt;
        s2 = fma(dlt, dlt, s2);
      }
    }
  }
  mo.n = cnt; mo.mean = 0.0; mo.m2 = 0.0;
  if (cnt > 0.0) {
    const double m = s1 / cnt;
    mo.mean = shift + m;
    const double m2 = fma(-s1, m, s2);
    mo.m2 = m2 > 0.0 ? m2 : (m2 == m2 ? 0.0 : m2);       // rounding may leave a tiny negative; NaN stays NaN
  }
  post();   // per-thread side results (ordered before the election by the barrier inside block_reduce)
  block_reduce(ml, mo, sm_ml, sm_mo);
  if (threadIdx.x == 0) {
    part_min[blockIdx.x] = ml.v; part_idx[blockIdx.x] = ml.i; part_mom[blockIdx.x] = mo;
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  ml.v = CUDART_INF; ml.i = ~0ull; mo.n = 0.0; mo.mean = 0.0; mo.m2 = 0.0;
  for (u32 b = threadIdx.x; b < gridDim.x; b += kBlock) {
    MinLoc o; o.v = __ldcg(part_min + b); o.i = __ldcg(part_idx + b);
    Moments p; p.n = __ldcg(&part_mom[b].n); p.mean = __ldcg(&part_mom[b].mean); p.m2 = __ldcg(&part_mom[b].m2);
    ml = minloc_merge(ml, o); mo = moments_merge(mo, p);
  }
  block_reduce(ml, mo, sm_ml, sm_mo);
  if (threadIdx.x == 0) { *ticket = 0; sm_ml[0] = ml; sm_mo[0] = mo; }
  __syncthreads();
  ml = sm_ml[0]; mo = sm_mo[0];
  return true;
}

// std_err exactly as the reference computes it (nlsolver.h:2037-2052): sequential sums in index order, the square
// evaluated in double.  One thread, O(n) dependent additions — only used when the pairwise value lands within 1e-9
// (relative) of eps, where the last bits decide whether the stop rule fires; everywhere else the two agree on the
// comparison and the pairwise value is used.
template <class T>
__device__ T sequential_std_err(const T *x, u64 n) {
  T mean_val = T(0), result = T(0);
  for (u64 i = 0; i < n; i++) mean_val = Ar<T>::add(mean_val, __ldcg(x + i));
  mean_val = mean_val / static_cast<T>(n);
  for (u64 i = 0; i < n; i++) {
    const double dlt = static_cast<double>(Ar<T>::sub(__ldcg(x + i), mean_val));
    result = static_cast<T>(__dadd_rn(static_cast<double>(result), __dmul_rn(dlt, dlt)));
  }
  result = result / static_cast<T>(n - 1);
  return static_cast<T>(sqrt(static_cast<double>(result)));
}
// the same over a population stored as consecutive segments (the shards of a swarm, in rank order)
template <class T>
__device__ T sequential_std_err_segments(const void *const *seg, const unsigned long long *count, int n_seg) {
  T mean_val = T(0), result = T(0);
  u64 n = 0;
  for (int k = 0; k < n_seg; k++) {
    const T *x = static_cast<const T *>(seg[k]);
    for (u64 i = 0; i < count[k]; i++) mean_val = Ar<T>::add(mean_val, __ldcg(x + i));
    n += count[k];
  }
  mean_val = mean_val / static_cast<T>(n);
  for (int k = 0; k < n_seg; k++) {
    const T *x = static_cast<const T *>(seg[k]);
    for (u64 i = 0; i < count[k]; i++) {
      const double dlt = static_cast<double>(Ar<T>::sub(__ldcg(x + i), mean_val));
      result = static_cast<T>(__dadd_rn(static_cast<double>(result), __dmul_rn(dlt, dlt)));
    }
  }
  result = result / static_cast<T>(n - 1);
  return static_cast<T>(sqrt(static_cast<double>(result)));
}
// The pairwise value decides the stop test except where the last bits matter: within `window` (relative) of eps the
// reference's own sequential evaluation is used.  The window covers the gap between the two summation orders: a few
// ulp of T per addition level for the pairwise tree against up to n ulp for the reference's running sums — generous
// multiples of both, so a wrong decision would need the two evaluations to disagree by more than their error bounds.
template <class T>
__device__ __forceinline__ double std_err_window(double n) {
  const double ulp = sizeof(T) == 8 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
  const double w = 8.0 * n * ulp;
  return w < 1e-9 ? 1e-9 : (w > 0.25 ? 0.25 : w);
}
template <class T>
__device__ __forceinline__ T stop_std_err(const Moments &mo, const T *values, u64 n, double eps) {
  T se = static_cast<T>(sqrt(mo.m2 / (mo.n - 1.0)));
  const double e = static_cast<double>(static_cast<T>(eps));
  if (values != nullptr && e > 0.0 && fabs(static_cast<double>(se) - e) <= std_err_window<T>(mo.n) * e)
    se = sequential_std_err<T>(values, n);
  return se;
}
template <class T>
__device__ __forceinline__ T stop_std_err_segments(const Moments &mo, const void *const *seg,
                                                   const unsigned long long *count, int n_seg, double eps) {
  T se = static_cast<T>(sqrt(mo.m2 / (mo.n - 1.0)));
  const double e = static_cast<double>(static_cast<T>(eps));
  if (seg[0] != nullptr && e > 0.0 && fabs(static_cast<double>(se) - e) <= std_err_window<T>(mo.n) * e)
    se = sequential_std_err_segments<T>(seg, count, n_seg);
  return se;
}

// Exchange record (island best / sharded-swarm candidate): 48-byte header followed by one row of dim elements.
struct RecordHeader {
  double value;                 // candidate objective value (already multiplied by -1 when maximising)
  unsigned long long index;     // global agent / particle index
  Moments moments;              // moments of the shard's stop statistic (scores / particle_best_values)
  int valid, _pad;              // 0: this shard has no candidate (all values NaN / +inf)
};
static_assert(sizeof(RecordHeader) == 48, "record header layout is part of the C ABI");

}  // namespace nls
