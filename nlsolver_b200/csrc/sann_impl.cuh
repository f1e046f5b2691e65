// sann_impl.cuh — simulated annealing as a batch of independent chains.
//
// The reference's SANN (nlsolver.h:2744-2815) is one sequential chain: per candidate, d rnorm proposals
// (nlsolver.h:2479-2485), one objective call, one Metropolis test against best_val.  Chains never interact, so a batch
// of them (multi-start from one point or from one point per chain) is data-parallel over chains with the reference's
// loop running unchanged inside each: W lanes cooperate on one chain (32, or 16 / 8 / 4 when one sweep of W lanes
// covers the row), the d proposals and the objective terms of a candidate are spread over the lanes, and the accept /
// improve decisions are taken redundantly by every lane of the group from the butterfly-reduced objective value.
//
//   sann_init_kernel   : p = x = x0, best_val = f(x0)                                   (nlsolver.h:2781-2784)
//   sann_steps_kernel  : n candidates of every chain, state kept in registers between them (:2786-2813)
//   sann_gather_kernel : the chains' x (or p) rows -> one dense [C][d] array
//   sann_best_kernel   : lowest best_val over the chains, lowest chain index on ties
//
// Draw tape (DESIGN.md "SANN chains"; the CPU checker restates the same rule): stream (epoch e, global chain id); epoch e opens right
// after the chain's objective call number e and numbers its draws from 0: the Metropolis draw of the candidate just
// evaluated — made only when difference > 0, the reference's short-circuit `||` (:2804) — then the 2*d rnorm draws of
// the next candidate.
#pragma once
#include <cstdlib>

#include "objectives.cuh"
#include "pso_impl.cuh"   // rnorm_from, launch helpers


namespace nls {

// rows a group re-reads are rows the same lanes wrote one candidate earlier: ordinary cached accesses
__device__ __forceinline__ void ld_own(const double *p, double (&x)[2]) {
  const double2 v = *reinterpret_cast<const double2 *>(p);
  x[0] = v.x; x[1] = v.y;
}
__device__ __forceinline__ void ld_own(const float *p, float (&x)[4]) {
  const float4 v = *reinterpret_cast<const float4 *>(p);
  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
__device__ __forceinline__ void st_own(double *p, const double (&x)[2]) {
  *reinterpret_cast<double2 *>(p) = make_double2(x[0], x[1]);
}
__device__ __forceinline__ void st_own(float *p, const float (&x)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(x[0], x[1], x[2], x[3]);
}

// select instead of indexing the by-value kernel parameter (a dynamic index would spill the array to local memory)
__device__ __forceinline__ void *sann_buf(const SANNState &s, u32 k) { return k == 0 ? s.buf[0] : (k == 1 ? s.buf[1] : s.buf[2]); }

template <class T> __device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ double t_log<double>(double x) { return log(x); }
template <> __device__ __forceinline__ float t_log<float>(float x) { return logf(x); }

// ------------------------------------------------------------------------------------------------ init
template <class T, int OBJ>
__global__ void __launch_bounds__(kBlock) sann_init_kernel(SANNState s, const T *x0, u64 x0_count) {
  constexpr int V = Vec<T>::V;
  typedef Ar<T> A;
  const int lane = threadIdx.x & 31;
  const u64 warp = (u64(blockIdx.x) * kBlock + threadIdx.x) >> 5, n_warps = (u64(gridDim.x) * kBlock) >> 5;
  const u32 d = static_cast<u32>(s.d);
  const u32 n_sweeps = (d + 32 * V - 1) / (32 * V);
  for (u64 c = warp; c < s.C; c += n_warps) {
    const T *src = x0 + (x0_count == 1 ? 0 : c * s.d);
    T *row = static_cast<T *>(s.buf[0]) + c * s.stride;
    Objective<T, OBJ> obj;
    obj.begin(lane, d);
    for (u32 sw = 0; sw < n_sweeps; sw++) {
      const u32 j0 = (sw * 32 + lane) * V;
      T x[V];
#pragma unroll
      for (int q = 0; q < V; q++) x[q] = (j0 + q < d) ? src[j0 + q] : T(0);
      if (j0 < d) st_own(row + j0, x);
      obj.step(x, j0, d, lane);
    }
    const T val = A::mul(static_cast<T>(s.fm), obj.finish(d));
    if (lane == 0) {
      static_cast<T *>(s.best)[c] = val;
      s.role[c] = 0;
      s.n_acc[c] = 0;
      s.n_imp[c] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ candidates
// S = accumulator slots per lane (objectives.cuh): 1 when one sweep of W lanes covers the row, 32 / W when a narrow group
// sweeps a longer row.  Narrow groups pay the per-candidate work (epoch key, Metropolis test with its exp and division,
// role bookkeeping — about 230 instructions) once per WARP for 32 / W chains instead of once per chain.
#ifndef NLS_SANN_SLOT_MINBLOCKS
#define NLS_SANN_SLOT_MINBLOCKS 2   // S > 1 keeps S accumulators per lane and an S-times unrolled sweep: no spills at 128 registers
#endif
template <class T, int OBJ, int W, int S>
__global__ void __launch_bounds__(kBlock, S == 1 ? NLS_PSO_MINBLOCKS : NLS_SANN_SLOT_MINBLOCKS) sann_steps_kernel(SANNState s, u64 step_begin, u64 n_steps) {
  constexpr int V = Vec<T>::V;
  constexpr u32 kStride = W * V;
  constexpr int G = 32 / W;
  typedef Ar<T> A;
  const int lane = (threadIdx.x & 31) % W, grp = (threadIdx.x & 31) / W;
  const u64 warp = (u64(blockIdx.x) * kBlock + threadIdx.x) >> 5, n_warps = (u64(gridDim.x) * kBlock) >> 5;
  const u32 d = static_cast<u32>(s.d);
  const u32 n_sweeps = (d + kStride - 1) / kStride;
  const T fm = static_cast<T>(s.fm), scale = static_cast<T>(s.scale), tmax = static_cast<T>(s.tmax);
  for (u64 c0 = warp * G; c0 < s.C; c0 += n_warps * G) {
    const bool active = c0 + grp < s.C;                  // idle groups shadow chain c0 and store nothing
    const u64 c = active ? c0 + grp : c0;
    const u64 gc = s.offset + c;
    const u32 role = s.role[c];
    u32 pi = role & 3u, xi = (role >> 2) & 3u, shift = (role >> 4) & 1u;
    T best = static_cast<const T *>(s.best)[c];
    u32 n_acc = s.n_acc[c], n_imp = s.n_imp[c];
    // candidate number n (1-based) belongs to outer iteration (n - 1) / inner and draws from epoch n - 1; both are
    // carried along instead of being re-derived per candidate (a 64-bit division and two mix64 rounds each)
    u64 iter = step_begin / s.inner, in_iter = step_begin % s.inner;
    u64 key = tape_key(tape_gen_key(s.seed, step_begin), gc);
    for (u64 n = step_begin + 1; n <= step_begin + n_steps; n++) {
      // cooling schedule (nlsolver.h:2792-2793); e - 1 is the reference's truncated literal
      const T t = iter < s.t_n ? static_cast<T>(s.t_table[iter])
                               : tmax / t_log<T>(A::add(static_cast<T>(iter), static_cast<T>(1.7182818)));
      const T cs = A::mul(t, scale);
      const u32 fi = (pi == xi) ? (pi + 1u) % 3u : 3u - pi - xi;
      const T *prow = static_cast<const T *>(sann_buf(s, pi)) + c * s.stride;
      T *trow = static_cast<T *>(sann_buf(s, fi)) + c * s.stride;
      Objective<T, OBJ, W, S> obj;
      obj.begin(lane, d);
      u32 j0 = lane * V;
      u64 st = tape_state(key, u64(shift) + 2 * u64(j0));   // coordinate j: draws shift + 2j (log), shift + 2j + 1 (cos)
      for (u32 sw0 = 0; sw0 < n_sweeps; sw0 += S) {
#pragma unroll
      for (int slot = 0; slot < S; slot++) {
        if (sw0 + slot >= n_sweeps) break;
        const bool in = j0 < d;
        const u32 jl = in ? j0 : 0u;                     // lanes past the row end recompute coordinate 0, unused
        T x[V];
        ld_own(prow + jl, x);
#pragma unroll
        for (int q = 0; q < V; q++) {
          x[q] = A::add(x[q], A::mul(cs, rnorm_tape<T>(mix64(st + kGolden * (2 * q)), mix64(st + kGolden * (2 * q + 1)))));   // :2799
        }
        if (in && active) {
#pragma unroll
          for (int q = 0; q < V; q++)
            if (j0 + q >= d) x[q] = T(0);                // padding inside the last vector stays zero
          st_own(trow + j0, x);
        }
        obj.step(x, j0, d, lane, slot);
        j0 += kStride;
        st += kGolden * (2 * kStride);
      }
      }
      const T val = A::mul(fm, obj.finish(d));
      const T diff = A::sub(val, best);                  // against best_val, not f(p) (:2803)
      bool accept = diff <= T(0);
      shift = 0;
      key = tape_key(tape_gen_key(s.seed, n), gc);       // the objective call above opened epoch n
      if (++in_iter == s.inner) { in_iter = 0; iter++; }
      if (!accept) {                                     // Metropolis draw: first draw of the new epoch
        const T u = unit<T>(tape_draw(key, 0));
        shift = 1;
        // the reference's unqualified exp() is the double overload for float too; the compare runs in double
        accept = static_cast<double>(u) < exp(static_cast<double>(-diff / t));
      }
      if (accept) {
        pi = fi; n_acc++;
        if (val <= best) { xi = fi; best = val; n_imp++; }
      }
    }
    if (lane == 0 && active) {
      s.role[c] = static_cast<uint8_t>(pi | (xi << 2) | (shift << 4));
      static_cast<T *>(s.best)[c] = best;
      s.n_acc[c] = n_acc;
      s.n_imp[c] = n_imp;
    }
  }
}

// ------------------------------------------------------------------------------------------------ read-out
template <class T>
__global__ void __launch_bounds__(kBlock) sann_gather_kernel(SANNState s, int which, T *out) {
  const int lane = threadIdx.x & 31;
  const u64 warp = (u64(blockIdx.x) * kBlock + threadIdx.x) >> 5, n_warps = (u64(gridDim.x) * kBlock) >> 5;
  for (u64 c = warp; c < s.C; c += n_warps) {
    const u32 role = s.role[c];
    const u32 b = which == 0 ? (role >> 2) & 3u : role & 3u;
    const T *row = static_cast<const T *>(sann_buf(s, b)) + c * s.stride;
    for (u64 j = lane; j < s.d; j += 32) out[c * s.d + j] = row[j];
  }
}

template <class T>
__global__ void __launch_bounds__(kBlock) sann_best_kernel(SANNState s) {
  __shared__ MinLoc sm[kWarpsPerBlock];
  MinLoc ml; ml.v = CUDART_INF; ml.i = ~0ull;
  const T *best = static_cast<const T *>(s.best);
  for (u64 c = threadIdx.x; c < s.C; c += kBlock) {
    const double v = static_cast<double>(best[c]);
    if (v < ml.v) { ml.v = v; ml.i = c; }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    MinLoc o; o.v = __shfl_down_sync(kFull, ml.v, off); o.i = __shfl_down_sync(kFull, ml.i, off);
    ml = minloc_merge(ml, o);
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ml;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kWarpsPerBlock; k++) ml = minloc_merge(ml, sm[k]);
    s.ctrl->best_valid = ml.i != ~0ull;
    s.ctrl->best_chain = ml.i != ~0ull ? ml.i : 0;
    s.ctrl->best_buf = (s.role[ml.i != ~0ull ? ml.i : 0] >> 2) & 3;
    s.ctrl->best_value = ml.i != ~0ull ? ml.v : static_cast<double>(best[0]);
  }
}

// ------------------------------------------------------------------------------------------------ launchers
template <class T>
cudaError_t sann_launch_init(const SANNState &s, const void *x0, u64 x0_count, const LaunchGeom &g, cudaStream_t st) {
  const u64 want = (s.C + kWarpsPerBlock - 1) / kWarpsPerBlock;
#define NLS_CALL(O)                                                                                             \
  sann_init_kernel<T, O><<<pso_clamp_grid(want, u64(g.sm_count) * pso_blocks_per_sm(sann_init_kernel<T, O>)), \
                           kBlock, 0, st>>>(s, static_cast<const T *>(x0), x0_count)
  NLS_PSO_OBJ_SWITCH(s.objective, NLS_CALL)
#undef NLS_CALL
  return cudaGetLastError();
}
template <class T, int O, int W, int S>
void sann_launch_steps_w(const SANNState &s, u64 b, u64 n, const LaunchGeom &g, cudaStream_t st) {
  const u64 per_block = u64(kWarpsPerBlock) * (32 / W);
  const u64 want = (s.C + per_block - 1) / per_block;
  sann_steps_kernel<T, O, W, S><<<pso_clamp_grid(want, u64(g.sm_count) * pso_blocks_per_sm(sann_steps_kernel<T, O, W, S>)),
                                  kBlock, 0, st>>>(s, b, n);
}
// Lane-group width for a row of `row_bytes`.  Narrow groups amortise the per-candidate work over 32 / W chains per warp,
// but every resident chain keeps two hot rows (p and the candidate): 4 lanes while those fit L1 (measured on B200:
// d = 20 fp64 +65 % over one 16-lane sweep), 8 lanes while they fit L2 (d = 64: +32 %, d = 256: +4 %), a full warp per
// chain beyond that (d = 1000: 8 lanes would spill the hot rows to HBM, -8 %).  NLS_SANN_LANES=4|8|16|32 overrides.
inline int sann_lanes_for(u64 row_bytes) {
  if (const char *e = std::getenv("NLS_SANN_LANES")) {
    const int v = std::atoi(e);
    if (v == 4 || v == 8 || v == 16 || v == 32) return v;
  }
  return row_bytes <= 320 ? 4 : (row_bytes <= 2304 ? 8 : 32);
}
template <class T, int O>
void sann_launch_steps_t(const SANNState &s, u64 b, u64 n, const LaunchGeom &g, cudaStream_t st) {
  const u64 vecs = (s.d + Vec<T>::V - 1) / Vec<T>::V;
  if constexpr (Objective<T, O>::kFullDim > 0) {         // closed forms: short fixed vectors, one sweep of 4 lanes
    sann_launch_steps_w<T, O, 4, 1>(s, b, n, g, st);
    return;
  } else {
    const int single = vecs <= 4 ? 4 : (vecs <= 8 ? 8 : (vecs <= 16 ? 16 : 32));   // narrowest one-sweep group
    const int want = sann_lanes_for(s.d * sizeof(T));
    const int lanes = want < single ? want : single;
    const bool multi = lanes < single;                   // several sweeps per candidate: S = 32 / lanes slots
    if (lanes == 4) { if (multi) sann_launch_steps_w<T, O, 4, 8>(s, b, n, g, st); else sann_launch_steps_w<T, O, 4, 1>(s, b, n, g, st); }
    else if (lanes == 8) { if (multi) sann_launch_steps_w<T, O, 8, 4>(s, b, n, g, st); else sann_launch_steps_w<T, O, 8, 1>(s, b, n, g, st); }
    else if (lanes == 16) { if (multi) sann_launch_steps_w<T, O, 16, 2>(s, b, n, g, st); else sann_launch_steps_w<T, O, 16, 1>(s, b, n, g, st); }
    else sann_launch_steps_w<T, O, 32, 1>(s, b, n, g, st);
  }
}
template <class T>
cudaError_t sann_launch_steps(const SANNState &s, u64 b, u64 n, const LaunchGeom &g, cudaStream_t st) {
#define NLS_CALL(O) sann_launch_steps_t<T, O>(s, b, n, g, st)
  NLS_PSO_OBJ_SWITCH(s.objective, NLS_CALL)
#undef NLS_CALL
  return cudaGetLastError();
}
template <class T>
cudaError_t sann_launch_gather(const SANNState &s, int which, void *out, const LaunchGeom &g, cudaStream_t st) {
  const u64 want = (s.C + kWarpsPerBlock - 1) / kWarpsPerBlock;
  sann_gather_kernel<T><<<pso_clamp_grid(want, u64(g.sm_count) * 8), kBlock, 0, st>>>(s, which, static_cast<T *>(out));
  return cudaGetLastError();
}
template <class T>
cudaError_t sann_launch_best(const SANNState &s, cudaStream_t st) {
  sann_best_kernel<T><<<1, kBlock, 0, st>>>(s);
  return cudaGetLastError();
}

#define NLS_DEFINE_SANN_OPS(T, NAME)                                                                       \
  const SANNOps *NAME() {                                                                                  \
    static const SANNOps ops = {sann_launch_init<T>, sann_launch_steps<T>, sann_launch_gather<T>,          \
                                sann_launch_best<T>};                                                      \
    return &ops;                                                                                           \
  }

}  // namespace nls
