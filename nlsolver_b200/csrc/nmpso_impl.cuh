// nmpso_impl.cuh — nlsolver::NelderMeadPSO (nlsolver.h:3546-3920) as a batch of independent solvers, one warp each.
//
// The hybrid keeps 3n + 1 particles; per iteration it sorts them by value (:3654), applies one Nelder–Mead step to the
// best n + 1 (apply_simplex, :3731-3822) and a PSO move to the other 2n (apply_pso, :3824-3868) — a sequential
// algorithm with a population tied to the dimension, so the only data parallelism worth having is ACROSS solvers
// (multi-start: one solver per start point).  A warp owns a solver: the lanes spread over the coordinates (row
// arithmetic, the objective's lane-strided sum) or over the particles (the rank sort), every decision is taken by all
// lanes from the same values, and the particles' rows live in global memory (L1 / L2 resident: the warp re-reads what
// it just wrote).  The accidents of the reference are kept (see the restatement in oracle/popsolve_oracle.cpp, pinned
// to the unmodified reference): velocities never change (`velocity` is a by-value copy, :3843-3845), the pair's
// reference particle is current_order[i + 1] from the second pair on (:3836-3838), best_val is never updated (:3650),
// vertex v of the initial simplex raises coordinate v (vertex n is x itself; the reference's store to [n][n] is out of
// bounds), only the unbounded overloads exist (the bounded ones index lower / upper with the particle counter, :3859).
// Sorting: the reference's std::sort is unstable; here ties are broken by the particle id, and parity cases are tie-free.
//
// Draw tape: stream (tag, global solver id); tag 0: init_solver_state — PSO particle p, coordinate j: position draw
// 2 (p n + j), velocity draw 2 (p n + j) + 1 (:3712-3721); tag it + 1: apply_pso of iteration it — q-th PSO particle in
// sorted order, coordinate j: r_p = draw 2 (q n + j), r_g = the next (:3847).
#pragma once
#include <math_constants.h>

#include "common.cuh"
#include "objectives.cuh"

namespace nls {

struct NMPSOState {
  void *pos, *vel;        // [C][N][stride] particle_positions / particle_velocities, N = 3 n + 1
  void *tmp;              // [C][4][stride] centroid, reflected, expanded, contracted point
  void *x0;               // [x0_count][d]
  void *x_best, *f_best;  // [C][d], [C]
  unsigned long long *iters, *evals;   // [C]
  unsigned long long C, d, stride, x0_count;
  unsigned long long seed, offset, max_iter, no_change_limit;
  double alpha, gamma, rho, sigma, inertia, cog, soc, eps, fm;
  int objective, _pad;
};

constexpr int kNMPSOWarps = 4;              // solvers per block
constexpr unsigned kNMPSOMaxDim = 256;      // 3 n + 1 = 769 values per solver in shared memory

template <class T> __device__ __forceinline__ T nm_abs(T v) { return v < T(0) ? -v : v; }

// f(row) by the warp, canonical order (objectives.cuh)
template <class T, int OBJ>
__device__ __forceinline__ T nm_eval(const T *row, u32 d, int lane) {
  constexpr int V = Vec<T>::V;
  Objective<T, OBJ> obj;
  obj.begin(lane, d);
  const u32 n_steps = (d + 32 * V - 1) / (32 * V);
  for (u32 st = 0; st < n_steps; st++) {
    const u32 j0 = (st * 32 + lane) * V;
    T x[V];
#pragma unroll
    for (int q = 0; q < V; q++) x[q] = (j0 + q < d) ? row[j0 + q] : T(0);
    obj.step(x, j0, d, lane);
  }
  return obj.finish(d);
}

template <class T, int OBJ>
__global__ void __launch_bounds__(kNMPSOWarps * 32) nmpso_solve_kernel(NMPSOState s) {
  typedef Ar<T> A;
  extern __shared__ __align__(16) unsigned char nm_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const u64 c = u64(blockIdx.x) * kNMPSOWarps + wib;
  if (c >= s.C) return;                                    // (whole warps: no block-wide barrier is used below)
  const u32 n = static_cast<u32>(s.d), ns = n + 1, N = 3 * n + 1;
  const u64 stride = s.stride;
  T *val = reinterpret_cast<T *>(nm_smem) + size_t(wib) * N;
  u32 *order = reinterpret_cast<u32 *>(nm_smem + size_t(kNMPSOWarps) * N * sizeof(T)) + size_t(wib) * N;
  u32 *scratch = order + size_t(kNMPSOWarps) * N;         // the sort's output before it replaces `order`
  T *pos = static_cast<T *>(s.pos) + c * N * stride;
  T *vel = static_cast<T *>(s.vel) + c * N * stride;
  T *cen = static_cast<T *>(s.tmp) + c * 4 * stride, *t_ref = cen + stride, *t_exp = cen + 2 * stride, *t_con = cen + 3 * stride;
  const T *x0 = static_cast<const T *>(s.x0) + (s.x0_count == 1 ? 0 : c * s.d);
  const T fm = static_cast<T>(s.fm), alpha = static_cast<T>(s.alpha), gamma = static_cast<T>(s.gamma),
          rho = static_cast<T>(s.rho), sigma = static_cast<T>(s.sigma), inertia = static_cast<T>(s.inertia),
          cog = static_cast<T>(s.cog), soc = static_cast<T>(s.soc), eps = static_cast<T>(s.eps);
  auto row = [&](u32 p) { return pos + u64(p) * stride; };

  // ---- init_solver_state (nlsolver.h:3679-3729)
  T x_inf = T(0);
  for (u32 j = lane; j < n; j += 32) { const T t = nm_abs(x0[j]); x_inf = t > x_inf ? t : x_inf; }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) { const T o = __shfl_xor_sync(kFull, x_inf, off); x_inf = o > x_inf ? o : x_inf; }
  const T a_ = x_inf < T(1.0) ? T(1.0) : x_inf;
  const T scale = a_ < T(10) ? a_ : T(10);
  const double nn = static_cast<double>(static_cast<T>(n));
  // x[i] + ((1.0 - sqrt(n + 1.0)) / n * scale), evaluated in double as the reference's expression is
  const double shift0 = __dmul_rn(__ddiv_rn(__dsub_rn(1.0, sqrt(__dadd_rn(nn, 1.0))), nn), static_cast<double>(scale));
  const u64 key0 = tape_key(tape_gen_key(s.seed, 0), s.offset + c);
  for (u32 p = 0; p < N; p++)
    for (u32 j = lane; j < n; j += 32) {
      T xv, vv = T(0);
      if (p == 0) xv = static_cast<T>(__dadd_rn(static_cast<double>(x0[j]), shift0));
      else if (p < ns) xv = (p == j) ? A::add(x0[j], scale) : x0[j];       // vertex n == x (the [n][n] store is dropped)
      else {
        const T up = static_cast<T>(fabs(__dmul_rn(2.5, static_cast<double>(x0[j])))), lo = -up;   // minimize(), :3583-3593
        const T span = A::sub(up, lo), temp = nm_abs(span);
        const u64 k = 2 * (u64(p - ns) * n + j);
        xv = A::add(lo, A::mul(span, unit<T>(tape_draw(key0, k))));
        vv = A::add(-temp, A::mul(unit<T>(tape_draw(key0, k + 1)), temp));
      }
      row(p)[j] = xv;
      vel[u64(p) * stride + j] = vv;
    }
  __syncwarp();
  u64 evals = 0;
  for (u32 p = 0; p < N; p++) {
    const T v = A::mul(fm, nm_eval<T, OBJ>(row(p), n, lane));
    if (lane == 0) val[p] = v;
    evals++;
  }
  for (u32 p = lane; p < N; p += 32) order[p] = p;
  __syncwarp();
  // rank sort by (value, particle id): ascending values, the id breaks ties
  auto sort_order = [&]() {
    for (u32 p = lane; p < N; p += 32) {
      const T v = val[p];
      u32 rank = 0;
      for (u32 q = 0; q < N; q++) {
        const T w = val[q];
        rank += (w < v || (w == v && q < p)) ? 1u : 0u;
      }
      scratch[rank] = p;
    }
    __syncwarp();
    for (u32 p = lane; p < N; p += 32) order[p] = scratch[p];
    __syncwarp();
  };
  auto transform = [&](const T *point, T *result, T coef, bool reflect) {      // simplex_transform, :1988-2010
    for (u32 j = lane; j < n; j += 32) {
      const T cj = cen[j];
      result[j] = reflect ? A::add(cj, A::mul(coef, A::sub(cj, point[j]))) : A::add(cj, A::mul(coef, A::sub(point[j], cj)));
    }
    __syncwarp();
  };
  auto copy_row = [&](T *dst, const T *src) {
    for (u32 j = lane; j < n; j += 32) dst[j] = src[j];
    __syncwarp();
  };
  u64 iter = 0, no_change = 0;
  const T best_val = val[0];                                                   // :3650, never updated
  for (;;) {
    sort_order();                                                              // :3654-3658
    const bool same = best_val == val[order[0]];
    no_change = same ? no_change + 1 : 0;                                      // :3663-3664
    bool stop = iter >= s.max_iter || no_change >= s.no_change_limit;
    if (!stop) {                                                               // simplex_std_err, :3901-3918 (every lane)
      T mean_val = T(0), result = T(0);
      for (u32 i = 0; i < ns; i++) mean_val = A::add(mean_val, val[order[i]]);
      mean_val = mean_val / static_cast<T>(ns);
      for (u32 i = 0; i < ns; i++) {
        const double dlt = static_cast<double>(A::sub(val[order[i]], mean_val));
        result = static_cast<T>(__dadd_rn(static_cast<double>(result), __dmul_rn(dlt, dlt)));
      }
      result = result / static_cast<T>(ns - 1);
      stop = static_cast<T>(sqrt(static_cast<double>(result))) < eps;
    }
    if (stop) break;
    // ---- apply_simplex (:3731-3822)
    {
      const T best_score = val[order[0]];
      const u32 worst = order[ns - 1], second = order[ns - 2];
      for (u32 j = lane; j < n; j += 32) {                                     // update_centroid, :3869-3885
        T acc = T(0);
        for (u32 i = 0; i + 1 < ns; i++) acc = A::add(acc, row(order[i])[j]);
        cen[j] = acc / static_cast<T>(ns - 1);
      }
      __syncwarp();
      transform(row(worst), t_ref, alpha, true);
      const T ref_score = A::mul(fm, nm_eval<T, OBJ>(t_ref, n, lane));
      evals++;
      if (ref_score >= best_score && ref_score < val[second]) {
        copy_row(row(worst), t_ref);
        if (lane == 0) val[worst] = ref_score;
      } else if (ref_score < best_score) {
        transform(t_ref, t_exp, gamma, false);
        const T exp_score = A::mul(fm, nm_eval<T, OBJ>(t_exp, n, lane));
        evals++;
        copy_row(row(worst), exp_score < ref_score ? t_exp : t_ref);
        if (lane == 0) val[worst] = exp_score < ref_score ? exp_score : ref_score;
      } else {
        const T worst_score = val[worst];
        transform(ref_score < worst_score ? t_ref : row(worst), t_con, rho, false);
        const T cont_score = A::mul(fm, nm_eval<T, OBJ>(t_con, n, lane));
        evals++;
        const T lim = ref_score < worst_score ? ref_score : worst_score;      // std::min(ref_score, worst_score)
        if (cont_score < lim) {
          copy_row(row(worst), t_con);
          if (lane == 0) val[worst] = cont_score;
        } else {
          const T *best = row(order[0]);                                       // shrink, :3886-3900
          for (u32 i = 1; i < ns; i++) {
            T *cur = row(order[i]);
            for (u32 j = lane; j < n; j += 32) cur[j] = A::add(best[j], A::mul(sigma, A::sub(cur[j], best[j])));
          }
          __syncwarp();
          for (u32 i = 1; i < ns; i++) {
            const T v = A::mul(fm, nm_eval<T, OBJ>(row(order[i]), n, lane));
            if (lane == 0) val[order[i]] = v;
          }
          evals += ns - 1;
          __syncwarp();
          sort_order();
        }
      }
      __syncwarp();
    }
    // ---- apply_pso (:3824-3868)
    {
      const u64 key = tape_key(tape_gen_key(s.seed, iter + 1), s.offset + c);
      bool order_flip = false;
      u32 best_in_pair = order[ns];
      const T *best = row(order[0]);
      for (u32 i = ns; i < N; i++) {
        const u32 id = order[i];
        if (order_flip) best_in_pair = order[i + 1];
        order_flip = ((i - ns) % 2) != 0;
        T *particle = row(id);
        const T *velocity = vel + u64(id) * stride, *pairwise = row(best_in_pair);
        for (u32 j = lane; j < n; j += 32) {
          const u64 k = 2 * (u64(i - ns) * n + j);
          const T r_p = unit<T>(tape_draw(key, k)), r_g = unit<T>(tape_draw(key, k + 1));
          const T pj = particle[j];
          const T t1 = A::mul(inertia, velocity[j]);
          const T t2 = A::mul(A::mul(cog, r_p), A::sub(pairwise[j], pj));
          const T t3 = A::mul(A::mul(soc, r_g), A::sub(best[j], pj));
          particle[j] = A::add(pj, A::add(A::add(t1, t2), t3));
        }
        __syncwarp();
        const T v = A::mul(fm, nm_eval<T, OBJ>(particle, n, lane));
        if (lane == 0) val[id] = v;
        evals++;
        __syncwarp();
      }
    }
    iter++;
  }
  // x = particle_positions[current_order[0]]; solver_status(value, iter, function_calls_used)   (:3672-3677)
  const u32 b = order[0];
  T *xb = static_cast<T *>(s.x_best) + c * s.d;
  for (u32 j = lane; j < n; j += 32) xb[j] = row(b)[j];
  if (lane == 0) {
    static_cast<T *>(s.f_best)[c] = val[b];
    s.iters[c] = iter;
    s.evals[c] = evals;
  }
}

inline size_t nmpso_smem_bytes(size_t n, size_t elem) {
  const size_t N = 3 * n + 1;
  return kNMPSOWarps * N * (elem + 2 * sizeof(unsigned int)) + 16;
}

#ifndef NLS_PLUGIN_BUILD
template <class T, int O>
cudaError_t nmpso_launch_o(const NMPSOState &s, cudaStream_t st) {
  auto kernel = nmpso_solve_kernel<T, O>;
  const size_t smem = nmpso_smem_bytes(s.d, sizeof(T));
  static size_t allowed_of[64] = {};                       // function attributes are per DEVICE
  int dev = 0;
  cudaGetDevice(&dev);
  size_t &allowed = allowed_of[dev % 64];
  if (smem > allowed) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    allowed = smem;
  }
  const unsigned grid = static_cast<unsigned>((s.C + kNMPSOWarps - 1) / kNMPSOWarps);
  kernel<<<grid, kNMPSOWarps * 32, smem, st>>>(s);
  return cudaGetLastError();
}
template <class T>
cudaError_t nmpso_launch(const NMPSOState &s, cudaStream_t st) {
  switch (s.objective) {
#define NLS_NM_CASE(O) case O: return nmpso_launch_o<T, O>(s, st);
    NLS_NM_CASE(OBJ_SPHERE) NLS_NM_CASE(OBJ_ROSENBROCK) NLS_NM_CASE(OBJ_RASTRIGIN) NLS_NM_CASE(OBJ_ACKLEY)
    NLS_NM_CASE(OBJ_ROSENBROCK_EX) NLS_NM_CASE(OBJ_BEALE) NLS_NM_CASE(OBJ_GOLDSTEIN_PRICE)
    NLS_NM_CASE(OBJ_THREE_HUMP_CAMEL) NLS_NM_CASE(OBJ_MCCORMICK) NLS_NM_CASE(OBJ_SCHAFFER_N2)
    NLS_NM_CASE(OBJ_STYBLINSKI_TANG) NLS_NM_CASE(OBJ_SHEKEL) NLS_NM_CASE(OBJ_BOOTH) NLS_NM_CASE(OBJ_BUKIN_N6)
    NLS_NM_CASE(OBJ_MATYAS) NLS_NM_CASE(OBJ_LEVI_N13)
#undef NLS_NM_CASE
    default: return cudaErrorInvalidValue;
  }
}
#endif

}  // namespace nls
