// de_impl.cuh — the DE generation on B200.  A warp owns a tile of agents: per-agent scalar work runs one agent per lane,
// row streaming runs one agent per warp (or per 4 / 8 / 16-lane group for short rows); rows live agent-major in HBM and
// are streamed with 128-bit loads; the objective is a lane-strided sum closed by a butterfly.
//
// Reference semantics being reproduced (DE::solve, nlsolver.h:2413-2476): agents are processed sequentially and IN
// PLACE — agent i reads donor r's row *after* r's own greedy selection when r < i and *before* it when r > i
// (nlsolver.h:2466-2471 vs 2368-2372).  A synchronous double-buffered generation is a different algorithm
// (SURVEY.md §7.3 item 1), so the generation is run as speculate + exact repair:
//
//   K2  de_generation_kernel : every agent builds and scores its trial against the PRE-generation rows (full HBM
//                              bandwidth, no ordering), records donors / dim / rejects, accepts greedily; an accepted
//                              trial is written to the agent's row in the OTHER buffer, so pre-generation rows stay.
//   K2r de_repair_kernel     : cooperative kernel.  The "donor r < i" relation is a shallow DAG (depth ~17 at 2^20).
//                              Round by round, an agent whose lower donors are all final becomes final; if any of
//                              them was accepted, its trial is re-evaluated against the now-known rows.  Each agent
//                              is therefore evaluated at most twice and the result equals the sequential loop.
//                              Returns at once when the speculative pass accepted nothing.
//   K3  de_commit_kernel     : commits accepted trials (score, row-location bit), then the population reduction:
//                              min-loc with the reference's tie rule (nlsolver.h:2432-2437), the std_err statistic
//                              (nlsolver.h:2037-2052) and the stop test (nlsolver.h:2439-2447), last block finalises.
#pragma once
#include <cooperative_groups.h>
#include <math_constants.h>

#include "objectives.cuh"
#include "reduce.cuh"
#include "launch.h"
#include "state.h"

namespace nls {
namespace cg = cooperative_groups;


// ------------------------------------------------------------------------------------------------ row pass
template <class T, bool RESOLVED>
__device__ __forceinline__ const T *de_row_of(const DEState &s, u64 r, u64 i) {
  u32 w = s.where[r];
  if (RESOLVED && r < i && s.acc[r]) w ^= 1u;   // a lower donor whose trial was accepted already holds its new row
  return static_cast<const T *>(s.buf[w]) + r * s.stride;
}

// One sweep over the d coordinates of agent i's trial (propose_new_agent, nlsolver.h:2357-2375):
//   trial[j] = mut ? A[r1][j] + F * (A[r2][j] - A[r3][j]) : A[r0][j],   mut = (draw_j < CR) || (j == dim)
// EVAL accumulates the objective; WRITE stores the trial into `dst`.  Each lane owns V consecutive coordinates per
// step (one 128-bit load per row); U steps are issued back to back so 4*U loads per lane are in flight.  Lanes past
// the end of the row read coordinate 0 instead (always valid, one cached line) and are masked out downstream.
// Tuning (B200, fp64 d=1000 Rastrigin, P=2^20, K2 time): U=2 / 2 blocks per SM 6.29 ms; U=2 / 3 blocks 5.68 ms;
// U=1 / 4 blocks (64 registers, 32 warps per SM, no spills) 5.63 ms — occupancy beats per-warp unrolling here.
#ifndef NLS_DE_UNROLL
#define NLS_DE_UNROLL 1
#endif
// W lanes cooperate on the agent (`lane` is the index inside that group); W < 32 requires d <= W * V.
// SHARED_BASE: the base row is the single best row (best recombination, speculative pass): read-only for the launch and
// shared by every agent, so it goes through L1 instead of being re-fetched from L2 per agent.
template <class T, int OBJ, bool EVAL, bool WRITE, int W = 32, bool SHARED_BASE = false>
__device__ __forceinline__ T de_sweep(const DEState &s, const T *p0, const T *p1, const T *p2, const T *p3, T *dst,
                                      u64 sbase, u32 dim, u64 i, int lane) {
  constexpr int V = Vec<T>::V;
  constexpr int U = NLS_DE_UNROLL;
  constexpr u32 kStride = W * V;                       // coordinates per group step
  typedef Ar<T> A;
  const u32 d = static_cast<u32>(s.d);
  const T F = static_cast<T>(s.F);
  const u64 cr_le = s.cr_le;
  const bool cr_any = !s.cr_none;
  const u32 n_steps = (d + kStride - 1) / kStride;
  Objective<T, OBJ, W> obj;
  if (EVAL) obj.begin(lane, d);
  u32 j0 = lane * V;                                   // first coordinate of this lane in the current step
  u64 st = sbase + kGolden * j0;                       // draw-stream state of coordinate j0
  for (u32 step = 0; step < n_steps; step += U) {
    T x1[U][V], x2[U][V], x3[U][V], x0[U][V];
    bool mut[U][V];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const u32 jj = j0 + u * kStride;
      bool all_mut = true;
#pragma unroll
      for (int q = 0; q < V; q++) {
        mut[u][q] = (cr_any && mix64(st + kGolden * (u * kStride + q)) <= cr_le) || (jj + q == dim);
        all_mut &= mut[u][q];
      }
      const u32 jl = jj < d ? jj : 0u;
      ld_row(p1 + jl, x1[u]); ld_row(p2 + jl, x2[u]); ld_row(p3 + jl, x3[u]);
      if (!all_mut) {                                  // the base row is only touched where a coordinate keeps it
        if (SHARED_BASE) ld_row_shared(p0 + jl, x0[u]);
        else ld_row(p0 + jl, x0[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const u32 jj = j0 + u * kStride;
      T t[V];
#pragma unroll
      for (int q = 0; q < V; q++)
        t[q] = mut[u][q] ? A::add(x1[u][q], A::mul(F, A::sub(x2[u][q], x3[u][q]))) : x0[u][q];
      if (WRITE && jj < d) st_row(dst + jj, t);
      if (EVAL) {
        obj.step(t, jj, d, lane);
        if (s.masks != nullptr) {
#pragma unroll
          for (int q = 0; q < V; q++)
            if (jj + q < d) s.masks[i * d + jj + q] = mut[u][q];
        }
      }
    }
    j0 += U * kStride;
    st += kGolden * (U * kStride);
  }
  return EVAL ? obj.finish(d) : T(0);
}

// Build, score and greedily select agent i's trial (loop body nlsolver.h:2459-2471).
template <class T, int OBJ, bool RESOLVED>
__device__ __forceinline__ void de_trial(const DEState &s, u64 i, u64 key, u64 r0, u64 r1, u64 r2, u64 r3, u32 dim,
                                         u32 rej, int lane) {
  const T *p0 = de_row_of<T, RESOLVED>(s, r0, i);
  const T *p1 = de_row_of<T, RESOLVED>(s, r1, i);
  const T *p2 = de_row_of<T, RESOLVED>(s, r2, i);
  const T *p3 = de_row_of<T, RESOLVED>(s, r3, i);
  T *dst = static_cast<T *>(s.buf[s.where[i] ^ 1u]) + i * s.stride;
  const u64 sbase = tape_state(key, 4 + rej);           // draws 0..2+rej: indices, 3+rej: dim, then one per coordinate
  const T raw = de_sweep<T, OBJ, true, false>(s, p0, p1, p2, p3, dst, sbase, dim, i, lane);
  const T score = Ar<T>::mul(static_cast<T>(s.fm), raw);
  const bool ok = score < static_cast<const T *>(s.score)[i];   // strict <, NaN never accepted (nlsolver.h:2466)
  if (ok)   // re-stream the (L2-warm) donor rows and materialise the trial; ~a few % of agents
    de_sweep<T, OBJ, false, true>(s, p0, p1, p2, p3, dst, sbase, dim, i, lane);
  if (lane == 0) {
    static_cast<T *>(s.tscore)[i] = score;
    s.acc[i] = ok;
  }
}

// ------------------------------------------------------------------------------------------------ K1 init
// init_agents / generate_sequence + initial scoring (nlsolver.h:2302-2323, 2423-2425):
//   agent[j] = (g() - 0.5) * x0[j]   evaluated in double even for T = float, like the reference expression
template <class T, int OBJ>
__global__ void __launch_bounds__(kBlock) de_init_kernel(DEState s, const T *__restrict__ x0) {
  constexpr int V = Vec<T>::V;
  const int lane = threadIdx.x & 31;
  const u64 warp = (u64(blockIdx.x) * kBlock + threadIdx.x) >> 5, n_warps = (u64(gridDim.x) * kBlock) >> 5;
  const u64 d = s.d, gen_key = tape_gen_key(s.seed, 0);
  const u64 n_steps = (d + 32 * V - 1) / (32 * V);
  for (u64 i = warp; i < s.P; i += n_warps) {
    const u64 key = tape_key(gen_key, s.offset + i);
    T *row = static_cast<T *>(s.buf[0]) + i * s.stride;
    Objective<T, OBJ> obj;
    obj.begin(lane, u32(d));
    for (u64 st = 0; st < n_steps; st++) {
      const u64 j0 = (st * 32 + lane) * V;
      T t[V];
#pragma unroll
      for (int q = 0; q < V; q++) {
        const u64 j = j0 + q;
        t[q] = T(0);
        if (j < d)
          t[q] = static_cast<T>(__dmul_rn(__dsub_rn(static_cast<double>(unit<T>(tape_draw(key, j))), 0.5),
                                          static_cast<double>(x0[j])));
      }
      if (j0 < d) st_row(row + j0, t);
      obj.step(t, u32(j0), u32(d), lane);
    }
    const T score = Ar<T>::mul(static_cast<T>(s.fm), obj.finish(u32(d)));
    if (lane == 0) {
      static_cast<T *>(s.score)[i] = score;
      static_cast<T *>(s.tscore)[i] = score;
      s.where[i] = 0;
      s.acc[i] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ K2 generation
// generate_indices (nlsolver.h:2331-2355): draw until three proposals differ from `fixed` and from each other.
template <class T>
__device__ __forceinline__ void de_select_donors(u64 key, u64 P, u64 fixed, u64 &r1, u64 &r2, u64 &r3, u32 &rej) {
  u64 k = 0;
  rej = 0;
  for (;;) { r1 = index_from<T>(tape_draw(key, k++), P); if (r1 != fixed) break; rej++; }
  for (;;) { r2 = index_from<T>(tape_draw(key, k++), P); if (r2 != fixed && r2 != r1) break; rej++; }
  for (;;) { r3 = index_from<T>(tape_draw(key, k++), P); if (r3 != fixed && r3 != r1 && r3 != r2) break; rej++; }
}

#ifndef NLS_DE_MINBLOCKS
#define NLS_DE_MINBLOCKS 4
#endif
// What the lane-parallel prologue hands to the cooperative part, one entry per agent of the tile (shared memory).
struct __align__(16) DETileEntry {
  unsigned long long key;      // draw stream of (generation, agent)
  unsigned int r1, r2, r3;     // donors ids[1..3]
  unsigned int dim;            // forced crossover coordinate
  unsigned int rej;            // rejected index proposals
  unsigned int wbits;          // where[] of ids[0], r1, r2, r3 and of the agent itself (bits 0..4)
  double score;                // scores[i] (T widened)
};

// A warp owns a TILE of 32 consecutive agents:
//   prologue  — one agent per lane: stream key, donor selection, forced coordinate, row-location bits, current score;
//               the per-agent scalar work and its dependent loads run 32 agents wide instead of 32 times redundantly;
//   body      — the 32 agents one after the other, all lanes streaming the rows of one agent (de_sweep);
//   epilogue  — one agent per lane again: coalesced stores of trial score / accept flag.
// The tile size (2^tile_shift <= 32 agents) is chosen by the launcher so that every warp gets many tiles: with 32-agent
// tiles a population of 2^18 would give 1.7 tiles per resident warp and a 15 % tail.
// W lanes per agent in the body: 32, or 16 / 8 / 4 when one step of W lanes covers the row (d <= W * V) — then the warp
// streams 32 / W agents of the tile at a time instead of leaving lanes idle.
template <class T, int OBJ, int W>
__global__ void __launch_bounds__(kBlock, NLS_DE_MINBLOCKS) de_generation_kernel(DEState s, int tile_shift) {
  __shared__ DETileEntry tile_mem[kWarpsPerBlock][32];
  DECtrl *ctrl = s.ctrl;
  if (ctrl->stop) return;
  const int lane = threadIdx.x & 31;
  DETileEntry *tile_entries = tile_mem[threadIdx.x >> 5];
  const u64 warp = (u64(blockIdx.x) * kBlock + threadIdx.x) >> 5, n_warps = (u64(gridDim.x) * kBlock) >> 5;
  const u64 gen_key = tape_gen_key(s.seed, ctrl->iter + 1), best_id = ctrl->best_id;
  const bool random_mode = s.strategy != 0;             // NLS_DE_RANDOM = 1: ids[0] = i; best: ids[0] = best_id
  const u64 P = s.P, tile_size = 1ull << tile_shift, n_tiles = (P + tile_size - 1) >> tile_shift;
  const T *score = static_cast<const T *>(s.score);
  for (u64 tile = warp; tile < n_tiles; tile += n_warps) {
    const u64 first = tile << tile_shift, mine = first + lane;
    const bool own = lane < int(tile_size) && mine < P;   // this lane carries an agent in the prologue / epilogue
    // ---- prologue
    bool level0 = true;
    if (own) {
      DETileEntry e;
      e.key = tape_key(gen_key, s.offset + mine);
      const u64 fixed = random_mode ? mine : best_id;
      u64 r1, r2, r3;
      de_select_donors<T>(e.key, P, fixed, r1, r2, r3, e.rej);
      e.dim = static_cast<u32>(index_from<T>(tape_draw(e.key, 3 + e.rej), s.d));
      e.r1 = u32(r1); e.r2 = u32(r2); e.r3 = u32(r3);
      e.wbits = u32(s.where[fixed]) | (u32(s.where[r1]) << 1) | (u32(s.where[r2]) << 2) | (u32(s.where[r3]) << 3) |
                (u32(s.where[mine]) << 4);
      e.score = static_cast<double>(score[mine]);
      tile_entries[lane] = e;
      s.dec[mine] = make_uint4(e.r1, e.r2, e.r3, e.dim);
      s.rej[mine] = e.rej;
      // agents without a lower donor are final as they stand: repair round 1 is decided here
      level0 = r1 > mine && r2 > mine && r3 > mine && (random_mode || best_id >= mine);
      s.fin[mine] = level0 ? 1 : 0;
    }
    {   // everybody else goes on the repair's first pending list (order is irrelevant)
      const u32 vote = __ballot_sync(kFull, own && !level0);
      if (vote) {
        u32 slot = 0;
        if (lane == 0) slot = atomicAdd(&ctrl->pending[1], __popc(vote));   // consumed by repair round 2
        slot = __shfl_sync(kFull, slot, 0);
        if (own && !level0) s.pend[0][slot + __popc(vote & ((1u << lane) - 1u))] = u32(mine);
      }
    }
    __syncwarp();
    // ---- body
    const int n_here = (P - first) < tile_size ? int(P - first) : int(tile_size);
    T my_score = T(0);
    bool my_ok = false;
    constexpr int G = 32 / W;                              // agents streamed at a time
    const int grp = lane / W, sub = lane % W;
    for (int a0 = 0; a0 < n_here; a0 += G) {
      const bool active = a0 + grp < n_here;               // idle groups redo the first agent, results discarded
      const int a = active ? a0 + grp : a0;
      const DETileEntry e = tile_entries[a];               // broadcast read inside the group
      const u64 i = first + a;
      const u64 r0 = random_mode ? i : best_id;
      const T *p0 = static_cast<const T *>(s.buf[e.wbits & 1u]) + r0 * s.stride;
      const T *p1 = static_cast<const T *>(s.buf[(e.wbits >> 1) & 1u]) + u64(e.r1) * s.stride;
      const T *p2 = static_cast<const T *>(s.buf[(e.wbits >> 2) & 1u]) + u64(e.r2) * s.stride;
      const T *p3 = static_cast<const T *>(s.buf[(e.wbits >> 3) & 1u]) + u64(e.r3) * s.stride;
      T *dst = static_cast<T *>(s.buf[((e.wbits >> 4) & 1u) ^ 1u]) + i * s.stride;
      const u64 sbase = tape_state(e.key, 4 + e.rej);
      // (warp-uniform branch: in best mode every agent's base row is the pre-generation best row)
      const T raw = random_mode ? de_sweep<T, OBJ, true, false, W, false>(s, p0, p1, p2, p3, dst, sbase, e.dim, i, sub)
                                : de_sweep<T, OBJ, true, false, W, true>(s, p0, p1, p2, p3, dst, sbase, e.dim, i, sub);
      const T sc = Ar<T>::mul(static_cast<T>(s.fm), raw);
      const bool ok = active && sc < static_cast<T>(e.score);   // strict <, NaN never accepted (nlsolver.h:2466)
      if (ok) de_sweep<T, OBJ, false, true, W>(s, p0, p1, p2, p3, dst, sbase, e.dim, i, sub);
      // hand the outcome of agent a0 + g to lane a0 + g (its owner in the epilogue)
#pragma unroll
      for (int g = 0; g < G; g++) {
        const T sc_g = __shfl_sync(kFull, sc, g * W);
        const bool ok_g = __shfl_sync(kFull, int(ok), g * W) != 0;
        if (lane == a0 + g) { my_score = sc_g; my_ok = ok_g; }
      }
    }
    // ---- epilogue
    if (own) {
      static_cast<T *>(s.tscore)[mine] = my_score;
      s.acc[mine] = my_ok;
    }
    const u32 n_ok = __popc(__ballot_sync(kFull, my_ok));
    if (lane == 0 && n_ok) atomicAdd(&ctrl->spec_accepted, n_ok);
    __syncwarp();                                          // tile_entries is rewritten by the next tile's prologue
  }
}

// ------------------------------------------------------------------------------------------------ K2r repair
template <class T, int OBJ>
__global__ void __launch_bounds__(kBlock) de_repair_kernel(DEState s) {
  cg::grid_group grid = cg::this_grid();
  DECtrl *ctrl = s.ctrl;
  if (ctrl->stop) return;                                // uniform over the grid: only K3 changes it
  // If the speculative pass accepted nothing, no row changed and every speculative result already equals the
  // sequential one (induction over the agent index) — nothing to repair.
  if (*reinterpret_cast<volatile unsigned int *>(&ctrl->spec_accepted) == 0) return;
  const int lane = threadIdx.x & 31;
  const u64 tid = u64(blockIdx.x) * kBlock + threadIdx.x, n_threads = u64(gridDim.x) * kBlock;
  const u64 warp = tid >> 5, n_warps = n_threads >> 5;
  const u64 gen_key = tape_gen_key(s.seed, ctrl->iter + 1), best_id = ctrl->best_id;
  const bool best_mode = s.strategy == 0;
  u32 reruns = 0, round = 2;                             // round 1 (agents without lower donors) was decided by K2
  // Round r consumes the pending list produced by round r - 1 (buffer pend[r & 1], length pending[(r - 1) % 3]; K2
  // produced the first one) and produces pend[(r & 1) ^ 1] / pending[r % 3] plus the re-evaluation list
  // list / list_count[r % 3].  Three counter slots let thread 0 recycle slot (r + 1) % 3 between the two barriers of
  // round r: its last reader finished a round ago, its next writer starts after the second barrier.
  for (;; round++) {
    const u32 cur = round % 3u, prev = (round + 2u) % 3u, nxt = (round + 1u) % 3u;
    const u32 *pin = s.pend[round & 1u];
    u32 *pout = s.pend[(round & 1u) ^ 1u];
    // phase A: classify the pending agents.  A donor is "final for this round" iff it was finalised in an EARLIER
    // round (0 < fin < round), which makes the outcome independent of the order in which threads run.
    const u32 n_in = *reinterpret_cast<volatile unsigned int *>(&ctrl->pending[prev]);
    // Phase A is a chain of dependent L2 reads per entry (list -> donors -> their state -> their donors -> state), so
    // each thread walks UA entries at once with the loads of one level issued together (predicated, no branches).
    constexpr int UA = 2;
    for (u64 base = warp * (32 * UA); base < n_in; base += n_warps * (32 * UA)) {
      u32 ii[UA];
      bool valid[UA], wait[UA], dirty[UA];
      uint4 dc[UA];
#pragma unroll
      for (int u = 0; u < UA; u++) {
        const u64 idx = base + u * 32 + lane;
        valid[u] = idx < n_in;
        ii[u] = valid[u] ? pin[idx] : 0u;
      }
#pragma unroll
      for (int u = 0; u < UA; u++) dc[u] = s.dec[ii[u]];
      // level 1: the (up to four) lower donors of the entry
      u32 don[UA][4], f1[UA][4];
      bool low[UA][4];
#pragma unroll
      for (int u = 0; u < UA; u++) {
        don[u][0] = dc[u].x; don[u][1] = dc[u].y; don[u][2] = dc[u].z; don[u][3] = u32(best_id);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          low[u][k] = valid[u] && don[u][k] < ii[u] && (k < 3 || best_mode);
          f1[u][k] = low[u][k] ? u32(s.fin[don[u][k]]) : 1u;
        }
      }
      // A donor is "settled" iff it was finalised in an EARLIER round (0 < fin < round): everything read here was
      // written before this round's barrier, so the outcome does not depend on the order in which threads run.
      u32 a1[UA][4];
      uint4 dd[UA][4];
      bool deep[UA][4];
#pragma unroll
      for (int u = 0; u < UA; u++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const bool settled = f1[u][k] != 0 && f1[u][k] < round;
          deep[u][k] = low[u][k] && !settled;
          a1[u][k] = low[u][k] ? u32(s.acc[don[u][k]]) : 0u;        // speculative / settled accept flag of the donor
          dd[u][k] = deep[u][k] ? s.dec[don[u][k]] : make_uint4(0, 0, 0, 0);
        }
      // level 2: a donor r that is still pending is looked through — if all of ITS lower donors are settled and none of
      // them was accepted, r becomes final this round WITHOUT re-evaluation, i.e. with the accept flag it already has,
      // and the entry can rely on that now instead of waiting a round (two DAG levels per barrier).
#pragma unroll
      for (int u = 0; u < UA; u++) {
        wait[u] = false; dirty[u] = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          bool clean = true;
          if (deep[u][k]) {
            const u32 r = don[u][k];
            const u32 q[4] = {dd[u][k].x, dd[u][k].y, dd[u][k].z, u32(best_id)};
#pragma unroll
            for (int m = 0; m < 4; m++) {
              const bool lowq = q[m] < r && (m < 3 || best_mode);
              const u32 fq = lowq ? u32(s.fin[q[m]]) : 1u;
              const u32 aq = lowq ? u32(s.acc[q[m]]) : 0u;
              clean &= (fq != 0 && fq < round) && aq == 0;
            }
          }
          if (low[u][k]) {
            if (deep[u][k] && !clean) wait[u] = true;
            else dirty[u] |= a1[u][k] != 0;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UA; u++) {
        const bool w_ = valid[u] && wait[u];
        const bool rerun = valid[u] && !wait[u] && dirty[u];
        if (valid[u] && !wait[u]) s.fin[ii[u]] = uint16_t(round);
        const u32 vote_w = __ballot_sync(kFull, w_), vote_r = __ballot_sync(kFull, rerun);
        if (vote_w | vote_r) {
          u32 slot_w = 0, slot_r = 0;
          if (lane == 0) {
            if (vote_w) slot_w = atomicAdd(&ctrl->pending[cur], __popc(vote_w));
            if (vote_r) slot_r = atomicAdd(&ctrl->list_count[cur], __popc(vote_r));
          }
          slot_w = __shfl_sync(kFull, slot_w, 0);
          slot_r = __shfl_sync(kFull, slot_r, 0);
          const u32 below = (1u << lane) - 1u;
          if (w_) pout[slot_w + __popc(vote_w & below)] = ii[u];
          if (rerun) s.list[slot_r + __popc(vote_r & below)] = ii[u];
        }
      }
    }
    grid.sync();
    const u32 n_list = *reinterpret_cast<volatile unsigned int *>(&ctrl->list_count[cur]);
    const u32 n_out = *reinterpret_cast<volatile unsigned int *>(&ctrl->pending[cur]);
    if (tid == 0) { ctrl->pending[nxt] = 0; ctrl->list_count[nxt] = 0; }
    // phase B: re-evaluate the listed agents against rows that are now known
    for (u64 e = warp; e < n_list; e += n_warps) {
      const u64 i = s.list[e];
      const uint4 dc = s.dec[i];
      const u64 key = tape_key(gen_key, s.offset + i);
      de_trial<T, OBJ, true>(s, i, key, best_mode ? best_id : i, dc.x, dc.y, dc.z, dc.w, s.rej[i], lane);
    }
    if (warp == 0) reruns += n_list;
    if (n_out == 0) break;
    if (round >= 65000u) { if (tid == 0) ctrl->error = 1; break; }
    grid.sync();   // phase B's rows / accept flags and the recycled counters are visible to the next round
  }
  if (tid == 0) { ctrl->reruns += reruns; ctrl->rounds += round - 1; }
}

// ------------------------------------------------------------------------------------------------ K3 commit + reduce
template <class T>
// mode 0: commit the generation in flight; 1: first scan after init; 2: re-scan after island migration (best only)
__global__ void __launch_bounds__(kBlock) de_commit_kernel(DEState s, int mode) {
  const bool initial = mode != 0;
  DECtrl *ctrl = s.ctrl;
  if (ctrl->stop) return;
  T *score = static_cast<T *>(s.score);
  const T *tscore = static_cast<const T *>(s.tscore);
  u32 n_acc = 0;
  auto item = [&](u64 i, double &for_min, double &for_moments) {
    if (!initial && s.acc[i]) { score[i] = tscore[i]; s.where[i] ^= 1u; n_acc++; }
    for_min = for_moments = static_cast<double>(score[i]);
  };
  auto fin = [&](MinLoc ml, Moments mo) {
    if (!initial) ctrl->iter += 1;                        // nlsolver.h:2474
    const u64 prev = ctrl->best_id;
    bool not_updated = true;                              // best scan, nlsolver.h:2430-2437: the scan keeps best_id
    if (ml.v < static_cast<double>(__ldcg(score + prev))) { ctrl->best_id = ml.i; not_updated = false; }
    ctrl->best_value = static_cast<double>(__ldcg(score + ctrl->best_id));
    ctrl->score_moments = mo;
    if (mode == 2) { if (!not_updated) ctrl->vnc = 0; return; }
    ctrl->vnc = not_updated ? ctrl->vnc + 1 : 0;          // nlsolver.h:2439
    int reason = 0;                                       // nlsolver.h:2441-2443, same short-circuit order
    if (ctrl->iter >= s.max_iter) reason = 1;
    else if (ctrl->vnc >= s.vnc_limit) reason = 2;
    else {
      const T se = stop_std_err<T>(mo, score, s.P, s.eps);
      ctrl->std_err = static_cast<double>(se);
      if (se < static_cast<T>(s.eps)) reason = 3;
    }
    ctrl->accepted += ctrl->acc_partial;
    ctrl->acc_partial = 0;
    ctrl->spec_accepted = 0;
    ctrl->pending[0] = ctrl->pending[1] = ctrl->pending[2] = 0;
    ctrl->list_count[0] = ctrl->list_count[1] = ctrl->list_count[2] = 0;
    if (ctrl->error) reason = reason ? reason : 4;
    ctrl->stop_reason = reason;
    __threadfence();
    ctrl->stop = reason != 0;
  };
  auto post = [&]() {
    n_acc = __reduce_add_sync(kFull, n_acc);
    if ((threadIdx.x & 31) == 0 && n_acc) atomicAdd(&ctrl->acc_partial, n_acc);
  };
  MinLoc ml;
  Moments mo;
  if (population_reduce(s.P, s.part_min, s.part_idx, s.part_mom, &ctrl->ticket, item, post, ml, mo) &&
      threadIdx.x == 0)
    fin(ml, mo);
}

// ------------------------------------------------------------------------------------------------ island hooks
// (no reference counterpart — SURVEY.md §8e: each GPU runs a reference-exact DE on its island; these kernels export the
//  island best, pick the k best emigrants and overwrite the k worst agents with immigrants, all deterministically.)
template <class T>
__global__ void __launch_bounds__(kBlock) de_export_best_kernel(DEState s, void *record) {
  const DECtrl *ctrl = s.ctrl;
  RecordHeader *h = static_cast<RecordHeader *>(record);
  const u64 b = ctrl->best_id;
  if (threadIdx.x == 0) {
    h->value = ctrl->best_value; h->index = s.offset + b; h->moments = ctrl->score_moments; h->valid = 1; h->_pad = 0;
  }
  T *row = reinterpret_cast<T *>(h + 1);
  const T *src = static_cast<const T *>(s.buf[s.where[b]]) + b * s.stride;
  for (u64 j = threadIdx.x; j < s.d; j += kBlock) row[j] = __ldcg(src + j);
}

// Top-k selection for migration, deterministic and in two kernels.  Order:
//   best-first  (sign = +1): ascending score, ascending index on ties
//   worst-first (sign = -1): descending score, descending index on ties
// Both are "ascending (key, visit)" with key = sign * score and visit = i or P - 1 - i.  Phase 1: every block extracts
// the k smallest (key, visit) pairs of its contiguous slice by k rounds of a block-wide arg-min that only admits pairs
// beyond the previous pick (no marking needed).  Phase 2: one block does the same over the blocks' candidates.
struct TopKey { double key; unsigned long long visit; };
__device__ __forceinline__ bool topkey_less(const TopKey &a, const TopKey &b) {
  return a.key < b.key || (a.key == b.key && a.visit < b.visit);
}
__device__ __forceinline__ TopKey topkey_block_min(TopKey mine, TopKey *sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    TopKey o; o.key = __shfl_down_sync(kFull, mine.key, off); o.visit = __shfl_down_sync(kFull, mine.visit, off);
    if (topkey_less(o, mine)) mine = o;
  }
  if (lane == 0) sm[w] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kWarpsPerBlock; k++) if (topkey_less(sm[k], mine)) mine = sm[k];
    sm[0] = mine;
  }
  __syncthreads();
  mine = sm[0];
  __syncthreads();
  return mine;
}
constexpr unsigned long long kNoVisit = ~0ull;

template <class T>
__global__ void __launch_bounds__(kBlock) de_topk_partial_kernel(DEState s, int sign, u32 k, u64 slice, TopKey *cand) {
  __shared__ TopKey sm[kWarpsPerBlock];
  const T *score = static_cast<const T *>(s.score);
  const u64 P = s.P, lo = u64(blockIdx.x) * slice, hi = (lo + slice < P) ? lo + slice : P;
  TopKey cursor; cursor.key = -CUDART_INF; cursor.visit = kNoVisit;   // kNoVisit + 1 wraps to 0: "before everything"
  bool started = false;
  for (u32 e = 0; e < k; e++) {
    TopKey best; best.key = CUDART_INF; best.visit = kNoVisit;
    for (u64 v = lo + threadIdx.x; v < hi; v += kBlock) {
      const u64 a = sign > 0 ? v : P - 1 - v;
      TopKey c; c.key = sign * static_cast<double>(score[a]); c.visit = v;
      const bool beyond = !started || topkey_less(cursor, c);
      if (beyond && topkey_less(c, best)) best = c;
    }
    best = topkey_block_min(best, sm);
    if (threadIdx.x == 0) cand[u64(blockIdx.x) * k + e] = best;
    if (best.visit == kNoVisit) {                       // slice exhausted: pad the remaining slots
      for (u32 r = e + 1 + threadIdx.x; r < k; r += kBlock) cand[u64(blockIdx.x) * k + r] = best;
      break;
    }
    cursor = best;
    started = true;
  }
}

// Phase 2: every block's candidate list is already in ascending (key, visit) order, so the global top-k is a k-step
// merge of the lists' heads: each thread looks at the heads of its lists (one load each), a block-wide arg-min picks the
// winner and advances that list.  (Scanning all candidates every round instead costs k * n_cand dependent L2 reads in a
// single block — 2 ms for 256 lists of 64, which used to dominate a migration.)
constexpr u32 kMaxTopkLists = 4096;
template <class T>
__global__ void __launch_bounds__(kBlock) de_topk_final_kernel(DEState s, int sign, u32 k, u32 n_lists, const TopKey *cand) {
  __shared__ TopKey sm[kWarpsPerBlock];
  __shared__ u32 sm_src[kWarpsPerBlock];
  __shared__ u32 heads[kMaxTopkLists];
  for (u32 b = threadIdx.x; b < n_lists; b += kBlock) heads[b] = 0;
  __syncthreads();
  for (u32 e = 0; e < k; e++) {
    TopKey best; best.key = CUDART_INF; best.visit = kNoVisit;
    u32 src = 0xffffffffu;
    for (u32 b = threadIdx.x; b < n_lists; b += kBlock) {
      const u32 h = heads[b];
      if (h >= k) continue;
      TopKey c; c.key = cand[u64(b) * k + h].key; c.visit = cand[u64(b) * k + h].visit;
      if (c.visit != kNoVisit && topkey_less(c, best)) { best = c; src = b; }
    }
    // block arg-min carrying the source list
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      TopKey o; o.key = __shfl_down_sync(kFull, best.key, off); o.visit = __shfl_down_sync(kFull, best.visit, off);
      const u32 os = __shfl_down_sync(kFull, src, off);
      if (topkey_less(o, best)) { best = o; src = os; }
    }
    if (lane == 0) { sm[w] = best; sm_src[w] = src; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int q = 1; q < kWarpsPerBlock; q++) if (topkey_less(sm[q], best)) { best = sm[q]; src = sm_src[q]; }
      s.list[e] = best.visit == kNoVisit ? 0xffffffffu : u32(sign > 0 ? best.visit : s.P - 1 - best.visit);
      if (best.visit != kNoVisit) heads[src] += 1;
    }
    __syncthreads();
  }
}

// copy the k selected rows / scores out (emigrants) ...
template <class T>
__global__ void __launch_bounds__(kBlock) de_gather_kernel(DEState s, u32 k, T *rows, T *scores) {
  for (u32 e = blockIdx.x; e < k; e += gridDim.x) {
    const u32 a = s.list[e];
    if (a == 0xffffffffu) continue;
    const T *src = static_cast<const T *>(s.buf[s.where[a]]) + u64(a) * s.stride;
    for (u64 j = threadIdx.x; j < s.d; j += kBlock) rows[u64(e) * s.d + j] = __ldcg(src + j);
    if (threadIdx.x == 0) scores[e] = static_cast<const T *>(s.score)[a];
  }
}
// ... or overwrite the k selected agents (immigrants): row in place, score, and nothing else
template <class T>
__global__ void __launch_bounds__(kBlock) de_scatter_kernel(DEState s, u32 k, const T *rows, const T *scores) {
  for (u32 e = blockIdx.x; e < k; e += gridDim.x) {
    const u32 a = s.list[e];
    if (a == 0xffffffffu) continue;
    T *dst = static_cast<T *>(s.buf[s.where[a]]) + u64(a) * s.stride;
    for (u64 j = threadIdx.x; j < s.d; j += kBlock) dst[j] = rows[u64(e) * s.d + j];
    if (threadIdx.x == 0) static_cast<T *>(s.score)[a] = scores[e];
  }
}

// compact rows [first, first + count) of the current population into `out` (host read-back path)
template <class T>
__global__ void __launch_bounds__(kBlock) de_gather_rows_kernel(DEState s, u64 first, u64 count, T *out) {
  for (u64 e = blockIdx.x; e < count; e += gridDim.x) {
    const u64 a = first + e;
    const T *src = static_cast<const T *>(s.buf[s.where[a]]) + a * s.stride;
    for (u64 j = threadIdx.x; j < s.d; j += kBlock) out[e * s.d + j] = __ldcg(src + j);
  }
}

// ------------------------------------------------------------------------------------------------ host launchers
template <class K>
inline int blocks_per_sm(K kernel) {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kBlock, 0);
  return n < 1 ? 1 : n;
}
inline unsigned int clamp_grid(u64 want, u64 cap) {
  const u64 g = want < cap ? want : cap;
  return static_cast<unsigned int>(g < 1 ? 1 : g);
}

#ifdef NLS_PLUGIN_BUILD   // an objective plugin instantiates the kernels for its own functor only
#define NLS_OBJ_SWITCH(obj, CALL)                             \
  switch (obj) {                                       \
    case OBJ_CUSTOM: { CALL(OBJ_CUSTOM); } break;      \
    default: return cudaErrorInvalidValue;             \
  }
#else
#define NLS_OBJ_SWITCH(obj, CALL)                      \
  switch (obj) {                                       \
    case OBJ_SPHERE: { CALL(OBJ_SPHERE); } break;      \
    case OBJ_ROSENBROCK: { CALL(OBJ_ROSENBROCK); } break; \
    case OBJ_RASTRIGIN: { CALL(OBJ_RASTRIGIN); } break; \
    case OBJ_ACKLEY: { CALL(OBJ_ACKLEY); } break;      \
    case OBJ_ROSENBROCK_EX: { CALL(OBJ_ROSENBROCK_EX); } break; \
    case OBJ_BEALE: { CALL(OBJ_BEALE); } break; \
    case OBJ_GOLDSTEIN_PRICE: { CALL(OBJ_GOLDSTEIN_PRICE); } break; \
    case OBJ_THREE_HUMP_CAMEL: { CALL(OBJ_THREE_HUMP_CAMEL); } break; \
    case OBJ_MCCORMICK: { CALL(OBJ_MCCORMICK); } break; \
    case OBJ_SCHAFFER_N2: { CALL(OBJ_SCHAFFER_N2); } break; \
    case OBJ_STYBLINSKI_TANG: { CALL(OBJ_STYBLINSKI_TANG); } break; \
    case OBJ_SHEKEL: { CALL(OBJ_SHEKEL); } break; \
    case OBJ_BOOTH: { CALL(OBJ_BOOTH); } break; \
    case OBJ_BUKIN_N6: { CALL(OBJ_BUKIN_N6); } break; \
    case OBJ_MATYAS: { CALL(OBJ_MATYAS); } break; \
    case OBJ_LEVI_N13: { CALL(OBJ_LEVI_N13); } break; \
    default: return cudaErrorInvalidValue;             \
  }
#endif

template <class T>
cudaError_t de_launch_commit(const DEState &s, int mode, const LaunchGeom &g, cudaStream_t st) {
  de_commit_kernel<T><<<g.reduce_blocks, kBlock, 0, st>>>(s, mode);
  return cudaGetLastError();
}
template <class T>
cudaError_t de_launch_export_best(const DEState &s, void *record, cudaStream_t st) {
  de_export_best_kernel<T><<<1, kBlock, 0, st>>>(s, record);
  return cudaGetLastError();
}
// sign +1: export the k best into rows/scores; sign -1: overwrite the k worst with rows/scores, then re-scan the best
template <class T>
cudaError_t de_launch_migrate(const DEState &s, int sign, unsigned long long k, void *rows, void *scores,
                              const LaunchGeom &g, cudaStream_t st) {
  // contiguous slices of >= 4096 agents, at most kMaxTopkLists of them
  u64 slice = (s.P + kMaxTopkLists - 1) / kMaxTopkLists;
  if (slice < 4096) slice = 4096;
  const u64 blocks = (s.P + slice - 1) / slice;
  if (!s.topk_scratch) return cudaErrorInvalidValue;   // sized by the host side: up to kMaxTopkLists * k pairs
  TopKey *cand = reinterpret_cast<TopKey *>(s.topk_scratch);
  de_topk_partial_kernel<T><<<static_cast<unsigned int>(blocks), kBlock, 0, st>>>(s, sign, u32(k), slice, cand);
  de_topk_final_kernel<T><<<1, kBlock, 0, st>>>(s, sign, u32(k), u32(blocks), cand);
  const unsigned int grid = clamp_grid(k, 4096);
  if (sign > 0) de_gather_kernel<T><<<grid, kBlock, 0, st>>>(s, u32(k), static_cast<T *>(rows), static_cast<T *>(scores));
  else {
    de_scatter_kernel<T><<<grid, kBlock, 0, st>>>(s, u32(k), static_cast<const T *>(rows), static_cast<const T *>(scores));
    de_commit_kernel<T><<<g.reduce_blocks, kBlock, 0, st>>>(s, 2);
  }
  return cudaGetLastError();
}

template <class T>
cudaError_t de_launch_init(const DEState &s, const void *x0_dev, const LaunchGeom &g, cudaStream_t st) {
  const u64 want = (s.P + kWarpsPerBlock - 1) / kWarpsPerBlock;
#define NLS_CALL(O)                                                                                       \
  de_init_kernel<T, O><<<clamp_grid(want, u64(g.sm_count) * blocks_per_sm(de_init_kernel<T, O>)), kBlock, 0, st>>>( \
      s, static_cast<const T *>(x0_dev))
  NLS_OBJ_SWITCH(s.objective, NLS_CALL)
#undef NLS_CALL
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return de_launch_commit<T>(s, 1, g, st);
}

template <class T, int O, int W>
void de_launch_k2_w(const DEState &s, const LaunchGeom &g, cudaStream_t st) {
  const u64 want = (s.P + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const unsigned int grid = clamp_grid(want, u64(g.sm_count) * blocks_per_sm(de_generation_kernel<T, O, W>));
  int shift = 5;                                       // largest tile that still gives >= 8 tiles per warp
  while (shift > 0 && ((s.P + (1ull << shift) - 1) >> shift) < 8ull * grid * kWarpsPerBlock) shift--;
  de_generation_kernel<T, O, W><<<grid, kBlock, 0, st>>>(s, shift);
}
// lanes per agent: the smallest of 4 / 8 / 16 whose single step covers the row, else the whole warp
template <class T, int O>
void de_launch_k2(const DEState &s, const LaunchGeom &g, cudaStream_t st) {
  const u64 vecs = (s.d + Vec<T>::V - 1) / Vec<T>::V;   // 128-bit vectors per row
  if constexpr (closed_form_dim(O) > 0) {               // fixed short vectors: only the 4-lane variant exists
    de_launch_k2_w<T, O, 4>(s, g, st);
    return;
  }
  if (vecs <= 4) de_launch_k2_w<T, O, 4>(s, g, st);
  else if (vecs <= 8) de_launch_k2_w<T, O, 8>(s, g, st);
  else if (vecs <= 16) de_launch_k2_w<T, O, 16>(s, g, st);
  else de_launch_k2_w<T, O, 32>(s, g, st);
}

// one generation: K2, K2r (cooperative), K3
template <class T>
cudaError_t de_launch_generation(const DEState &s, const LaunchGeom &g, cudaStream_t st, cudaEvent_t *ev) {
  cudaError_t e = cudaSuccess;
  if (ev) cudaEventRecord(ev[0], st);
#define NLS_CALL(O)                                                                                                 \
  de_launch_k2<T, O>(s, g, st);                                                                                     \
  e = cudaGetLastError();                                                                                           \
  if (e != cudaSuccess) return e;                                                                                   \
  if (ev) cudaEventRecord(ev[1], st);                                                                               \
  {                                                                                                                 \
    DEState arg = s;                                                                                                \
    void *args[] = {&arg};                                                                                          \
    const unsigned int grid = clamp_grid((s.P + kBlock - 1) / kBlock,                                               \
                                         u64(g.sm_count) * blocks_per_sm(de_repair_kernel<T, O>));                  \
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(de_repair_kernel<T, O>), dim3(grid), dim3(kBlock),     \
                                    args, 0, st);                                                                   \
  }
  NLS_OBJ_SWITCH(s.objective, NLS_CALL)
#undef NLS_CALL
  if (e != cudaSuccess) return e;
  if (ev) cudaEventRecord(ev[2], st);
  e = de_launch_commit<T>(s, 0, g, st);
  if (ev) cudaEventRecord(ev[3], st);
  return e;
}


template <class T>
cudaError_t de_launch_gather_rows(const DEState &s, unsigned long long first, unsigned long long count, void *out,
                                  cudaStream_t st) {
  de_gather_rows_kernel<T><<<clamp_grid(count, 1u << 16), kBlock, 0, st>>>(s, first, count, static_cast<T *>(out));
  return cudaGetLastError();
}

#define NLS_DEFINE_DE_OPS(T, NAME)                                                                        \
  const DEOps *NAME() {                                                                                   \
    static const DEOps ops = {de_launch_init<T>, de_launch_generation<T>, de_launch_export_best<T>,       \
                              de_launch_migrate<T>, de_launch_gather_rows<T>};                            \
    return &ops;                                                                                          \
  }

}  // namespace nls
