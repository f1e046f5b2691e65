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
//   K2r de_repair_kernel     : cooperative kernel, a fixed-point iteration over the "donor r < i" relation: in
//                              iteration k an agent is re-evaluated iff one of its lower donors changed its visible
//                              state (accept flag / accepted row) in iteration k - 1, against the rows as they stand.
//                              Two grid barriers per iteration (scan, re-evaluate); it stops when an iteration
//                              changed nothing, and the state then equals the sequential loop's (see below).
//                              Returns at once when the speculative pass accepted nothing.
//   K3  de_commit_kernel     : commits accepted trials (score, row-location bit), then the population reduction:
//                              min-loc with the reference's tie rule (nlsolver.h:2432-2437), the std_err statistic
//                              (nlsolver.h:2037-2052) and the stop test (nlsolver.h:2439-2447), last block finalises.
#pragma once
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <math_constants.h>

#include "de_bulk.cuh"
#include "objectives.cuh"
#include "reduce.cuh"
#include "launch.h"
#include "state.h"

namespace nls {
namespace cg = cooperative_groups;

// fin[i] of an agent whose visible state never changed in this generation (see de_repair_kernel)
constexpr uint16_t kNeverChanged = 0xFFFFu;


// ------------------------------------------------------------------------------------------------ row pass
// One sweep over the d coordinates of agent i's trial (propose_new_agent, nlsolver.h:2357-2375):
//   trial[j] = mut ? A[r1][j] + F * (A[r2][j] - A[r3][j]) : A[r0][j],   mut = (draw_j < CR) || (j == dim)
// EVAL accumulates the objective; WRITE stores the trial into `dst`.  W lanes cooperate on the agent (`lane` is the
// index inside that group); each lane owns V consecutive coordinates per step (one 128-bit load per row) and U steps
// are issued back to back, so 4 * U loads per lane are in flight.  Lanes past the end of the row read coordinate 0
// instead (always valid, one cached line) and are masked out downstream.
//   S = accumulator slots per lane (objectives.cuh): a group of W < 32 lanes that sweeps a row LONGER than one step
// (d > W * V) keeps S = 32 / W canonical accumulators per lane, step k feeding slot k.  The whole row must then fit
// the one unrolled iteration (d <= U * W * V, U <= S) so that the slot is a compile-time constant.
//   SKIP_BASE (long rows): the base row is only loaded where a coordinate of the lane keeps it, which needs the draws
// before the loads.  Short rows load all four rows FIRST and draw while the loads are in flight: a skipped 16-byte
// piece saves no DRAM traffic there (128-byte lines), and the ~25 integer instructions per draw are the only
// independent work a lane has to cover the load latency with.
//   SHARED_BASE: the base row is the single best row (best recombination, speculative pass): read-only for the launch
// and shared by every agent, so it goes through L1 instead of being re-fetched from L2 per agent.
// Tuning (B200, fp64 d=1000 Rastrigin, P=2^20, K2 time): U=2 / 2 blocks per SM 6.29 ms; U=2 / 3 blocks 5.68 ms;
// U=1 / 4 blocks (64 registers, 32 warps per SM, no spills) 5.63 ms — occupancy beats per-warp unrolling for long rows.
//   KEEP (rows that one trip covers): the trial's coordinates are also handed back in `keep[u][q]`, so that an accepted
// trial is stored from registers instead of being rebuilt by a second sweep over the four rows.
template <class T, int OBJ, bool EVAL, bool WRITE, int W = 32, bool SHARED_BASE = false, int U = 1, int S = 1,
          bool SKIP_BASE = true, bool KEEP = false>
__device__ __forceinline__ T de_sweep(const DEState &s, const T *p0, const T *p1, const T *p2, const T *p3, T *dst,
                                      u64 sbase, u32 dim, u64 i, int lane, T (*keep)[Vec<T>::V] = nullptr) {
  constexpr int V = Vec<T>::V;
  constexpr u32 kStride = W * V;                       // coordinates per group step
  static_assert(S == 1 || U <= S, "with accumulator slots the row is one unrolled iteration: step u feeds slot u");
  typedef Ar<T> A;
  const u32 d = static_cast<u32>(s.d);
  const T F = static_cast<T>(s.F);
  const u64 cr_le = s.cr_le;
  const bool cr_any = !s.cr_none;
  const u32 n_steps = (d + kStride - 1) / kStride;
  Objective<T, OBJ, W, S> obj;
  if (EVAL) obj.begin(lane, d);
  u32 j0 = lane * V;                                   // first coordinate of this lane in the current step
  u64 st = sbase + kGolden * j0;                       // draw-stream state of coordinate j0
  for (u32 step = 0; step < n_steps; step += U) {
    T x1[U][V], x2[U][V], x3[U][V], x0[U][V];
    bool mut[U][V];
    if (!SKIP_BASE) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const u32 jj = j0 + u * kStride;
        const u32 jl = jj < d ? jj : 0u;
        ld_row(p1 + jl, x1[u]); ld_row(p2 + jl, x2[u]); ld_row(p3 + jl, x3[u]);
        if (SHARED_BASE) ld_row_shared(p0 + jl, x0[u]);
        else ld_row(p0 + jl, x0[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const u32 jj = j0 + u * kStride;
      bool all_mut = true;
#pragma unroll
      for (int q = 0; q < V; q++) {
        mut[u][q] = (cr_any && mix64(st + kGolden * (u * kStride + q)) <= cr_le) || (jj + q == dim);
        all_mut &= mut[u][q];
      }
      if (SKIP_BASE) {
        const u32 jl = jj < d ? jj : 0u;
        // (the predicated base-row load goes FIRST: issued after the donor loads, ptxas has been seen to sink it below
        //  the first use of the donors' data to save a register pair, which exposes a second DRAM latency per step)
        if (!all_mut) {                                // the base row is only touched where a coordinate keeps it
          if (SHARED_BASE) ld_row_shared(p0 + jl, x0[u]);
          else ld_row(p0 + jl, x0[u]);
        }
        ld_row(p1 + jl, x1[u]); ld_row(p2 + jl, x2[u]); ld_row(p3 + jl, x3[u]);
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
      if (KEEP) {
#pragma unroll
        for (int q = 0; q < V; q++) keep[u][q] = t[q];
      }
      if (EVAL) {
        obj.step(t, jj, d, lane, S == 1 ? 0 : u);
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
// What the lane-parallel prologue hands to the cooperative part, one entry per agent of the tile (shared memory).
struct __align__(16) DETileEntry {
  unsigned long long key;      // draw stream of (generation, agent)
  const void *p0, *p1, *p2, *p3;   // the rows of ids[0..3] this evaluation reads
  double score;                // scores[i] (T widened)
  unsigned int agent;          // local agent index i
  unsigned int dim;            // forced crossover coordinate
  unsigned int rej;            // rejected index proposals
  unsigned int flags;          // bit 0: where[i] (an accepted trial goes to the OTHER buffer); bit 1: acc[i] before
};
static_assert(sizeof(DETileEntry) == 64, "four 128-bit shared-memory loads per entry");

// The lane-parallel prologue of a tile of the speculative pass: lane l prepares agent first + l (generate_indices,
// nlsolver.h:2331-2355, the forced coordinate, the addresses of the four PRE-generation rows), stores the entry in
// shared memory and the decisions in global memory.
template <class T>
__device__ __forceinline__ void de_tile_prologue(const DEState &s, DETileEntry *tile_entries, u64 gen_key, u64 best_id,
                                                 bool random_mode, u64 first, int tile_size, int lane) {
  const u64 mine = first + lane;
  if (lane < tile_size && mine < s.P) {
    const char *buf0 = static_cast<const char *>(s.buf[0]), *buf1 = static_cast<const char *>(s.buf[1]);
    const u64 row_bytes = s.stride * sizeof(T);
    DETileEntry e;
    e.key = tape_key(gen_key, s.offset + mine);
    const u64 fixed = random_mode ? mine : best_id;
    u64 r1, r2, r3;
    de_select_donors<T>(e.key, s.P, fixed, r1, r2, r3, e.rej);
    e.dim = static_cast<u32>(index_from<T>(tape_draw(e.key, 3 + e.rej), s.d));
    const u32 w0 = s.where[fixed], w1 = s.where[r1], w2 = s.where[r2], w3 = s.where[r3];
    e.p0 = (w0 ? buf1 : buf0) + fixed * row_bytes;
    e.p1 = (w1 ? buf1 : buf0) + r1 * row_bytes;
    e.p2 = (w2 ? buf1 : buf0) + r2 * row_bytes;
    e.p3 = (w3 ? buf1 : buf0) + r3 * row_bytes;
    e.score = static_cast<double>(static_cast<const T *>(s.score)[mine]);
    e.agent = u32(mine);
    e.flags = s.where[mine];
    tile_entries[lane] = e;
    s.dec[mine] = make_uint4(u32(r1), u32(r2), u32(r3), e.dim);
    s.rej[mine] = e.rej;
  }
}

// The prologue of a tile of the repair: lane l prepares list entry first + l from the recorded decisions, with the
// rows RESOLVED — a lower donor (r < i) whose trial currently counts as accepted contributes its new row.
template <class T>
__device__ __forceinline__ void de_tile_prologue_repair(const DEState &s, DETileEntry *tile_entries, u64 gen_key,
                                                        u64 best_id, bool random_mode, u64 first, u32 n_list,
                                                        int tile_size, int lane) {
  const u64 slot = first + lane;
  if (lane < tile_size && slot < n_list) {
    const char *buf0 = static_cast<const char *>(s.buf[0]), *buf1 = static_cast<const char *>(s.buf[1]);
    const u64 row_bytes = s.stride * sizeof(T);
    const u64 mine = s.list[slot];
    const uint4 dc = s.dec[mine];
    DETileEntry e;
    e.key = tape_key(gen_key, s.offset + mine);
    e.rej = s.rej[mine];
    e.dim = dc.w;
    const u64 r[4] = {random_mode ? mine : best_id, dc.x, dc.y, dc.z};
    const void *p[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      u32 w = s.where[r[q]];
      if (r[q] < mine && s.acc[r[q]]) w ^= 1u;
      p[q] = (w ? buf1 : buf0) + r[q] * row_bytes;
    }
    e.p0 = p[0]; e.p1 = p[1]; e.p2 = p[2]; e.p3 = p[3];
    e.score = static_cast<double>(static_cast<const T *>(s.score)[mine]);
    e.agent = u32(mine);
    e.flags = u32(s.where[mine]) | (u32(s.acc[mine]) << 1);
    tile_entries[lane] = e;
  }
}

// The cooperative part of a tile: the prepared agents, 32 / W at a time, W lanes streaming the rows of one agent
// (de_sweep: trial, objective, greedy selection — loop body nlsolver.h:2459-2471); an accepted trial (a few % of the
// agents) is re-streamed from the L2-warm rows and written to the agent's row in the other buffer.  The first lane of a
// group stores the trial score and accept flag and hands the outcome to `done(entry, ok)`.
template <class T, int OBJ, int W, int U, int S, bool SKIP_BASE, class Done>
__device__ __forceinline__ void de_tile_body(const DEState &s, const DETileEntry *tile_entries, int n_here,
                                             bool shared_base, int lane, Done done) {
  constexpr int G = 32 / W;                              // agents streamed at a time
  const int grp = lane / W, sub = lane % W;
  for (int a0 = 0; a0 < n_here; a0 += G) {
    const bool active = a0 + grp < n_here;               // idle groups redo the first agent, nothing is stored
    const DETileEntry e = tile_entries[active ? a0 + grp : a0];   // broadcast read inside the group
    const u64 i = e.agent;
    const T *p0 = static_cast<const T *>(e.p0), *p1 = static_cast<const T *>(e.p1);
    const T *p2 = static_cast<const T *>(e.p2), *p3 = static_cast<const T *>(e.p3);
    const u64 sbase = tape_state(e.key, 4 + e.rej);     // draws 0..2+rej: indices, 3+rej: dim, then one per coordinate
    // (warp-uniform branch: in the speculative pass of best mode every base row is the pre-generation best row)
    // short rows (the kernels without the base-row skip): one trip of U steps covers the row, the trial stays in registers
    constexpr bool kKeep = !SKIP_BASE;
    constexpr int V = Vec<T>::V;
    T keep[U][V];
    const T raw = shared_base
                      ? de_sweep<T, OBJ, true, false, W, true, U, S, SKIP_BASE, kKeep>(s, p0, p1, p2, p3, nullptr, sbase, e.dim, i, sub, keep)
                      : de_sweep<T, OBJ, true, false, W, false, U, S, SKIP_BASE, kKeep>(s, p0, p1, p2, p3, nullptr, sbase, e.dim, i, sub, keep);
    const T sc = Ar<T>::mul(static_cast<T>(s.fm), raw);
    const bool ok = active && sc < static_cast<T>(e.score);   // strict <, NaN never accepted (nlsolver.h:2466)
    if (ok) {
      T *dst = static_cast<T *>(s.buf[(e.flags & 1u) ^ 1u]) + i * s.stride;
      if (kKeep && static_cast<u32>(s.d) <= u32(U * W * V)) {
#pragma unroll
        for (int u = 0; u < U; u++) {
          const u32 jj = u32(sub * V + u * W * V);
          if (jj < static_cast<u32>(s.d)) st_row(dst + jj, keep[u]);
        }
      } else {
        de_sweep<T, OBJ, false, true, W, false, U, S, SKIP_BASE>(s, p0, p1, p2, p3, dst, sbase, e.dim, i, sub);
      }
    }
    if (active && sub == 0) {
      static_cast<T *>(s.tscore)[i] = sc;
      s.acc[i] = ok;
      done(e, ok);
    }
  }
}

// A warp owns a TILE of up to 32 consecutive agents:
//   prologue  — one agent per lane: stream key, donor selection, forced coordinate, row addresses, current score;
//               the per-agent scalar work and its dependent loads run 32 agents wide instead of 32 times redundantly;
//   body      — the agents of the tile, 32 / W at a time, W lanes streaming the rows of one agent (de_sweep); the
//               first lane of a group stores the agent's trial score / accept flag / repair stamp.
// The tile size (2^tile_shift <= 32 agents) is chosen by the launcher so that every warp gets many tiles: with 32-agent
// tiles a population of 2^18 would give 1.7 tiles per resident warp and a 15 % tail.
// Lane-group shapes (de_launch_k2): W = 4 / 8 lanes with one step when it covers the row; W = 8 / 16 / 32 lanes with
// U = 2 steps and S = 32 / W accumulator slots for rows of up to 16 / 32 / 64 vectors (fp32 d = 64, fp64 d = 64, ...) —
// twice the loads in flight per lane and the per-agent overhead spread over twice the coordinates; W = 32, U = 1 with
// the base-row skip for long rows.
#ifndef NLS_DE_BULK_STAGES
#define NLS_DE_BULK_STAGES 2
#endif
#ifndef NLS_DE_BULK_STEPS
#define NLS_DE_BULK_STEPS 2
#endif
#ifndef NLS_DE_BULK_BLOCKS
#define NLS_DE_BULK_BLOCKS 2
#endif
// Sweep steps in flight per lane when the repair re-evaluates LONG rows.  A listed agent is four random rows read by
// one warp, and a warp gets a handful of agents per iteration, so what an iteration costs is the latency of a row, i.e.
// (steps per row) / (steps in flight) DRAM round trips.  K2r at the config-2 shape with 2-3 % of the trials accepted:
// 0.52 / 0.42 / 0.36 ms with one step in flight, 0.37 / 0.29 / 0.26 with two, 0.34 / 0.26 / 0.23 with four (128
// registers, two blocks per SM — occupancy is not what this pass needs).
#ifndef NLS_DE_REPAIR_U
#define NLS_DE_REPAIR_U 4
#endif
// Tile size (log2, <= 5) for a pass over n agents by n_warps warps that stream G agents at a time.  A tile costs one
// prologue — a chain of dependent round trips: draws -> donor ids -> row locations -> scores — plus ceil(tile / G) body
// trips of one row latency each, and the pass lasts as long as the warp with the most tiles.  `prologue` is the cost of
// a prologue in body trips (about 2 for short rows, a fraction for long ones).  Large populations end up with the
// largest tile (fewest prologues; every warp still gets many tiles), small ones with few, larger tiles per warp instead
// of many one-agent tiles whose prologue latency nothing hides: K2 at P = 2^16, d = 64 is 9 trips of (prologue + one row)
// per warp with 2-agent tiles, and 3 x (prologue + 4 rows) with 8-agent tiles.
__host__ __device__ inline int de_tile_shift(unsigned long long n, unsigned long long n_warps, int G, double prologue) {
  int best = 0;
  double best_cost = 1e300;
  for (int shift = 0; shift <= 5; shift++) {
    const unsigned long long tiles = (n + (1ull << shift) - 1) >> shift;
    const unsigned long long per_warp = (tiles + n_warps - 1) / n_warps;
    const double cost = double(per_warp) * (prologue + double(((1u << shift) + G - 1) / G));
    if (cost <= best_cost) { best_cost = cost; best = shift; }     // ties: the larger tile
  }
  return best;
}

template <int U> struct DEBlocksPerSM { static constexpr int value = U >= 4 ? 2 : (U >= 2 ? 3 : 4); };
// barriers of the multi-phase passes: the whole grid (cooperative launch) or one thread-block cluster (the one-launch
// path for small populations, where a generation is a handful of microseconds and the barrier latency is what counts)
struct GridSync {
  cg::grid_group g;
  __device__ __forceinline__ void operator()() { g.sync(); }
};
struct ClusterSync {
  __device__ __forceinline__ void operator()() { cg::this_cluster().sync(); }
};

template <class T, int OBJ, int W, int U, int S, bool SKIP_BASE>
__device__ __forceinline__ void de_generation_pass(const DEState &s, DETileEntry *tile_entries, int tile_shift) {
  DECtrl *ctrl = s.ctrl;
  const int lane = threadIdx.x & 31;
  const u64 warp = (u64(blockIdx.x) * kBlock + threadIdx.x) >> 5, n_warps = (u64(gridDim.x) * kBlock) >> 5;
  const u64 gen_key = tape_gen_key(s.seed, ctrl->iter + 1), best_id = ctrl->best_id;
  const bool random_mode = s.strategy != 0;             // NLS_DE_RANDOM = 1: ids[0] = i; best: ids[0] = best_id
  const u64 P = s.P, tile_size = 1ull << tile_shift, n_tiles = (P + tile_size - 1) >> tile_shift;
  u32 n_ok = 0;
  for (u64 tile = warp; tile < n_tiles; tile += n_warps) {
    const u64 first = tile << tile_shift;
    de_tile_prologue<T>(s, tile_entries, gen_key, best_id, random_mode, first, int(tile_size), lane);
    __syncwarp();
    const int n_here = (P - first) < tile_size ? int(P - first) : int(tile_size);
    de_tile_body<T, OBJ, W, U, S, SKIP_BASE>(s, tile_entries, n_here, !random_mode, lane, [&](const DETileEntry &e, bool ok) {
      s.fin[e.agent] = ok ? uint16_t(0) : kNeverChanged;     // pass 0 of the repair's fixed-point iteration
      if (ok && s.coarse != nullptr) atomicOr(s.coarse + (e.agent >> 10), 1u << ((e.agent >> 5) & 31u));   // bitmap 0
      n_ok += ok;
    });
    __syncwarp();                                          // tile_entries is rewritten by the next tile's prologue
  }
  n_ok = __reduce_add_sync(kFull, n_ok);
  if (lane == 0 && n_ok) atomicAdd(&ctrl->spec_accepted, n_ok);
}
template <class T, int OBJ, int W, int U, int S, bool SKIP_BASE>
__global__ void __launch_bounds__(kBlock, DEBlocksPerSM<U>::value) de_generation_kernel(DEState s, int tile_shift) {
  __shared__ DETileEntry tile_mem[kWarpsPerBlock][32];
  if (s.ctrl->stop) return;
  de_generation_pass<T, OBJ, W, U, S, SKIP_BASE>(s, tile_mem[threadIdx.x >> 5], tile_shift);
}

// ------------------------------------------------------------------------------------------------ K2, TMA-staged
// The same generation pass for long rows, the four rows of every agent staged through shared memory by bulk copies
// (de_bulk.cuh).  One warp per agent; kSteps sweep steps per chunk; a ring of kStages chunks per warp that runs across
// the agents of the tile.  The base row is staged whole like the donors (in best mode it is the single best row and
// comes out of L2; in random mode that costs the ~3 % of traffic the LDG version saves by skipping fully mutated
// 128-byte lines).  An accepted trial is materialised by the ordinary second sweep.
template <class T, int OBJ, int kStages, int kSteps>
__global__ void __launch_bounds__(kBlock, NLS_DE_BULK_BLOCKS) de_generation_bulk_kernel(DEState s, int tile_shift) {
  constexpr int V = Vec<T>::V;
  constexpr u32 kStride = 32 * V;                        // coordinates per sweep step
  constexpr u32 kChunkBytes = 512u * kSteps;             // per row and stage
  constexpr u32 kStageBytes = 4u * kChunkBytes;
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  __shared__ DETileEntry tile_mem[kWarpsPerBlock][32];
  __shared__ __align__(8) unsigned long long bars[kWarpsPerBlock][kStages];
  typedef Ar<T> A;
  DECtrl *ctrl = s.ctrl;
  if (ctrl->stop) return;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  DETileEntry *tile_entries = tile_mem[wib];
  unsigned long long *bar = bars[wib];
  unsigned char *ring = bulk_smem + size_t(wib) * kStages * kStageBytes;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kStages; k++) mbar_init(bar + k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const u64 warp = (u64(blockIdx.x) * kBlock + threadIdx.x) >> 5, n_warps = (u64(gridDim.x) * kBlock) >> 5;
  const u64 gen_key = tape_gen_key(s.seed, ctrl->iter + 1), best_id = ctrl->best_id;
  const bool random_mode = s.strategy != 0;
  const u64 P = s.P, tile_size = 1ull << tile_shift, n_tiles = (P + tile_size - 1) >> tile_shift;
  const u32 d = static_cast<u32>(s.d);
  const u32 row_bytes = static_cast<u32>(s.stride * sizeof(T));
  const u32 n_chunks = (row_bytes + kChunkBytes - 1) / kChunkBytes;
  const T F = static_cast<T>(s.F);
  const u64 cr_le = s.cr_le;
  const bool cr_any = !s.cr_none;
  u32 phases = 0;                                        // parity each stage's barrier completes with next
  u32 n_ok = 0;
  for (u64 tile = warp; tile < n_tiles; tile += n_warps) {
    const u64 first = tile << tile_shift;
    de_tile_prologue<T>(s, tile_entries, gen_key, best_id, random_mode, first, int(tile_size), lane);
    __syncwarp();
    const int n_here = (P - first) < tile_size ? int(P - first) : int(tile_size);
    // producer cursor: (agent of the tile, chunk of its rows, stage); the ring is empty at a tile boundary
    int pa = 0;
    u32 pc = 0, pstage = 0, cstage = 0;
    auto produce = [&]() {
      if (pa >= n_here) return;
      if (lane == 0) {
        const DETileEntry &pe = tile_entries[pa];
        const u32 off = pc * kChunkBytes;
        const u32 bytes = row_bytes - off < kChunkBytes ? row_bytes - off : kChunkBytes;
        unsigned char *st = ring + pstage * kStageBytes;
        fence_proxy_async();
        mbar_arrive_expect_tx(bar + pstage, 4u * bytes);
        bulk_load(st, static_cast<const char *>(pe.p0) + off, bytes, bar + pstage);
        bulk_load(st + kChunkBytes, static_cast<const char *>(pe.p1) + off, bytes, bar + pstage);
        bulk_load(st + 2 * kChunkBytes, static_cast<const char *>(pe.p2) + off, bytes, bar + pstage);
        bulk_load(st + 3 * kChunkBytes, static_cast<const char *>(pe.p3) + off, bytes, bar + pstage);
      }
      pstage = pstage + 1 == kStages ? 0 : pstage + 1;
      if (++pc == n_chunks) { pc = 0; pa++; }
    };
#pragma unroll
    for (int k = 0; k < kStages - 1; k++) produce();
    for (int a = 0; a < n_here; a++) {
      const DETileEntry e = tile_entries[a];
      const u64 i = e.agent;
      Objective<T, OBJ, 32, 1> obj;
      obj.begin(lane, d);
      u32 j0 = lane * V;
      u64 st = tape_state(e.key, 4 + e.rej) + kGolden * j0;
      for (u32 c = 0; c < n_chunks; c++) {
        produce();                                         // refills the stage that was consumed one iteration ago
        mbar_wait(bar + cstage, (phases >> cstage) & 1u);
        phases ^= 1u << cstage;
        const unsigned char *sb = ring + cstage * kStageBytes + lane * 16;
#pragma unroll
        for (int u = 0; u < kSteps; u++) {
          const u32 jj = j0 + u * kStride;
          T x0[V], x1[V], x2[V], x3[V], t[V];
          lds_row(sb + u * 512, x0);
          lds_row(sb + kChunkBytes + u * 512, x1);
          lds_row(sb + 2 * kChunkBytes + u * 512, x2);
          lds_row(sb + 3 * kChunkBytes + u * 512, x3);
#pragma unroll
          for (int q = 0; q < V; q++) {
            const bool mut = (cr_any && mix64(st + kGolden * (u * kStride + q)) <= cr_le) || (jj + q == e.dim);
            t[q] = mut ? A::add(x1[q], A::mul(F, A::sub(x2[q], x3[q]))) : x0[q];
            if (s.masks != nullptr && jj + q < d) s.masks[i * d + jj + q] = mut;
          }
          obj.step(t, jj, d, lane);
        }
        j0 += kSteps * kStride;
        st += kGolden * (kSteps * kStride);
        __syncwarp();                                      // every lane has read the stage before it is refilled
        cstage = cstage + 1 == kStages ? 0 : cstage + 1;
      }
      const T sc = A::mul(static_cast<T>(s.fm), obj.finish(d));
      const bool ok = sc < static_cast<T>(e.score);        // strict <, NaN never accepted (nlsolver.h:2466)
      if (ok)
        de_sweep<T, OBJ, false, true, 32, false, 1, 1, true>(s, static_cast<const T *>(e.p0), static_cast<const T *>(e.p1),
                                                             static_cast<const T *>(e.p2), static_cast<const T *>(e.p3),
                                                             static_cast<T *>(s.buf[(e.flags & 1u) ^ 1u]) + i * s.stride,
                                                             tape_state(e.key, 4 + e.rej), e.dim, i, lane);
      if (lane == 0) {
        static_cast<T *>(s.tscore)[i] = sc;
        s.acc[i] = ok;
        s.fin[i] = ok ? uint16_t(0) : kNeverChanged;
        if (ok && s.coarse != nullptr) atomicOr(s.coarse + (i >> 10), 1u << ((u32(i) >> 5) & 31u));   // bitmap 0
        n_ok += ok;
      }
    }
    __syncwarp();
  }
  if (lane == 0 && n_ok) atomicAdd(&ctrl->spec_accepted, n_ok);
}

// ------------------------------------------------------------------------------------------------ K2r repair
// Exact in-place semantics as a fixed point.  After the speculative pass (iteration 0) fin[i] holds the last iteration
// in which agent i's VISIBLE state — its accept flag and, if accepted, its new row — may have changed (kNeverChanged:
// the agent was never accepted).  Iteration k >= 1 re-evaluates every agent that has a lower donor r (r < i, the only
// donors whose new state the sequential loop lets i see, nlsolver.h:2466-2471) with fin[r] == k - 1, against the rows
// as they stand (RESOLVED lookup), and stamps fin[i] = k if it was or now is accepted.  The iteration stops when nothing
// was stamped.
//   Why this equals the sequential loop: let L be the last iteration in which a lower donor of i was stamped.  In
// iteration L + 1 agent i is re-evaluated (fin[r] == L is stable: a re-stamp would be a later change), and no lower
// donor changes visibly while it reads them, so it sees exactly the final states of everything below it; by induction
// over i those are the sequential loop's states.  Evaluations made EARLIER than that may have read rows in flux (a
// donor re-evaluated in the same iteration) — their results are garbage by design and always overwritten, because the
// donor that was in flux stamps itself and so re-triggers its dependents.  With no lower donor ever stamped, the
// speculative result stands.  Cost: two grid barriers per iteration; the hit sets shrink geometrically (about 1.5 a P,
// then 3 a of that, ...) and the number of iterations is bounded by the depth of the "donor r < i" DAG (~17 at 2^20).
// Each iteration is two phases with a grid barrier after each: SCAN (one thread per agent: which agents have a lower
// donor stamped in the previous iteration?  hits go to a list through warp-aggregated atomics) and RE-EVALUATE (the
// listed agents spread evenly over all lane groups of the grid — a tile-local loop left most warps idle behind the few
// that drew several hits: 9.6 % issue utilisation at a 5 % hit rate).
template <class T, int OBJ, int W, int U, int S, bool SKIP_BASE, class Sync>
__device__ __forceinline__ void de_repair_pass(const DEState &s, DETileEntry *tile_entries, Sync &sync) {
  DECtrl *ctrl = s.ctrl;
  // If the speculative pass accepted nothing, no row changed and every speculative result already equals the
  // sequential one (induction over the agent index) — nothing to repair.
  if (*reinterpret_cast<volatile unsigned int *>(&ctrl->spec_accepted) == 0) return;
  const int lane = threadIdx.x & 31;
  const u64 tid = u64(blockIdx.x) * kBlock + threadIdx.x, n_threads = u64(gridDim.x) * kBlock;
  const u64 warp = tid >> 5, n_warps = n_threads >> 5;
  const u64 gen_key = tape_gen_key(s.seed, ctrl->iter + 1), best_id = ctrl->best_id;
  const bool best_mode = s.strategy == 0;
  const u64 P = s.P;
  u32 k = 1;
  for (;; k++) {
    const u32 want = k - 1, cur = k % 3u;
    // ---- scan: kScan agents per lane and trip.  All decision loads of the trip are issued first, then all stamp loads
    // (two dependent L2 round trips per trip instead of two per agent), and the warp reserves list space with ONE atomic
    // per trip.  A scan is bound by the random 32-byte sectors of the stamp lookups (1.5 per agent), so the
    // lookups are filtered through a coarse bitmap the previous re-evaluation filled (bitmap 0: the speculative pass) —
    // one bit per 32 agents, small enough to live in L1: wherever few agents were stamped — the late iterations, or every
    // iteration of a generation that accepts little — almost every lookup stops there and the scan streams the decisions
    // only.  (Ordinary loads: the kernel boundary / grid / cluster barrier that separates the bitmap's writers from these
    // readers makes them visible.)  K3 clears bitmap 0 for the next generation.
    constexpr int kScan = 4;
    const u32 *cb_prev = s.coarse != nullptr ? s.coarse + u64((k - 1) % 3u) * s.coarse_words : nullptr;
    u32 *cb_cur = s.coarse != nullptr ? s.coarse + u64(cur) * s.coarse_words : nullptr;
    // the bitmap this iteration's re-evaluation fills: its last readers were the scan two iterations ago
    if (cb_cur != nullptr)
      for (u64 w = tid; w < s.coarse_words; w += n_threads) cb_cur[w] = 0u;
    for (u64 base = warp * (32u * kScan); base < P; base += n_threads * kScan) {
      uint4 dc[kScan];
      u32 f[kScan][4];
#pragma unroll
      for (int u = 0; u < kScan; u++) {
        const u64 mine = base + u64(u) * 32u + lane;
        dc[u] = mine < P ? s.dec[mine] : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);
      }
      bool look[kScan][4];
#pragma unroll
      for (int u = 0; u < kScan; u++) {
        const u64 mine = base + u64(u) * 32u + lane;
        const u32 don[4] = {dc[u].x, dc[u].y, dc[u].z, u32(best_id)};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          // (predicated loads: the stamp of a donor that is not lower is never needed)
          look[u][q] = mine < P && don[q] < mine && (q < 3 || best_mode);
          if (cb_prev != nullptr && look[u][q]) look[u][q] = (cb_prev[don[q] >> 10] >> ((don[q] >> 5) & 31u)) & 1u;
        }
      }
#pragma unroll
      for (int u = 0; u < kScan; u++) {
        const u32 don[4] = {dc[u].x, dc[u].y, dc[u].z, u32(best_id)};
#pragma unroll
        for (int q = 0; q < 4; q++) f[u][q] = look[u][q] ? u32(__ldcg(s.fin + don[q])) : 0x10000u;
      }
      u32 vote[kScan], n_hits = 0;
#pragma unroll
      for (int u = 0; u < kScan; u++) {
        const bool hit = f[u][0] == want || f[u][1] == want || f[u][2] == want || f[u][3] == want;
        vote[u] = __ballot_sync(kFull, hit);
        n_hits += __popc(vote[u]);
      }
      if (n_hits) {                                         // warp-uniform
        u32 slot = 0;
        if (lane == 0) slot = atomicAdd(&ctrl->list_count[cur], n_hits);
        slot = __shfl_sync(kFull, slot, 0);
#pragma unroll
        for (int u = 0; u < kScan; u++) {
          if (vote[u] & (1u << lane)) s.list[slot + __popc(vote[u] & ((1u << lane) - 1u))] = u32(base + u64(u) * 32u + lane);
          slot += __popc(vote[u]);
        }
      }
    }
    sync();
    const u32 n_list = *reinterpret_cast<volatile unsigned int *>(&ctrl->list_count[cur]);
    // recycle the counter slots iteration k + 2 will use: their last readers passed an earlier barrier, their next
    // writers start after the barrier that ends this iteration
    if (tid == 0) { ctrl->list_count[(k + 2u) % 3u] = 0; ctrl->changed[(k + 2u) % 3u] = 0; }
    if (n_list == 0) break;
    // ---- re-evaluate: tiles of the list, sized so that every warp gets work (a short list is spread thin)
    const int shift = de_tile_shift(n_list, n_warps, 32 / W, W == 32 ? 0.5 : 2.0);
    const u64 tile_size = 1ull << shift, n_tiles = (u64(n_list) + tile_size - 1) >> shift;
    u32 n_changed = 0;
    for (u64 tile = warp; tile < n_tiles; tile += n_warps) {
      const u64 first = tile << shift;
      de_tile_prologue_repair<T>(s, tile_entries, gen_key, best_id, !best_mode, first, n_list, int(tile_size), lane);
      __syncwarp();
      const int n_here = (n_list - first) < tile_size ? int(n_list - first) : int(tile_size);
      de_tile_body<T, OBJ, W, U, S, SKIP_BASE>(s, tile_entries, n_here, false, lane, [&](const DETileEntry &e, bool ok) {
        if ((e.flags >> 1) | u32(ok)) {
          s.fin[e.agent] = uint16_t(k);
          if (cb_cur != nullptr) atomicOr(cb_cur + (e.agent >> 10), 1u << ((e.agent >> 5) & 31u));
          n_changed++;
        }
      });
      __syncwarp();
    }
    n_changed = __reduce_add_sync(kFull, n_changed);
    if (lane == 0 && n_changed) atomicAdd(&ctrl->changed[cur], n_changed);
    if (tid == 0) ctrl->reruns += n_list;
    sync();
    if (*reinterpret_cast<volatile unsigned int *>(&ctrl->changed[cur]) == 0) break;
    if (k >= 65000u) { if (tid == 0) ctrl->error = 1; break; }
  }
  if (tid == 0) ctrl->rounds += k;
}

template <class T>
__device__ __forceinline__ void de_commit_pass(const DEState &s, int mode);

// commit != 0: K3 runs behind the repair in the same launch (small populations, where a generation is bound by launch
// latency; the launcher only asks for it when this grid is the commit kernel's grid, so partials and results are the
// same).  Every way out of the repair pass is right behind a grid barrier — or nothing was written at all — so the
// commit pass may read what the repair wrote.
template <class T, int OBJ, int W, int U, int S, bool SKIP_BASE>
__global__ void __launch_bounds__(kBlock, DEBlocksPerSM<U>::value) de_repair_kernel(DEState s, int commit) {
  __shared__ DETileEntry tile_mem[kWarpsPerBlock][32];
  if (s.ctrl->stop) return;                              // uniform over the grid: only K3 changes it
  GridSync sync{cg::this_grid()};
  de_repair_pass<T, OBJ, W, U, S, SKIP_BASE>(s, tile_mem[threadIdx.x >> 5], sync);
  if (commit) de_commit_pass<T>(s, 0);
}

// ------------------------------------------------------------------------------------------------ K3 commit + reduce
// mode 0: commit the generation in flight; 1: first scan after init; 2: re-scan after island migration (best only)
template <class T>
__device__ __forceinline__ void de_commit_pass(const DEState &s, int mode) {
  const bool initial = mode != 0;
  DECtrl *ctrl = s.ctrl;
  T *score = static_cast<T *>(s.score);
  const T *tscore = static_cast<const T *>(s.tscore);
  u32 n_acc = 0;
  // an accepted trial commits as: score = trial score, row location flipped (the row itself is already in place)
  auto item = [&](u64 i, double &for_min, double &for_moments) {
    const bool a = !initial && s.acc[i];
    for_min = for_moments = static_cast<double>(a ? tscore[i] : score[i]);
  };
  auto store = [&](u64 i) {
    if (!initial && s.acc[i]) { score[i] = tscore[i]; s.where[i] ^= 1u; n_acc++; }
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
    ctrl->changed[0] = ctrl->changed[1] = ctrl->changed[2] = 0;
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
  // the repair's coarse bitmap 0 (agents accepted by the speculative pass) is rebuilt by the next generation's K2
  if (s.coarse != nullptr)
    for (u64 w = u64(blockIdx.x) * kBlock + threadIdx.x; w < s.coarse_words; w += u64(gridDim.x) * kBlock) s.coarse[w] = 0u;
  MinLoc ml;
  Moments mo;
  if (!population_reduce(s.P, s.part_min, s.part_idx, s.part_mom, &ctrl->ticket, item, store, post, ml, mo)) return;
  if (threadIdx.x == 0) fin(ml, mo);
  // Islands: the last block stores the island's record (best value, global id, score moments, best row) straight into
  // every peer's exchange window over NVLink and releases a sequence flag — the per-generation "all-gather of the
  // island bests" without a collective, a launch or a wait: readers look at their own window when they need the
  // global best (nls_de_read_exchange).  Sequence = iterations completed + 1, slot parity = sequence & 1.
  if (s.xw != nullptr) {
    __syncthreads();                                     // fin()'s updates of the control block
    const XchgWindow &w = *s.xw;
    const u64 seq = *reinterpret_cast<volatile unsigned long long *>(&ctrl->iter) + 1ull;
    const u64 b = *reinterpret_cast<volatile unsigned long long *>(&ctrl->best_id);
    const u64 slot = ((seq & 1ull) * u64(w.world) + u64(w.rank)) * w.record_bytes;
    // (the location byte may have been flipped by another block of this launch: read it from L2, not from this SM's L1)
    const T *src = static_cast<const T *>(s.buf[__ldcg(s.where + b)]) + b * s.stride;
    for (int r = 0; r < w.world; r++) {
      RecordHeader *h = reinterpret_cast<RecordHeader *>(w.records[r] + slot);
      if (threadIdx.x == 0) {
        h->value = *reinterpret_cast<volatile double *>(&ctrl->best_value); h->index = s.offset + b;
        h->moments = ctrl->score_moments; h->valid = 1; h->_pad = 0;
      }
      T *row = reinterpret_cast<T *>(h + 1);
      for (u64 j = threadIdx.x; j < s.d; j += kBlock) row[j] = __ldcg(src + j);
    }
    __syncthreads();
    if (threadIdx.x < w.world) {
      __threadfence_system();
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(w.flags[threadIdx.x] + (seq & 1ull) * u64(w.world) + u64(w.rank)), "l"(seq) : "memory");
    }
  }
}
template <class T>
__global__ void __launch_bounds__(kBlock) de_commit_kernel(DEState s, int mode) {
  if (s.ctrl->stop) return;
  de_commit_pass<T>(s, mode);
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
  if (s.ctrl->stop) return;                              // an island whose stop rule has fired stays as it is
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
constexpr int kMaxDevices = 64;
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
cudaError_t de_launch_rescan(const DEState &s, const LaunchGeom &g, cudaStream_t st) {
  return de_launch_commit<T>(s, 2, g, st);
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

template <class T, int O, int W, int U, int S, bool SKIP_BASE>
void de_launch_k2_w(const DEState &s, const LaunchGeom &g, cudaStream_t st) {
  const u64 want = (s.P + kWarpsPerBlock - 1) / kWarpsPerBlock;
  auto kernel = de_generation_kernel<T, O, W, U, S, SKIP_BASE>;
  const unsigned int grid = clamp_grid(want, u64(g.sm_count) * blocks_per_sm(kernel));
  const int shift = de_tile_shift(s.P, u64(grid) * kWarpsPerBlock, 32 / W, W == 32 ? 0.5 : 2.0);
  kernel<<<grid, kBlock, 0, st>>>(s, shift);
}
// NLS_DE_BULK in the environment overrides the staging policy: 0 never, 1 best recombination only, 2 both
// recombination modes (the default), 3 both and for every row length (tests)
inline int de_bulk_mode() {
  const char *e = std::getenv("NLS_DE_BULK");
  return e ? std::atoi(e) : 2;
}
template <class T, int O, int kStages, int kSteps>
void de_launch_k2_bulk(const DEState &s, const LaunchGeom &g, cudaStream_t st) {
  auto kernel = de_generation_bulk_kernel<T, O, kStages, kSteps>;
  constexpr size_t smem = size_t(kWarpsPerBlock) * kStages * 4 * 512 * kSteps;
  // function attributes are per DEVICE: a device group launches the same kernel on several of them
  static int per_sm_of[kMaxDevices] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int &per_sm = per_sm_of[dev % kMaxDevices];
  if (per_sm == 0) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kBlock, smem);
    per_sm = n < 1 ? 1 : n;
  }
  const u64 want = (s.P + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const unsigned int grid = clamp_grid(want, u64(g.sm_count) * per_sm);
  const int shift = de_tile_shift(s.P, u64(grid) * kWarpsPerBlock, 1, 0.25);
  kernel<<<grid, kBlock, smem, st>>>(s, shift);
}
// lane-group shape by row length in 128-bit vectors (see de_generation_kernel)
template <class T, int O>
void de_launch_k2(const DEState &s, const LaunchGeom &g, cudaStream_t st) {
  const u64 vecs = (s.d + Vec<T>::V - 1) / Vec<T>::V;   // 128-bit vectors per row
  if constexpr (closed_form_dim(O) == 0) {
    const int mode = de_bulk_mode();
    const bool long_row = vecs > 64;
    if (mode >= 3 || (long_row && (mode == 2 || (mode == 1 && s.strategy == 0)))) {
      de_launch_k2_bulk<T, O, NLS_DE_BULK_STAGES, NLS_DE_BULK_STEPS>(s, g, st);
      return;
    }
  }
  if constexpr (closed_form_dim(O) > 0) {               // fixed short vectors: only the 4-lane variant exists
    de_launch_k2_w<T, O, 4, 1, 1, false>(s, g, st);
    return;
  }
  if (vecs <= 4) de_launch_k2_w<T, O, 4, 1, 1, false>(s, g, st);
  else if (vecs <= 8) de_launch_k2_w<T, O, 8, 1, 1, false>(s, g, st);
  else if (vecs <= 16) de_launch_k2_w<T, O, 8, 2, 4, false>(s, g, st);
  else if (vecs <= 32) de_launch_k2_w<T, O, 16, 2, 2, false>(s, g, st);
  else if (vecs <= 64) de_launch_k2_w<T, O, 32, 2, 1, false>(s, g, st);
  else de_launch_k2_w<T, O, 32, 1, 1, true>(s, g, st);
}

// *commit (in): the caller would like K3 fused behind the repair; (out): whether this launch does it
template <class T, int O, int W, int U, int S, bool SKIP_BASE>
cudaError_t de_launch_repair_w(const DEState &s, const LaunchGeom &g, cudaStream_t st, bool *commit) {
  DEState arg = s;
  auto kernel = de_repair_kernel<T, O, W, U, S, SKIP_BASE>;
  const unsigned int grid = clamp_grid((s.P + kBlock - 1) / kBlock, u64(g.sm_count) * blocks_per_sm(kernel));
  *commit = *commit && grid == static_cast<unsigned int>(g.reduce_blocks);
  int fuse = *commit ? 1 : 0;
  void *args[] = {&arg, &fuse};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kernel), dim3(grid), dim3(kBlock), args, 0, st);
}
// one sweep step per agent where it covers the row (W = 4 / 8 / 16 lanes), else a full warp looping over the row
template <class T, int O>
cudaError_t de_launch_repair(const DEState &s, const LaunchGeom &g, cudaStream_t st, bool *commit) {
  const u64 vecs = (s.d + Vec<T>::V - 1) / Vec<T>::V;
  if constexpr (closed_form_dim(O) > 0) return de_launch_repair_w<T, O, 4, 1, 1, false>(s, g, st, commit);
  if (vecs <= 4) return de_launch_repair_w<T, O, 4, 1, 1, false>(s, g, st, commit);
  if (vecs <= 8) return de_launch_repair_w<T, O, 8, 1, 1, false>(s, g, st, commit);
  if (vecs <= 16) return de_launch_repair_w<T, O, 16, 1, 1, false>(s, g, st, commit);
  // long rows: NLS_DE_REPAIR_U sweep steps in flight per lane (NLS_DE_REPAIR_U=2 in the environment: two)
  static const int steps = [] { const char *e = std::getenv("NLS_DE_REPAIR_U"); return e ? std::atoi(e) : NLS_DE_REPAIR_U; }();
  if (steps == 2) return de_launch_repair_w<T, O, 32, 2, 1, true>(s, g, st, commit);
  return de_launch_repair_w<T, O, 32, NLS_DE_REPAIR_U, 1, true>(s, g, st, commit);
}

// one generation: K2, K2r (cooperative), K3 — K3 inside the K2r launch for launch-bound populations (two launches
// per generation instead of three; not while the kernels are being timed one by one)
template <class T>
cudaError_t de_launch_generation(const DEState &s, const LaunchGeom &g, cudaStream_t st, cudaEvent_t *ev) {
  cudaError_t e = cudaSuccess;
  static const int fuse_mode = [] { const char *e = std::getenv("NLS_DE_FUSE_COMMIT"); return e ? std::atoi(e) : 1; }();
  bool commit = ev == nullptr && fuse_mode != 0 && (fuse_mode == 2 || s.P * s.d <= (1ull << 24));
  if (ev) cudaEventRecord(ev[0], st);
#define NLS_CALL(O)                                                                                                 \
  de_launch_k2<T, O>(s, g, st);                                                                                     \
  e = cudaGetLastError();                                                                                           \
  if (e != cudaSuccess) return e;                                                                                   \
  if (ev) cudaEventRecord(ev[1], st);                                                                               \
  e = de_launch_repair<T, O>(s, g, st, &commit);
  NLS_OBJ_SWITCH(s.objective, NLS_CALL)
#undef NLS_CALL
  if (e != cudaSuccess) return e;
  if (ev) cudaEventRecord(ev[2], st);
  if (!commit) e = de_launch_commit<T>(s, 0, g, st);
  if (ev) cudaEventRecord(ev[3], st);
  return e;
}

// the one-launch path (de_persist.cuh, compiled in its own translation units)
template <class T>
cudaError_t de_launch_persistent(const DEState &s, unsigned long long n_generations, cudaStream_t st);

template <class T>
cudaError_t de_launch_gather_rows(const DEState &s, unsigned long long first, unsigned long long count, void *out,
                                  cudaStream_t st) {
  de_gather_rows_kernel<T><<<clamp_grid(count, 1u << 16), kBlock, 0, st>>>(s, first, count, static_cast<T *>(out));
  return cudaGetLastError();
}

// (objective plugins are one translation unit compiled at run time: they leave the one-launch kernels out and small
//  populations take the graph-replay path there)
#ifdef NLS_PLUGIN_BUILD
#define NLS_PERSISTENT_OR_NULL(f) nullptr
#else
#define NLS_PERSISTENT_OR_NULL(f) f
#endif
#define NLS_DEFINE_DE_OPS(T, NAME)                                                                        \
  const DEOps *NAME() {                                                                                   \
    static const DEOps ops = {de_launch_init<T>, de_launch_generation<T>, de_launch_export_best<T>,       \
                              de_launch_migrate<T>, de_launch_gather_rows<T>, NLS_PERSISTENT_OR_NULL(de_launch_persistent<T>), \
                              de_launch_rescan<T>};                                                      \
    return &ops;                                                                                          \
  }

}  // namespace nls
