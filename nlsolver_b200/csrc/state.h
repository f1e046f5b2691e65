// state.h — device-resident solver state shared between the host side of the C ABI (api.cu) and the kernels.
// Plain structs, passed to kernels by value; element buffers are void* and reinterpreted by the kernel's dtype.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nls {

// Loop counters and stop flags of DE::solve (nlsolver.h:2426-2447), kept on the device so that generations can be
// enqueued back to back: every kernel of a generation begins by reading `stop` and returns if a rule has fired.
struct Moments {  // count / mean / sum of squared deviations (Chan et al. pairwise combination)
  double n, mean, m2;
};

struct DECtrl {
  unsigned long long iter, best_id, vnc;
  unsigned long long reruns, rounds, accepted;
  double best_value, std_err;
  int stop, stop_reason, error, _pad;
  unsigned int ticket;          // last-block election of the reduction
  unsigned int acc_partial;     // accepted trials of the generation being committed
  unsigned int spec_accepted;   // trials accepted by the speculative pass K2 (0 => nothing to repair)
  unsigned int _pad2;
  unsigned int changed[3];      // repair: agents whose visible state changed in iteration k, slot k % 3
  unsigned int list_count[3];   // repair: agents to re-evaluate in iteration k, slot k % 3
  Moments score_moments;        // moments of the scores at the last scan (island exchange record)
};

struct XchgWindow;
struct DEState {
  void *buf[2];          // two agent-major row buffers [P][stride]; where[i] says which one holds agent i's row
  uint8_t *where;        // [P]
  void *score;           // [P] scores[i]
  void *tscore;          // [P] trial score of the generation in flight
  uint8_t *acc;          // [P] trial accepted (in flight) / decisions of the last generation
  uint16_t *fin;         // [P] repair: last iteration in which the agent's accept flag / accepted row changed (0xFFFF: never)
  uint4 *dec;            // [P] {ids[1], ids[2], ids[3], dim}
  uint32_t *rej;         // [P] rejected index proposals (draw offset of the crossover draws = 4 + rej)
  uint8_t *masks;        // [P*d] crossover mask of the last generation, or NULL
  uint32_t *list;        // [P] repair: agents to re-evaluate in the current iteration; migration: the top-k picks
  uint32_t *coarse;      // [3][coarse_words] repair: one bit per 32 agents, "an agent of this group was stamped in iteration k"
                         // (bitmap k % 3); NULL: no filter
  unsigned long long coarse_words;
  void *topk_scratch;    // migration top-k candidates: ceil(P / 4096) * k (key, visit) pairs
  DECtrl *ctrl;
  // reduction partials [n_partials]
  double *part_min; unsigned long long *part_idx; Moments *part_mom;
  unsigned long long P, d, stride;
  unsigned long long seed, offset;
  unsigned long long cr_le;    // crossover: mutate iff raw draw <= cr_le (integer form of `unit(raw) < CR`)
  int cr_none;                 // ... unless no raw draw satisfies it
  int strategy, objective;
  double F, fm, eps;
  unsigned long long max_iter, vnc_limit;
  // islands: exchange window (device copy of the descriptor) the commit kernel publishes the island's best record into,
  // or NULL
  const XchgWindow *xw;
};

struct PSOCtrl {
  unsigned long long iter, vnc, best_index;   // best_index: global index of the particle that last set the swarm best
  double best_value, std_err;
  int stop, stop_reason, best_valid, error;
  unsigned int ticket, _pad;
};

struct PSOState {
  void *pos, *vel;          // [P][stride]
  void *pbest, *last;       // [P] particle_best_values, values of the evaluation in flight
  void *sbest;              // [stride] swarm_best_position
  void *lower, *upper;      // [d]
  PSOCtrl *ctrl;
  double *part_min; unsigned long long *part_idx; Moments *part_mom;
  unsigned long long P, d, stride, P_global;
  unsigned long long seed, offset;
  int pso_type, objective, constrained, social_j;
  double init_inertia, cog, soc, fm, eps;
  unsigned long long max_iter, vnc_limit;
  // accelerated type: inertia of loop iteration k = pow(init_inertia, k) (nlsolver.h:2613), tabulated on the host with
  // the reference's own libm call so that the device never depends on how many generations the host has enqueued
  const double *inertia_table;
  unsigned long long inertia_n;
};

// Peer exchange window of a sharded swarm (one per rank, mapped into every peer through CUDA IPC): for each of the two
// generation parities, one record slot and one sequence flag per source rank.  A rank PUBLISHES its record by storing
// it into slot [parity][rank] of every peer's window over NVLink, then releases the flag; it GATHERS by waiting on the
// flags of its own window.  No host-side collective is involved.
constexpr int kMaxPeers = 16;
struct XchgWindow {
  char *records[kMaxPeers];                 // base of rank r's record area  [2][world][record_bytes]
  unsigned long long *flags[kMaxPeers];     // base of rank r's flag area    [2][world]
  int world, rank;
  unsigned long long record_bytes;
  // device groups (one process, every shard addressable): the shards' stop-statistic arrays in rank order, so that
  // std_err can be re-evaluated exactly as the reference does when it lands near eps; NULL otherwise
  const void *values[kMaxPeers];
  unsigned long long counts[kMaxPeers];
};

// A batch of independent simulated-annealing chains (SANN::solve, nlsolver.h:2778-2815, once per chain).  Each chain
// owns one row in each of three buffers; two of them hold its current point p and its best point x (the same buffer
// right after an improvement), the remaining one receives the next candidate — accepting a candidate is a role swap,
// never a copy.
struct SANNCtrl {
  double best_value;
  unsigned long long best_chain;   // local index of the best chain (lowest value, lowest index on ties)
  int best_valid;
  int best_buf;                    // which of the three row buffers holds that chain's x
};
struct SANNState {
  void *buf[3];            // [C][stride] each
  uint8_t *role;           // [C] bits 0-1: buffer of p, bits 2-3: buffer of x, bit 4: the last step consumed its Metropolis draw
  void *best;              // [C] best_val
  uint32_t *n_acc, *n_imp; // [C] accepted candidates / improvements of the best so far
  SANNCtrl *ctrl;
  unsigned long long C, d, stride;
  unsigned long long seed, offset;
  unsigned long long inner;       // temperature_iter - 1: candidates per temperature
  unsigned long long total_steps; // max_iter * inner
  // temperature of outer iteration k, temperature_max / log(k + 1.7182818) (nlsolver.h:2792-2793), tabulated on the host
  // with the reference's own libm call in the reference's precision
  const double *t_table;
  unsigned long long t_n;
  double tmax, scale, fm;
  int objective, _pad;
};

struct LaunchGeom {
  int sm_count;
  int reduce_blocks;   // grid of the reduction kernels (= number of partials)
};

}  // namespace nls
