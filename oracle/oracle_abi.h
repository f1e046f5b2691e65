/*
 * oracle_abi.h — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Plain-C call surface shared by the two CPU checkers under oracle/:
 *   - popsolve_oracle.cpp : a from-scratch restatement ("port") of the reference's DE / PSO population loop
 *                           (nlsolver.h:2302-2477 DE, 2479-2742 PSO, 2037-2052 std_err, 1343-1381 xorshift);
 *   - ref_harness.cpp     : drives the UNMODIFIED reference templates (compiled from /root/reference where
 *                           they lie, output only into oracle/_ref/) through a tape RNG + objective hook.
 * Both export the same entry points with the prefix `oracle_` / `ref_` so the tests can diff them.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load these.
 */
#ifndef ORACLE_ABI_H_
#define ORACLE_ABI_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_F32 = 0, ORC_F64 = 1 };
/* objective ids: N-D forms that reduce to test_functions.h:51-92 at d = 2; id 4 is example.cpp:41-48 */
enum { ORC_SPHERE = 0, ORC_ROSENBROCK = 1, ORC_RASTRIGIN = 2, ORC_ACKLEY = 3, ORC_ROSENBROCK_EX = 4,
       /* the other problems of the reference's test driver, test_functions.h:94-318 */
       ORC_BEALE = 5, ORC_GOLDSTEIN_PRICE = 6, ORC_THREE_HUMP_CAMEL = 7, ORC_MCCORMICK = 8, ORC_SCHAFFER_N2 = 9,
       ORC_STYBLINSKI_TANG = 10, ORC_SHEKEL = 11, ORC_BOOTH = 12, ORC_BUKIN_N6 = 13, ORC_MATYAS = 14, ORC_LEVI_N13 = 15,
       ORC_CUSTOM = 100 /* callbacks installed with oracle_set_custom_objective (objective-plugin tests) */ };
/* same order as the reference enums (nlsolver.h:2377, 2496) */
enum { ORC_DE_BEST = 0, ORC_DE_RANDOM = 1 };
enum { ORC_PSO_VANILLA = 0, ORC_PSO_ACCELERATED = 1 };
/* draw source: the counter tape (DESIGN.md "RNG tape") or the reference's sequential xorshift128+ */
enum { ORC_RNG_TAPE = 0, ORC_RNG_XORSHIFT = 1 };

typedef struct {
  int32_t dtype, objective, strategy, minimize;
  uint64_t pop_size, dim;
  double crossover_prob, differential_weight, eps;
  uint64_t max_iter, best_val_no_change;
  int32_t rng_mode, _pad;
  uint64_t seed;          /* tape seed (ORC_RNG_TAPE) */
  uint64_t agent_offset;  /* global id of local agent 0 in the tape key (islands) */
  uint64_t xs_state[2];   /* xorshift128+ state (ORC_RNG_XORSHIFT); {0,0} = reference default seeding */
} orc_de_cfg;

typedef struct {
  int32_t dtype, objective, pso_type, minimize;
  uint64_t n_particles, dim;
  double inertia, cognitive_coef, social_coef, eps;
  uint64_t max_iter, best_val_no_change;
  int32_t constrained;     /* 1 = bounded overloads (threshold_positions), 0 = unbounded */
  int32_t social_index_j;  /* vanilla only: 0 = reference quirk swarm_best_position[i] (needs P <= d), 1 = [j] */
  int32_t rng_mode, _pad;
  uint64_t seed;
  uint64_t particle_offset, n_particles_global; /* sharded swarm: slice [offset, offset+n_particles) of a global swarm */
  uint64_t xs_state[2];
} orc_pso_cfg;

typedef struct {
  double f_value;
  uint64_t iterations, function_calls;
  uint64_t best_index;     /* DE: best_id ; PSO: index of the particle whose row is swarm_best_position (if valid) */
  uint64_t val_no_change;
  uint64_t draws_consumed; /* number of generator() calls made */
  int32_t best_valid;      /* PSO: 0 if swarm_best_position was never assigned */
  int32_t stop_reason;     /* 1 max_iter, 2 val_no_change, 3 std_err < eps */
  double std_err;          /* last std_err evaluated */
} orc_status;

/*
 * Buffers are caller-allocated, element type = cfg.dtype, any may be NULL.
 *  x0[d] in; x_best[d] out; rows[P*d], scores[P] = final population;
 *  decisions of the LAST executed generation: donors[P*3] (ids[1..3]), dim_idx[P], rejects[P] (rejected index
 *  proposals), accepted[P] (0/1), masks[P*d] (1 = mutated coordinate), trial_scores[P].
 */
typedef struct {
  void *x_best, *rows, *scores, *trial_scores;
  uint32_t *donors, *dim_idx, *rejects;
  uint8_t *accepted, *masks;
} orc_de_out;

typedef struct {
  void *x_best, *positions, *velocities, *pbest_values, *last_values;
} orc_pso_out;

/* ---- SANN as a batch of independent chains (nlsolver.h:2744-2815 is one chain) ----
 * Tape: chain c draws from stream (epoch e, global chain id); epoch e starts right after the chain's objective call
 * number e (0 = the initial f(x)) and its draws are numbered from 0 in the order the reference makes them: the
 * Metropolis draw of the step just evaluated (only if difference > 0), then the 2*d rnorm draws of the next step. */
typedef struct {
  int32_t dtype, objective, minimize, rng_mode;
  uint64_t n_chains, dim;
  uint64_t max_iter, temperature_iter;   /* reference defaults 5000, 10 */
  double temperature_max;                /* 10.0 */
  uint64_t seed, chain_offset;
  uint64_t xs_state[2];   /* ORC_RNG_XORSHIFT: one generator shared by the chains, run one after the other */
  uint64_t x0_count;      /* 1: every chain starts from the same x0[d]; n_chains: x0[n_chains*d] */
  uint64_t max_steps;     /* restatement only: stop after this many inner steps (0 = run to max_iter) */
} orc_sann_cfg;

typedef struct {
  void *x_best;                      /* [n_chains*d] best points (the reference's x on return) */
  void *f_best;                      /* [n_chains] best_val */
  void *p_cur;                       /* [n_chains*d] current points p (restatement only) */
  uint32_t *n_accepted, *n_improved; /* [n_chains] (restatement only) */
  uint64_t *draws;                   /* [n_chains] generator() calls made by the chain */
  uint64_t *iterations, *function_calls; /* [n_chains] as solver_status reports them */
} orc_sann_out;

/* ---- NelderMeadPSO as a batch of independent solvers (nlsolver.h:3546-3920 is one solver) ----
 * Tape: solver s draws from stream (tag, global solver id); tag 0 serves init_solver_state (nlsolver.h:3712-3721: for
 * every PSO particle p = 0 .. 2n-1 and coordinate j the position draw k = 2 (p n + j), then the velocity draw k + 1),
 * tag it + 1 serves apply_pso of loop iteration it (nlsolver.h:3843-3847: for the q-th PSO particle in sorted order
 * and coordinate j, r_p = draw 2 (q n + j), r_g = the next one).  Both consume exactly 4 n^2 draws, so the unmodified
 * reference is fed the same tape by counting.  Only the unbounded overloads exist here: the bounded ones clamp with
 * lower[i] / upper[i] indexed by the PARTICLE loop counter (nlsolver.h:3859), which is out of bounds for every particle.
 * The reference's initial simplex writes particle_positions[n][n] and reads x[n] (nlsolver.h:3697-3700, one element past
 * the end); the restatement ignores that store.  With glibc it lands in allocator slack for some n (fp64: even n; fp32:
 * 2, 4, 8, 12, ...) and corrupts the heap for the others, which the reference harness therefore refuses (returns -1). */
typedef struct {
  int32_t dtype, objective, minimize, rng_mode;
  uint64_t n_solvers, dim;
  double alpha, gamma, rho, sigma, inertia, cognitive_coef, social_coef, eps; /* defaults 1, 2, .5, .5, .8, 1.8, 1.8, 1e-6 */
  uint64_t max_iter, no_change_best_iter;                                     /* defaults 1000, 20 */
  uint64_t seed, solver_offset;
  uint64_t xs_state[2];
  uint64_t x0_count; /* 1: every solver starts from the same x0[d]; n_solvers: x0[n_solvers*d] */
} orc_nmpso_cfg;

typedef struct {
  void *x_best;  /* [n_solvers*d] */
  void *f_best;  /* [n_solvers] */
  uint64_t *iterations, *function_calls, *draws; /* [n_solvers] */
  uint8_t *ties; /* [n_solvers] 1 if a sort ever compared two equal values (std::sort is unstable: the reference's
                    order is then whatever libstdc++ produces; parity cases must be tie-free) */
} orc_nmpso_out;

#ifdef __cplusplus
}
#endif
#endif  /* ORACLE_ABI_H_ */
