/*
 * nls_b200.h — C ABI of libnls_b200.so: the B200 (sm_100a) engine for nlsolver's DE / PSO population loop.
 *
 * This is the drop-in boundary.  The reference (JSzitas/nlsolver) is a header-only C++17 template library with no
 * FFI layer of its own; the only process/device boundary on this path is the one introduced here:
 *
 *     host C++17 header (include/nlsolver_b200.hpp, same template API as the reference)
 *         -> extern "C" (this file: plain pointers, sizes and POD structs; no CUDA or torch types)
 *             -> CUDA kernels (nlsolver_b200/csrc/)
 *
 * Each entry point names the reference interface it replaces (paths are into the reference tree).  Every function
 * returns NLS_OK (0) or a negative nls_error; the message for the calling thread is nls_last_error().  Nothing
 * throws across this boundary.  There is NO CPU fallback: without a CUDA device nls_ctx_create fails.
 *
 * Threading: one solve at a time per solver handle; handles and contexts are independent of one another.
 * nls_load_objective is not synchronised against concurrent handle creation: load plugins before starting solver threads.
 * Lifetime: destroy every solver handle (nls_de / nls_pso / nls_sann / nls_xchg) before the context it was created on.
 */
#ifndef NLS_B200_H_
#define NLS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLS_B200_VERSION 101

#if defined(__GNUC__)
#define NLS_API __attribute__((visibility("default")))
#else
#define NLS_API
#endif

typedef enum {
  NLS_OK = 0,
  NLS_ERR_INVALID = -1, /* bad argument (NULL, pop_size < 4, dim < 1, unknown enum, ...) */
  NLS_ERR_CUDA = -2,    /* a CUDA runtime call failed; nls_last_error() has the CUDA message */
  NLS_ERR_NOMEM = -3,   /* device allocation failed */
  NLS_ERR_STATE = -4,   /* call not valid in the handle's current state */
  NLS_ERR_INTERNAL = -5 /* the in-place repair did not converge (should be unreachable) */
} nls_error;

enum { NLS_F32 = 0, NLS_F64 = 1 };

/* Device objective functors: N-D forms that reduce to test_functions.h:51-92 at d = 2 (summation order is the
 * canonical one described in DESIGN.md); NLS_ROSENBROCK_EX is example.cpp:41-48 / README.md:83-90. */
enum { NLS_SPHERE = 0, NLS_ROSENBROCK = 1, NLS_RASTRIGIN = 2, NLS_ACKLEY = 3, NLS_ROSENBROCK_EX = 4,
       /* the other problems of the reference's test driver (test_functions.h:94-318, 485-524).  Closed forms of fixed
        * dimension: dim must be 2 (NLS_SHEKEL: 4); NLS_STYBLINSKI_TANG is a sum of any dimension. */
       NLS_BEALE = 5, NLS_GOLDSTEIN_PRICE = 6, NLS_THREE_HUMP_CAMEL = 7, NLS_MCCORMICK = 8, NLS_SCHAFFER_N2 = 9,
       NLS_STYBLINSKI_TANG = 10, NLS_SHEKEL = 11, NLS_BOOTH = 12, NLS_BUKIN_N6 = 13, NLS_MATYAS = 14, NLS_LEVI_N13 = 15,
       NLS_OBJECTIVE_COUNT = 16 };

/* Same enumerator order as nlsolver::RecombinationStrategy {best, random} (nlsolver.h:2377) */
enum { NLS_DE_BEST = 0, NLS_DE_RANDOM = 1 };
/* Same enumerator order as nlsolver::PSOType {Vanilla, Accelerated} (nlsolver.h:2496) */
enum { NLS_PSO_VANILLA = 0, NLS_PSO_ACCELERATED = 1 };

/* nls_de_cfg.flags / nls_pso_cfg.flags */
#define NLS_FLAG_RECORD_MASKS 1u   /* DE: keep the crossover mask of the last generation (P*d bytes) for parity checks */
#define NLS_FLAG_SOCIAL_INDEX_J 2u /* vanilla PSO: social term reads swarm_best_position[j] (corrected) instead of the
                                      reference's [i] (nlsolver.h:2674), which is only defined for n_particles <= dim */

typedef struct nls_ctx nls_ctx; /* one per (process, GPU): device, stream, scratch */
typedef struct nls_de nls_de;   /* a DE population resident in HBM */
typedef struct nls_pso nls_pso; /* a PSO swarm resident in HBM */

/* Mirrors the constructor of nlsolver::DE (nlsolver.h:2390-2402) plus what the template parameters carried. */
typedef struct {
  int32_t dtype;     /* NLS_F32 | NLS_F64         <- template parameter scalar_t */
  int32_t objective; /* NLS_SPHERE ...            <- template parameter Callable (device functor tag) */
  int32_t strategy;  /* NLS_DE_BEST | NLS_DE_RANDOM  <- template parameter RecombinationType */
  int32_t minimize;  /* 1 = minimize(), 0 = maximize() (scores multiplied by -1, nlsolver.h:2418) */
  uint64_t pop_size, dim;
  double crossover_prob, differential_weight, eps; /* defaults 0.9, 0.8, 10e-4 */
  uint64_t max_iter, best_val_no_change;           /* defaults 1000, 50 */
  uint64_t seed;         /* draw-tape seed; the C++ header derives it from two draws of the user's RNG */
  uint64_t agent_offset; /* global id of local agent 0 in the tape key (island model); 0 for a single population */
  uint32_t flags, _reserved;
} nls_de_cfg;

/* Mirrors the constructor of nlsolver::PSO (nlsolver.h:2522-2551). */
typedef struct {
  int32_t dtype, objective;
  int32_t pso_type; /* NLS_PSO_VANILLA | NLS_PSO_ACCELERATED */
  int32_t minimize;
  uint64_t n_particles, dim;
  double inertia, cognitive_coef, social_coef, eps; /* defaults 0.8, 1.8, 1.8, 10e-4 */
  uint64_t max_iter, best_val_no_change;            /* defaults 5000, 50 */
  int32_t constrained; /* 1 = the (x, lower, upper) overloads: clamp positions (nlsolver.h:2577-2589, 2701-2715) */
  uint32_t flags;
  uint64_t seed;
  uint64_t particle_offset;    /* sharded swarm: this handle owns global particles [offset, offset + n_particles) */
  uint64_t n_particles_global; /* 0 = n_particles */
} nls_pso_cfg;

/* solver_status<scalar_t> (nlsolver.h:2054-2097) plus the loop counters a stepping caller needs. */
typedef struct {
  double f_value;          /* scores[best_id] (DE) / swarm_best_value (PSO), as the reference reports it */
  uint64_t iterations;     /* iter */
  uint64_t function_calls; /* function_calls_used / f_evals == P * (iter + 1) */
  uint64_t best_index;     /* DE best_id; PSO: global index of the particle that set swarm_best_position */
  uint64_t val_no_change;
  int32_t stopped;     /* a stop rule has fired; further steps are no-ops */
  int32_t stop_reason; /* 1 iter >= max_iter, 2 val_no_change >= best_val_no_change, 3 std_err < eps */
  int32_t best_valid;  /* PSO: 0 while swarm_best_position was never assigned (the reference then returns an empty x) */
  int32_t _reserved;
  double std_err;           /* last value of the std_err stop statistic (nlsolver.h:2037-2052) */
  uint64_t repair_reruns;   /* DE diagnostics: trials re-evaluated by the in-place repair passes so far */
  uint64_t repair_rounds;   /* DE diagnostics: repair rounds executed so far */
  uint64_t accepted_total;  /* DE diagnostics: trials accepted so far */
} nls_status;

NLS_API const char *nls_last_error(void);
NLS_API int nls_version(void);

/* stream: a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL for a stream owned by the context */
NLS_API int nls_ctx_create(int device, void *stream, nls_ctx **out);
NLS_API int nls_ctx_destroy(nls_ctx *ctx);
/* A context caches the device buffers of destroyed solver handles for the next solve (allocating and freeing multi-GB
 * populations costs hundreds of milliseconds).  The cache is bounded: at most `bytes` (nls_ctx_set_pool_limit; 0 = the
 * default: no more than the largest single solver handle released so far), oldest buffers evicted first; a cached
 * buffer also serves a smaller request if less than half of it would be wasted.  nls_ctx_trim returns everything to
 * the driver now — call it before another CUDA user of the process (a framework allocator, NCCL) needs the memory. */
NLS_API int nls_ctx_trim(nls_ctx *ctx);
NLS_API int nls_ctx_set_pool_limit(nls_ctx *ctx, uint64_t bytes);
NLS_API uint64_t nls_ctx_pool_bytes(const nls_ctx *ctx);
/* Debug aid: with NLS_B200_GUARD=1 in the environment every device buffer is allocated between two 256-byte guard zones
 * that are verified when its solver handle is destroyed; returns how many buffers were found overwritten so far. */
NLS_API unsigned long long nls_debug_guard_violations(void);
NLS_API int nls_ctx_device(const nls_ctx *ctx);
NLS_API int nls_ctx_sm_count(const nls_ctx *ctx);

/* ---- one-shot calls: DE::minimize / DE::maximize (nlsolver.h:2404-2410) and the four PSO::minimize / maximize
 *      overloads (nlsolver.h:2553-2589).  Host buffers; x_best receives agents[best_id] / swarm_best_position. ---- */
NLS_API int nls_de_solve(nls_ctx *ctx, const nls_de_cfg *cfg, const void *x0_host, void *x_best_host, nls_status *status);
NLS_API int nls_pso_solve(nls_ctx *ctx, const nls_pso_cfg *cfg, const void *lower_host, const void *upper_host,
                  void *x_best_host, nls_status *status);

/* ---- stepwise DE: the same loop (DE::solve, nlsolver.h:2413-2476) cut at generation boundaries ---- */
/* init_agents + initial scoring + first best scan / stop test (nlsolver.h:2415-2447); synchronous */
NLS_API int nls_de_create(nls_ctx *ctx, const nls_de_cfg *cfg, const void *x0_host, nls_de **out);
/* enqueue n generations (loop body nlsolver.h:2449-2474 + the next best scan / stop test); asynchronous */
NLS_API int nls_de_step(nls_de *de, uint64_t n_generations);
/* wait for enqueued work; status may be NULL */
NLS_API int nls_de_sync(nls_de *de, nls_status *status);
NLS_API int nls_de_read_best(nls_de *de, void *x_host);             /* agents[best_id], dim elements */
NLS_API int nls_de_read_population(nls_de *de, void *rows_host);    /* agents, pop_size * dim elements, agent-major */
NLS_API int nls_de_read_scores(nls_de *de, void *scores_host);      /* scores, pop_size elements */
/* agents[first .. first+count), count * dim elements (spot checks on populations too large to copy whole) */
NLS_API int nls_de_read_rows(nls_de *de, uint64_t first, uint64_t count, void *rows_host);
/* decisions of the last executed generation; any pointer may be NULL.
 * donors: pop_size*3 (ids[1..3] of generate_indices, nlsolver.h:2331-2355); dim_idx / rejects / accepted: pop_size;
 * trial_scores: pop_size elements of dtype; masks: pop_size*dim bytes (needs NLS_FLAG_RECORD_MASKS) */
NLS_API int nls_de_read_decisions(nls_de *de, uint32_t *donors, uint32_t *dim_idx, uint32_t *rejects, uint8_t *accepted,
                          void *trial_scores, uint8_t *masks);
NLS_API int nls_de_destroy(nls_de *de);
/* measurement hook: when enabled, CUDA events bracket the three kernels of every generation enqueued afterwards;
 * nls_de_kernel_times waits for them and returns the accumulated device milliseconds and the generation count
 * since the last call (ms[0] K2 generation pass, ms[1] K2r repair, ms[2] K3 commit + reduce). */
NLS_API int nls_de_enable_kernel_timing(nls_de *de, int enable);
NLS_API int nls_de_kernel_times(nls_de *de, double ms[3], uint64_t *generations);

/* island-model hooks (device pointers, enqueued on the context stream; no reference counterpart) */
/* record = [f_value (double), best global id (uint64), row (dim elements of dtype, padded to 8 bytes)] */
NLS_API uint64_t nls_record_bytes(int32_t dtype, uint64_t dim);
NLS_API int nls_de_export_best(nls_de *de, void *record_dev);
/* the k best agents (lowest score first, lowest index on ties): rows k*dim elements, scores k elements */
NLS_API int nls_de_export_top(nls_de *de, uint64_t k, void *rows_dev, void *scores_dev);
/* replace the k worst agents (highest score first, highest index on ties) by the given rows / scores */
NLS_API int nls_de_import_migrants(nls_de *de, uint64_t k, const void *rows_dev, const void *scores_dev);

/* ---- stepwise PSO (PSO::solve, nlsolver.h:2592-2624) ---- */
/* init_solver_state + first update_best_positions (nlsolver.h:2626-2657, 2595); synchronous.
 * For a sharded swarm the first global exchange must follow before stepping (see nls_pso_apply_candidates). */
NLS_API int nls_pso_create(nls_ctx *ctx, const nls_pso_cfg *cfg, const void *lower_host, const void *upper_host,
                   nls_pso **out);
NLS_API int nls_pso_step(nls_pso *pso, uint64_t n_generations); /* single-GPU swarm: whole generations, asynchronous */
NLS_API int nls_pso_sync(nls_pso *pso, nls_status *status);
NLS_API int nls_pso_read_best(nls_pso *pso, void *x_host);            /* swarm_best_position */
NLS_API int nls_pso_read_positions(nls_pso *pso, void *rows_host);    /* particle_positions */
NLS_API int nls_pso_read_velocities(nls_pso *pso, void *rows_host);   /* particle_velocities (vanilla) */
NLS_API int nls_pso_read_pbest_values(nls_pso *pso, void *vals_host); /* particle_best_values */
NLS_API int nls_pso_read_last_values(nls_pso *pso, void *vals_host);  /* objective values of the last evaluation */
NLS_API int nls_pso_destroy(nls_pso *pso);

/* sharded swarm (one handle per GPU): a generation is
 *   nls_pso_step_local   : move + evaluate this shard's particles, reduce to the shard's candidate record
 *   (all-gather of the records — NCCL / torch.distributed, done by the caller)
 *   nls_pso_apply_candidates : strict-< min-loc over the gathered records against the running swarm best, lowest
 *                              global index on ties; adopts the winner's row; updates the counters / stop test */
NLS_API int nls_pso_step_local(nls_pso *pso, void *record_dev);
/* copy the shard's current candidate record (e.g. the one produced by nls_pso_create) to record_dev */
NLS_API int nls_pso_export_candidate(nls_pso *pso, void *record_dev);
NLS_API int nls_pso_apply_candidates(nls_pso *pso, const void *records_dev, uint64_t n_records);

/* ---- objective plugins: the reference accepts ANY functor as Callable (README.md:127-136); on
 * the GPU a functor must be device code, so a user objective is a small .cu file built on
 * nlsolver_b200/csrc/objective_plugin.cuh and compiled with nvcc into a shared library.  nls_load_objective loads it and
 * returns an objective id (>= 100) valid wherever NLS_SPHERE ... are (nls_de_cfg.objective, nls_pso_cfg.objective). */
NLS_API int nls_load_objective(const char *plugin_path, int32_t *objective_id);

/* ---- fused exchange over peer memory (NVLink / NVSwitch), no host-side collective ----
 * Each rank owns an exchange WINDOW in its HBM that every peer maps through CUDA IPC.  nls_pso_step_fused runs, per
 * generation: move kernel -> candidate kernel whose last block stores the shard's record directly into every peer's
 * window and releases a sequence flag -> apply kernel that waits on its own window's flags and then scans the records.
 * Results are identical to nls_pso_step_local + all-gather + nls_pso_apply_candidates.
 *   1. nls_xchg_create on every rank (world <= 16, same record size = nls_record_bytes(dtype, dim))
 *   2. nls_xchg_get_handle -> exchange the NLS_XCHG_HANDLE_BYTES-byte handles between ranks (any transport)
 *   3. nls_xchg_open_peers with the rank-ordered handles
 *   4. nls_pso_attach_exchange (performs the pending initial exchange), then nls_pso_step_fused
 * A window serves exactly one swarm (its sequence flags only grow); all ranks must step the same number of times. */
#define NLS_XCHG_HANDLE_BYTES 64
typedef struct nls_xchg nls_xchg;
NLS_API int nls_xchg_create(nls_ctx *ctx, uint64_t record_bytes, int world, int rank, nls_xchg **out);
NLS_API int nls_xchg_get_handle(nls_xchg *x, void *handle_out);
NLS_API int nls_xchg_open_peers(nls_xchg *x, const void *handles);
NLS_API int nls_xchg_destroy(nls_xchg *x);
NLS_API int nls_pso_attach_exchange(nls_pso *pso, nls_xchg *x);
NLS_API int nls_pso_step_fused(nls_pso *pso, uint64_t n_generations);
/* Islands: with a window attached, the commit kernel of every generation stores the island's record (best value,
 * global agent id, score moments, best row) into every peer's window — the per-generation exchange of the island bests
 * without a collective and without a wait.  nls_de_read_exchange copies the newest record of each of the `world`
 * islands from its OWN window to the host (rank order, nls_record_bytes each; valid = 0 where nothing was published);
 * the caller makes sure every rank has synchronised its stream first (a host barrier), and that no rank steps on before
 * everybody has read.  Destroy the solver before its window. */
NLS_API int nls_de_attach_exchange(nls_de *de, nls_xchg *x);
NLS_API int nls_de_read_exchange(nls_de *de, void *records_host);

/* ---- NelderMeadPSO as a batch of independent solvers ----
 * nlsolver::NelderMeadPSO (nlsolver.h:3546-3920) is a sequential hybrid over 3 dim + 1 particles: per iteration a sort,
 * one Nelder-Mead step on the best dim + 1 and a PSO move of the other 2 dim.  Solvers never interact, so a batch — one
 * solver per start point, or n_solvers from one point on different draw streams — runs one warp per solver with the
 * reference's loop (and its accidents: DESIGN.md §10) unchanged inside each.  Only the unbounded minimize / maximize
 * exist: the reference's bounded overloads index the bounds with the particle counter (nlsolver.h:3859) and read out
 * of bounds for every particle.  2 <= dim <= 256.  With n_solvers = 1 this is NelderMeadPSO::minimize / maximize. */
typedef struct {
  int32_t dtype, objective, minimize;
  uint32_t flags;
  uint64_t n_solvers, dim;
  double alpha, gamma, rho, sigma, inertia, cognitive_coef, social_coef, eps; /* 1, 2, 0.5, 0.5, 0.8, 1.8, 1.8, 1e-6 */
  uint64_t max_iter, no_change_best_iter;                                     /* 1000, 20 */
  uint64_t seed;
  uint64_t solver_offset; /* global id of local solver 0 in the tape key */
} nls_nmpso_cfg;
/* x0_host: x0_count rows of dim elements (1: every solver starts there, or n_solvers).  Per solver: x_best_host
 * n_solvers * dim elements, f_best_host n_solvers elements, iterations / function_calls n_solvers counters (any may be
 * NULL).  status: the best solver (lowest value, lowest id on ties); function_calls summed over the batch. */
NLS_API int nls_nmpso_solve(nls_ctx *ctx, const nls_nmpso_cfg *cfg, const void *x0_host, uint64_t x0_count,
                            void *x_best_host, void *f_best_host, uint64_t *iterations_host,
                            uint64_t *function_calls_host, nls_status *status);

/* ---- several GPUs from ONE process: device groups ----
 * The reference is single-threaded C++ with no notion of devices; a C++ caller that wants the whole box gets it here
 * without a process group: a group opens one context per device and enables peer access between them.
 *   sharded swarm : contiguous slices of the GLOBAL particle ids per device, draw streams keyed by the global id, the
 *                   per-generation min-loc exchange done by the fused kernels over peer memory (no host collective):
 *                   results are identical to the same swarm on one GPU.  The stop statistic std_err is evaluated
 *                   exactly as the reference does (sequential sums over all shards) whenever it lands near eps.
 *   islands       : one reference-exact DE population per device (global agent ids rank * pop_size + i); every
 *                   migrate_every generations the `migrants` best rows move around the ring rank -> rank + 1 over
 *                   NVLink and replace the receiver's worst rows.  The status is the best island's.
 * Devices of a group must be distinct (the exchange kernels of different shards wait on one another). */
typedef struct nls_group nls_group;
typedef struct nls_pso_sharded nls_pso_sharded;
typedef struct nls_de_islands nls_de_islands;
/* devices: n_devices ordinals, or NULL for 0 .. n_devices - 1 */
NLS_API int nls_group_create(int n_devices, const int *devices, nls_group **out);
NLS_API int nls_group_destroy(nls_group *g);
NLS_API int nls_group_size(const nls_group *g);
/* cfg describes the GLOBAL swarm (n_particles = all particles; particle_offset / n_particles_global are ignored) */
NLS_API int nls_pso_sharded_create(nls_group *g, const nls_pso_cfg *cfg, const void *lower_host, const void *upper_host,
                                   nls_pso_sharded **out);
NLS_API int nls_pso_sharded_step(nls_pso_sharded *h, uint64_t n_generations);
NLS_API int nls_pso_sharded_sync(nls_pso_sharded *h, nls_status *status);
NLS_API int nls_pso_sharded_read_best(nls_pso_sharded *h, void *x_host);
/* the shard of device `rank` (read-back of positions etc. through the nls_pso_read_* calls); owned by `h` */
NLS_API int nls_pso_sharded_shard(nls_pso_sharded *h, int rank, nls_pso **shard);
NLS_API int nls_pso_sharded_destroy(nls_pso_sharded *h);
/* PSO::minimize / maximize over all devices of the group: same arguments and results as nls_pso_solve */
NLS_API int nls_pso_solve_sharded(nls_group *g, const nls_pso_cfg *cfg, const void *lower_host, const void *upper_host,
                                  void *x_best_host, nls_status *status);
/* cfg describes ONE island (pop_size agents per device; agent_offset of island r is cfg->agent_offset + r * pop_size) */
NLS_API int nls_de_islands_create(nls_group *g, const nls_de_cfg *cfg, const void *x0_host, uint64_t migrate_every,
                                  uint64_t migrants, nls_de_islands **out);
NLS_API int nls_de_islands_step(nls_de_islands *h, uint64_t n_generations);
/* status: f_value / best_index (global agent id) of the best island (lowest rank on ties), iterations of island 0,
 * function_calls summed over the islands, stopped = every island has stopped */
NLS_API int nls_de_islands_sync(nls_de_islands *h, nls_status *status);
NLS_API int nls_de_islands_read_best(nls_de_islands *h, void *x_host);
NLS_API int nls_de_islands_island(nls_de_islands *h, int rank, nls_de **island);
NLS_API int nls_de_islands_destroy(nls_de_islands *h);
/* DE::minimize / maximize as islands over all devices of the group; runs until every island's stop rule has fired */
NLS_API int nls_de_solve_islands(nls_group *g, const nls_de_cfg *cfg, const void *x0_host, uint64_t migrate_every,
                                 uint64_t migrants, void *x_best_host, nls_status *status);

/* ---- simulated annealing as a batch of independent chains ----
 * nlsolver::SANN (nlsolver.h:2744-2815) is one sequential chain; chains never interact, so many of them — multi-start
 * from one point, or one start point per chain — run side by side with the reference's loop unchanged inside each.
 * Chain c draws from tape stream (epoch, chain_offset + c); with n_chains = 1 this is SANN::minimize / maximize. */
typedef struct nls_sann nls_sann;
/* Mirrors the constructor of nlsolver::SANN (nlsolver.h:2757-2766). */
typedef struct {
  int32_t dtype, objective, minimize;
  uint32_t flags;
  uint64_t n_chains, dim;
  uint64_t max_iter, temperature_iter; /* defaults 5000, 10: temperature_iter - 1 candidates per temperature */
  double temperature_max;              /* default 10.0 */
  uint64_t seed;
  uint64_t chain_offset; /* global id of local chain 0 in the tape key (chains sharded over GPUs); 0 otherwise */
} nls_sann_cfg;
/* x0_host: x0_count rows of dim elements; x0_count = 1 (every chain starts there) or n_chains.  Evaluates f(x0). */
NLS_API int nls_sann_create(nls_ctx *ctx, const nls_sann_cfg *cfg, const void *x0_host, uint64_t x0_count, nls_sann **out);
/* enqueue up to n_candidates candidates per chain (clamped to what is left of max_iter * (temperature_iter - 1)) */
NLS_API int nls_sann_step(nls_sann *sa, uint64_t n_candidates);
/* status of the batch: f_value / best_index = the best chain (lowest best_val, lowest global chain id on ties),
 * iterations = outer iterations completed, function_calls = n_chains * (1 + candidates evaluated per chain) */
NLS_API int nls_sann_sync(nls_sann *sa, nls_status *status);
NLS_API int nls_sann_read_best(nls_sann *sa, void *x_host); /* the best chain's x, dim elements */
/* per-chain results, any pointer may be NULL: x_best / p_cur n_chains * dim elements (the reference's x and p),
 * f_best n_chains elements, n_accepted / n_improved n_chains counters */
NLS_API int nls_sann_read_chains(nls_sann *sa, void *x_best_host, void *f_best_host, void *p_cur_host,
                                 uint32_t *n_accepted, uint32_t *n_improved);
NLS_API int nls_sann_destroy(nls_sann *sa);
/* one-shot: all chains to max_iter; x_best_host receives the best chain's x (dim elements) */
NLS_API int nls_sann_solve(nls_ctx *ctx, const nls_sann_cfg *cfg, const void *x0_host, uint64_t x0_count,
                           void *x_best_host, nls_status *status);

#ifdef __cplusplus
}
#endif
#endif /* NLS_B200_H_ */
