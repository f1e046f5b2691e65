/*
 * popsolve_oracle.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch CPU restatement ("port") of the one hot path this repo accelerates: the population
 * generation loop of nlsolver::DE and nlsolver::PSO.  It is the checker the CUDA path is compared with; it is
 * never linked into, imported by, or called from the product library (nlsolver_b200/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pinning: this file is itself checked (tests/test_oracle_vs_reference.py) against
 *   (1) the UNMODIFIED reference templates driven by oracle/ref_harness.cpp (built into oracle/_ref/) on the same
 *       draw tape — bit-for-bit in x, f_value, iterations, function_calls, populations and decisions;
 *   (2) the known-answer vectors the survey derived from the reference: xorshift default state
 *       and first draws, the README DE snippet result, the example.cpp DE stdout.
 *
 * Every function cites the reference lines it follows (paths are into /root/reference).
 * Build: g++ -std=c++17 -O2 -ffp-contract=off (FMA contraction changes result bits).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "oracle_abi.h"

namespace {

typedef uint64_t u64;

/* ---------------------------------------------------------------- RNG ------------------------------------ */

const u64 kGolden = 0x9E3779B97F4A7C15ull;

/* splitmix64 output function, nlsolver.h:1267-1270 / 1274-1277 (the part after the state increment) */
inline u64 mix64(u64 z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* Counter tape (DESIGN.md "RNG tape"): one splitmix64 stream per (generation tag, global agent id). */
inline u64 tape_key(u64 seed, u64 gen, u64 agent) { return mix64(mix64(seed + kGolden * (gen + 1)) ^ agent); }
inline u64 tape_draw(u64 key, u64 k) { return mix64(key + kGolden * (k + 1)); }

/* u64 -> [0,1] exactly as the reference generators do: T(u) / T(2^64-1)  (nlsolver.h:1358-1359, 1270, 1177) */
template <class T>
inline T unit(u64 u) {
  return static_cast<T>(u / static_cast<T>(18446744073709551615U));
}

/* rng::splitmix::yield_init with the fixed seed 12374563468, nlsolver.h:1265, 1273-1278 */
inline u64 splitmix_first() { return mix64(12374563468ull + kGolden); }

/* rng::xorshift (xorshift128+, shifts 23/18/5), nlsolver.h:1343-1361 */
struct XorShift {
  u64 a, b;
  XorShift() { a = splitmix_first(); b = a >> 32; }   /* nlsolver.h:1345-1349 */
  u64 next() {
    u64 t = a;
    const u64 s = b;
    a = s;
    t ^= t << 23;
    t ^= t >> 18;
    t ^= s ^ (s >> 5);
    b = t;
    return t + s;
  }
};

/* A draw source hands out raw u64s; `begin(gen, agent)` marks the start of an agent's draws in a generation. */
struct TapeSource {
  u64 seed, offset, key = 0, k = 0, total = 0;
  TapeSource(u64 s, u64 off) : seed(s), offset(off) {}
  void begin(u64 gen, u64 agent) { key = tape_key(seed, gen, offset + agent); k = 0; }
  u64 next() { total++; return tape_draw(key, k++); }
};
struct SeqSource {
  XorShift g; u64 total = 0;
  explicit SeqSource(const u64 st[2]) { if (st[0] | st[1]) { g.a = st[0]; g.b = st[1]; } }
  void begin(u64, u64) {}
  u64 next() { total++; return g.next(); }
};

/* ---------------------------------------------------------------- objectives ----------------------------- */
/*
 * N-D objectives (ours; the reference has only 2-D closed forms, test_functions.h:51-92).  Canonical summation
 * order, shared with the device functors: term j is added, in increasing j, to accumulator ((j / V) % 32) with
 * V = 16 / sizeof(T); the 32 accumulators are then combined by an xor-butterfly (offsets 16,8,4,2,1).
 * Each form is arranged so that at d = 2 it performs exactly the reference's operations in the reference's order.
 */
template <class T>
struct Lanes {
  T acc[32];
  explicit Lanes(T first = 0) { for (auto &a : acc) a = 0; acc[0] = first; }
  static size_t lane(size_t j) { return (j / (16 / sizeof(T))) % 32; }
  void add(size_t j, T term) { T &a = acc[lane(j)]; a = a + term; }
  T total() {
    for (int off = 16; off >= 1; off >>= 1) {
      T nxt[32];
      for (int l = 0; l < 32; l++) nxt[l] = acc[l] + acc[l ^ off];
      std::memcpy(acc, nxt, sizeof(acc));
    }
    return acc[0];
  }
};

/* custom objective of the objective-plugin tests: f = finish(seed + sum_j term(x[j], x[j-1], j, d), d) in the canonical
 * lane order, the contract of nlsolver_b200/csrc/objective_plugin.cuh */
typedef double (*orc_term_fn)(double x, double x_prev, uint64_t j, uint64_t d);
typedef double (*orc_finish_fn)(double sum, uint64_t d);
typedef double (*orc_full_fn)(const double *x, uint64_t d);
struct CustomObjective {
  int pairwise = 0; double seed = 0; orc_term_fn term = nullptr; orc_finish_fn finish = nullptr;
  orc_full_fn full = nullptr;   /* closed form over the whole vector (plugin full_dim mode); takes precedence */
} g_custom;

template <class T>
T objective(int id, const T *x, size_t d) {
  const T two_pi = static_cast<T>(2 * M_PI);
  switch (id) {
    case ORC_CUSTOM: {
      if (g_custom.full) {
        std::vector<double> xd(x, x + d);
        return static_cast<T>(g_custom.full(xd.data(), d));
      }
      if (!g_custom.term || !g_custom.finish) return std::nan("");
      Lanes<T> s(static_cast<T>(g_custom.seed));
      for (size_t j = g_custom.pairwise ? 1 : 0; j < d; j++)
        s.add(j, static_cast<T>(g_custom.term(x[j], j ? x[j - 1] : 0, j, d)));
      return static_cast<T>(g_custom.finish(s.total(), d));
    }
    case ORC_SPHERE: {  /* test_functions.h:55  x0*x0 + x1*x1 */
      Lanes<T> s;
      for (size_t j = 0; j < d; j++) s.add(j, x[j] * x[j]);
      return s.total();
    }
    case ORC_ROSENBROCK: {  /* test_functions.h:63-66  100*pow(x0*x0 - x1, 2) + pow(x0 - 1, 2) */
      Lanes<T> s;
      for (size_t j = 1; j < d; j++) {
        const T p = x[j - 1] * x[j - 1] - x[j], q = x[j - 1] - 1;
        s.add(j, static_cast<T>(100.0) * (p * p) + q * q);
      }
      return s.total();
    }
    case ORC_ROSENBROCK_EX: {  /* example.cpp:41-48  t1*t1 + 100*t2*t2, t1 = 1-x0, t2 = x1 - x0*x0 */
      Lanes<T> s;
      for (size_t j = 1; j < d; j++) {
        const T t1 = 1 - x[j - 1], t2 = x[j] - x[j - 1] * x[j - 1];
        s.add(j, t1 * t1 + static_cast<T>(100) * t2 * t2);
      }
      return s.total();
    }
    case ORC_RASTRIGIN: {  /* test_functions.h:74-77  2*10 + (x0*x0 - 10*cos(2*pi*x0)) + (...) */
      Lanes<T> s(static_cast<T>(10) * static_cast<T>(d));
      for (size_t j = 0; j < d; j++) s.add(j, x[j] * x[j] - static_cast<T>(10) * std::cos(two_pi * x[j]));
      return s.total();
    }
    case ORC_ACKLEY: {  /* test_functions.h:85-90 */
      Lanes<T> sq, cs;
      for (size_t j = 0; j < d; j++) { sq.add(j, x[j] * x[j]); cs.add(j, std::cos(two_pi * x[j])); }
      const T inv_d = static_cast<T>(1.0) / static_cast<T>(d);   /* 0.5 at d = 2 */
      const T a = static_cast<T>(-20) * std::exp(static_cast<T>(-0.2) * std::sqrt(inv_d * sq.total()));
      const T b = -std::exp(inv_d * cs.total());
      return a + b + static_cast<T>(std::exp(1.0)) + static_cast<T>(20);
    }
    /* ---- the other problems of the reference's test driver (test_functions.h:94-318).  pow(v, 2) is restated as v*v,
     *      pow(v, 4) as (v*v)*(v*v), pow(v, 6) as ((v*v)*(v*v))*(v*v); everything in T. ---- */
    case ORC_BEALE: {  /* :98-102 */
      const T xy = x[0] * x[1], xyy = xy * x[1], xyyy = xyy * x[1];
      const T a = static_cast<T>(1.5) - x[0] + xy, b = static_cast<T>(2.25) - x[0] + xyy, c = static_cast<T>(2.625) - x[0] + xyyy;
      return a * a + b * b + c * c;
    }
    case ORC_GOLDSTEIN_PRICE: {  /* :109-117 */
      const T x0 = x[0], x1 = x[1];
      const T s1 = x0 + x1 + 1, s2 = 2 * x0 - 3 * x1;
      const T a = 1 + (s1 * s1) * (19 - 14 * x0 + 3 * x0 * x0 - 14 * x1 + 6 * x0 * x1 + 3 * x1 * x1);
      const T b = 30 + (s2 * s2) * (18 - 32 * x0 + 12 * x0 * x0 + 48 * x1 - 36 * x0 * x1 + 27 * x1 * x1);
      return a * b;
    }
    case ORC_THREE_HUMP_CAMEL: {  /* :146-148 */
      const T x2 = x[0] * x[0], x4 = x2 * x2, x6 = x4 * x2;
      return 2 * x[0] * x[0] - static_cast<T>(1.05) * x4 + x6 / 6 + x[0] * x[1] + x[1] * x[1];
    }
    case ORC_MCCORMICK: {  /* :209-211 */
      const T dlt = x[0] - x[1];
      return std::sin(x[0] + x[1]) + dlt * dlt - static_cast<T>(1.5) * x[0] + static_cast<T>(2.5) * x[1] + 1;
    }
    case ORC_SCHAFFER_N2: {  /* :219-221 */
      const T sn = std::sin(x[0] * x[0] - x[1] * x[1]);
      const T dn = 1 + static_cast<T>(0.001) * (x[0] * x[0] + x[1] * x[1]);
      return static_cast<T>(0.5) + (sn * sn - static_cast<T>(0.5)) / (dn * dn);
    }
    case ORC_STYBLINSKI_TANG: {  /* :246-252, any dimension */
      Lanes<T> s;
      for (size_t j = 0; j < d; j++) { const T x2 = x[j] * x[j]; s.add(j, x2 * x2 - 16 * x2 + 5 * x[j]); }
      return s.total() / static_cast<T>(2.0);
    }
    case ORC_SHEKEL: {  /* :258-276 */
      const T a[40] = {4, 4, 4, 4, 1, 1, 1, 1, 8, 8, 8, 8, 6, 6, 6, 6, 3, 7, 3, 7,
                       2, 9, 2, 9, 5, 5, 3, 3, 8, 1, 8, 1, 6, 2, 6, 2, 7, static_cast<T>(3.6), 7, static_cast<T>(3.2)};
      const T c[10] = {static_cast<T>(0.1), static_cast<T>(0.2), static_cast<T>(0.2), static_cast<T>(0.4), static_cast<T>(0.4),
                       static_cast<T>(0.6), static_cast<T>(0.3), static_cast<T>(0.7), static_cast<T>(0.5), static_cast<T>(0.5)};
      T sum = 0;
      for (int i = 0; i < 10; i++) {
        T inner = 0;
        for (int j = 0; j < 4; j++) { const T dlt = x[j] - a[i * 4 + j]; inner += dlt * dlt; }
        sum += static_cast<T>(1.0) / (inner + c[i]);
      }
      return -sum;
    }
    case ORC_BOOTH: {  /* :283-285 */
      const T a = x[0] + 2 * x[1] - 7, b = 2 * x[0] + x[1] - 5;
      return a * a + b * b;
    }
    case ORC_BUKIN_N6:  /* :292-295 */
      return 100 * std::sqrt(std::abs(x[1] - static_cast<T>(0.01) * x[0] * x[0])) + static_cast<T>(0.01) * std::abs(x[0] + 10);
    case ORC_MATYAS:  /* :302-304 */
      return static_cast<T>(0.26) * (x[0] * x[0] + x[1] * x[1]) - static_cast<T>(0.48) * x[0] * x[1];
    case ORC_LEVI_N13: {  /* :311-317 */
      const T pi3 = static_cast<T>(3 * M_PI), pi2 = static_cast<T>(2 * M_PI);
      const T s0 = std::sin(pi3 * x[0]), s1 = std::sin(pi3 * x[1]), s2 = std::sin(pi2 * x[1]);
      const T a = x[0] - 1, b = x[1] - 1;
      return s0 * s0 + (a * a) * (1 + s1 * s1) + (b * b) * (1 + s2 * s2);
    }
  }
  return std::nan("");
}

/* std_err, nlsolver.h:2037-2052 (sequential sums; pow(.,2) is evaluated in double for float input) */
template <class T>
T std_err(const std::vector<T> &x) {
  T mean_val = 0, result = 0;
  for (size_t i = 0; i < x.size(); i++) mean_val += x[i];
  mean_val /= static_cast<T>(x.size());
  for (size_t i = 0; i < x.size(); i++) result += std::pow(x[i] - mean_val, 2);
  result /= static_cast<T>(x.size() - 1);
  return std::sqrt(result);
}

/* generate_index, nlsolver.h:2325-2329.  The reference indexes out of bounds when the draw is exactly 1.0;
 * that case is defined here (and on the device) as max-1 and is excluded from parity tapes. */
template <class T>
size_t gen_index(T u, size_t max) {
  size_t v = static_cast<size_t>(u * max);
  return v >= max ? max - 1 : v;
}

/* ---------------------------------------------------------------- DE ------------------------------------- */
/* DE::solve (nlsolver.h:2413-2476) cut at generation boundaries: `scan()` is the top of the while loop (best scan,
 * val_no_change, stop test), `generation()` the body.  de_run() below strings them together exactly like the
 * reference; the island tests call them one at a time and inject migrants in between. */
template <class T, class Src>
struct DELoop {
  const orc_de_cfg c;
  Src src;
  const size_t P, d;
  const T CR, F, eps, fm;
  std::vector<T> A, scores, trial, tscores;
  std::vector<uint32_t> donors, dimv, rej;
  std::vector<uint8_t> acc, masks;
  u64 fcalls = 0, iter = 0, best_id = 0, vnc = 0;
  int stop_reason = 0;
  T se = 0;

  DELoop(const orc_de_cfg &cfg, const T *x0, Src source, bool want_masks)
      : c(cfg), src(source), P(cfg.pop_size), d(cfg.dim), CR(static_cast<T>(cfg.crossover_prob)),
        F(static_cast<T>(cfg.differential_weight)), eps(static_cast<T>(cfg.eps)),
        fm(cfg.minimize ? static_cast<T>(1.0) : static_cast<T>(-1.0)),   /* nlsolver.h:2418 */
        A(P * d), scores(P), trial(d), tscores(P), donors(P * 3), dimv(P), rej(P), acc(P),
        masks(want_masks ? P * d : 0) {
    /* init_agents / generate_sequence, nlsolver.h:2302-2323: agent[j] = (g() - 0.5) * x0[j], agent-major */
    for (size_t i = 0; i < P; i++) {
      src.begin(0, i);
      for (size_t j = 0; j < d; j++) A[i * d + j] = static_cast<T>((unit<T>(src.next()) - 0.5) * x0[j]);
    }
    for (size_t i = 0; i < P; i++) scores[i] = fm * objective<T>(c.objective, &A[i * d], d);   /* :2423-2425 */
    fcalls = P;
  }

  /* top of the loop, nlsolver.h:2430-2447; returns true when a stop rule fires */
  bool scan() {
    bool not_updated = true;
    for (size_t i = 0; i < P; i++)
      if (scores[i] < scores[best_id]) { best_id = i; not_updated = false; }
    vnc = not_updated * (vnc + 1);
    stop_reason = 0;
    if (iter >= c.max_iter) stop_reason = 1;      /* short-circuit order preserved */
    else if (vnc >= c.best_val_no_change) stop_reason = 2;
    else { se = std_err(scores); if (se < eps) stop_reason = 3; }
    return stop_reason != 0;
  }

  /* loop body, nlsolver.h:2449-2474: sequential and IN PLACE */
  void generation() {
    for (size_t i = 0; i < P; i++) {
      src.begin(iter + 1, i);
      const size_t fixed = (c.strategy == ORC_DE_RANDOM) ? i : best_id;   /* :2451-2457 */
      size_t ids[4] = {fixed, 0, 0, 0};
      uint32_t n = 1, rejected = 0;
      while (n < 4) {                             /* generate_indices, :2331-2355 */
        const size_t prop = gen_index<T>(unit<T>(src.next()), P);
        bool used = false;
        for (uint32_t q = 0; q < n; q++) used |= (ids[q] == prop);
        if (used) rejected++; else ids[n++] = prop;
      }
      const size_t dim = gen_index<T>(unit<T>(src.next()), d);   /* propose_new_agent, :2357-2375 */
      for (size_t j = 0; j < d; j++) {
        const bool mut = (unit<T>(src.next()) < CR) || (j == dim);
        trial[j] = mut ? A[ids[1] * d + j] + F * (A[ids[2] * d + j] - A[ids[3] * d + j]) : A[ids[0] * d + j];
        if (!masks.empty()) masks[i * d + j] = mut;
      }
      const T score = fm * objective<T>(c.objective, trial.data(), d);   /* :2463-2464 */
      fcalls++;
      const bool ok = score < scores[i];          /* :2466-2471 (NaN never accepted) */
      if (ok) { std::memcpy(&A[i * d], trial.data(), d * sizeof(T)); scores[i] = score; }
      donors[i * 3] = ids[1]; donors[i * 3 + 1] = ids[2]; donors[i * 3 + 2] = ids[3];
      dimv[i] = dim; rej[i] = rejected; acc[i] = ok; tscores[i] = score;
    }
    iter++;
  }

  /* island hooks (no reference counterpart; mirror of the semantics in include/nls_b200.h):
   * the k best agents, ascending score then ascending index ... */
  void export_top(size_t k, T *rows, T *sc) const {
    std::vector<size_t> order(P);
    for (size_t i = 0; i < P; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return scores[a] < scores[b]; });
    for (size_t e = 0; e < k; e++) {
      std::memcpy(rows + e * d, &A[order[e] * d], d * sizeof(T));
      sc[e] = scores[order[e]];
    }
  }
  /* ... replace the k worst (descending score, then descending index) and re-scan the best without touching the
   * stop counters, except that an improved best resets val_no_change */
  void import_migrants(size_t k, const T *rows, const T *sc) {
    std::vector<size_t> order(P);
    for (size_t i = 0; i < P; i++) order[i] = P - 1 - i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return scores[a] > scores[b]; });
    for (size_t e = 0; e < k; e++) {
      std::memcpy(&A[order[e] * d], rows + e * d, d * sizeof(T));
      scores[order[e]] = sc[e];
    }
    bool updated = false;
    for (size_t i = 0; i < P; i++)
      if (scores[i] < scores[best_id]) { best_id = i; updated = true; }
    if (updated) vnc = 0;
  }

  void report(const orc_de_out *out, orc_status *st) const {
    if (st) {
      st->f_value = scores[best_id]; st->iterations = iter; st->function_calls = fcalls;
      st->best_index = best_id; st->val_no_change = vnc; st->draws_consumed = src.total;
      st->best_valid = 1; st->stop_reason = stop_reason; st->std_err = se;
    }
    if (!out) return;
    if (out->x_best) std::memcpy(out->x_best, &A[best_id * d], d * sizeof(T));
    if (out->rows) std::memcpy(out->rows, A.data(), P * d * sizeof(T));
    if (out->scores) std::memcpy(out->scores, scores.data(), P * sizeof(T));
    if (out->trial_scores) std::memcpy(out->trial_scores, tscores.data(), P * sizeof(T));
    if (out->donors) std::memcpy(out->donors, donors.data(), P * 3 * sizeof(uint32_t));
    if (out->dim_idx) std::memcpy(out->dim_idx, dimv.data(), P * sizeof(uint32_t));
    if (out->rejects) std::memcpy(out->rejects, rej.data(), P * sizeof(uint32_t));
    if (out->accepted) std::memcpy(out->accepted, acc.data(), P);
    if (out->masks && !masks.empty()) std::memcpy(out->masks, masks.data(), P * d);
  }
};

template <class T, class Src>
void de_run(const orc_de_cfg &c, const T *x0, Src &src, const orc_de_out *out, orc_status *st) {
  DELoop<T, Src> loop(c, x0, src, out && out->masks);
  while (!loop.scan()) loop.generation();
  loop.report(out, st);
  src = loop.src;
}

/* ---------------------------------------------------------------- PSO ------------------------------------ */

/* rnorm, nlsolver.h:2479-2485: sqrt(-2*log(g())) * cos(2*pi_*g()), pi_ = 3.141593; g++ draws the log operand
 * first.  For T = float the unqualified log/cos/sqrt calls resolve to the double
 * versions, which is what the promotions below reproduce. */
template <class T>
T rnorm(T u_log, T u_cos) {
  constexpr T pi_ = 3.141593;
  return static_cast<T>(std::sqrt(-2 * std::log(static_cast<double>(u_log))) *
                        std::cos(static_cast<double>(2 * pi_ * u_cos)));
}

/* PSO::solve (nlsolver.h:2592-2624) cut into its phases.  A shard owns particles [offset, offset + P) of a global
 * swarm: `evaluate()` is the per-particle part of update_best_positions and returns the shard's strict-< candidate,
 * `adopt()` the swarm-level part (called with the winner over all shards), `move()` the position update. */
template <class T, class Src>
struct PSOLoop {
  const orc_pso_cfg c;
  Src src;
  const size_t P, d;
  const bool vanilla;
  const T init_inertia, cog, soc, eps, fm;
  T inertia;
  std::vector<T> lower, upper, X, V, pbest, last, sbest;
  T sbest_val = static_cast<T>(100000.0);          /* init_solver_state, nlsolver.h:2631 */
  u64 f_evals = 0, vnc = 0, iter = 0, sbest_idx = 0;
  int stop_reason = 0;
  T se = 0;

  PSOLoop(const orc_pso_cfg &cfg, const T *lo, const T *up, Src source)
      : c(cfg), src(source), P(cfg.n_particles), d(cfg.dim), vanilla(cfg.pso_type == ORC_PSO_VANILLA),
        init_inertia(static_cast<T>(cfg.inertia)), cog(static_cast<T>(cfg.cognitive_coef)),
        soc(static_cast<T>(cfg.social_coef)), eps(static_cast<T>(cfg.eps)),
        fm(cfg.minimize ? static_cast<T>(1.0) : static_cast<T>(-1.0)), inertia(init_inertia),
        lower(lo, lo + d), upper(up, up + d), X(P * d), V(vanilla ? P * d : 0), pbest(P, static_cast<T>(10000)),
        last(P) {
    for (size_t i = 0; i < P; i++) {               /* init_solver_state, nlsolver.h:2626-2657 */
      src.begin(0, i);
      for (size_t j = 0; j < d; j++) {
        const T temp = std::abs(upper[j] - lower[j]);
        X[i * d + j] = lower[j] + ((upper[j] - lower[j]) * unit<T>(src.next()));
        if (vanilla) V[i * d + j] = -temp + (unit<T>(src.next()) * temp);
      }
    }
  }
  /* per-particle part of update_best_positions, :2721-2733; candidate = lowest value, lowest index on ties */
  bool evaluate(T *cand_value, size_t *cand_local) {
    bool any = false;
    for (size_t i = 0; i < P; i++) {
      const T temp = fm * objective<T>(c.objective, &X[i * d], d);
      last[i] = temp;
      if (temp < pbest[i]) pbest[i] = temp;
      if (!any ? temp < std::numeric_limits<T>::infinity() : temp < *cand_value) { *cand_value = temp; *cand_local = i; any = true; }
    }
    return any;
  }
  /* swarm-level part, :2723-2740: `value` / `global_index` / `row` describe the best candidate over all shards
   * (have = false if no shard had one) */
  void adopt(bool have, T value, u64 global_index, const T *row, size_t n_global) {
    u64 best_index = 0;
    if (have && value < sbest_val) { sbest_val = value; best_index = global_index; sbest.assign(row, row + d); sbest_idx = global_index; }
    f_evals += n_global;
    vnc = (best_index == 0) * (vnc + 1);
  }
  /* top of the loop, :2599-2600, with the std_err of ALL particle_best_values (moments supplied by the caller for a
   * sharded swarm, computed here for a whole one) */
  bool stop_test(T std_err_all) {
    stop_reason = 0;
    if (iter >= c.max_iter) stop_reason = 1;
    else if (vnc >= c.best_val_no_change) stop_reason = 2;
    else { se = std_err_all; if (se < eps) stop_reason = 3; }
    return stop_reason != 0;
  }
  void move() {
    if (vanilla) {                                  /* update_velocities, :2658-2677 (both quirks kept) */
      for (size_t i = 0; i < P; i++) {
        src.begin(iter + 1, i);
        const size_t gi = c.particle_offset + i;
        for (size_t j = 0; j < d; j++) {
          const T r_p = unit<T>(src.next()), r_g = unit<T>(src.next());
          const T x = X[i * d + j];
          const T sb = sbest.empty() ? static_cast<T>(0) : sbest[c.social_index_j ? j : gi];
          V[i * d + j] = (inertia * V[i * d + j]) + cog * r_p * (x - x) + soc * r_g * (sb - x);
        }
      }
      for (size_t q = 0; q < P * d; q++) X[q] += V[q];   /* update_positions, :2679-2686 */
    } else {
      inertia = std::pow(init_inertia, iter);       /* :2613 */
      for (size_t i = 0; i < P; i++) {              /* update_positions, :2687-2699 */
        src.begin(iter + 1, i);
        for (size_t j = 0; j < d; j++) {
          const T u_log = unit<T>(src.next()), u_cos = unit<T>(src.next());
          const T sb = sbest.empty() ? static_cast<T>(0) : sbest[j];
          X[i * d + j] = inertia * rnorm<T>(u_log, u_cos) + (1 - cog) * X[i * d + j] + soc * sb;
        }
      }
    }
    if (c.constrained)                              /* threshold_positions, :2701-2715 */
      for (size_t i = 0; i < P; i++)
        for (size_t j = 0; j < d; j++) {
          T &p = X[i * d + j];
          p = p < lower[j] ? lower[j] : p;
          p = p > upper[j] ? upper[j] : p;
        }
  }
  void report(const orc_pso_out *out, orc_status *st) const {
    if (st) {
      st->f_value = sbest_val; st->iterations = iter; st->function_calls = f_evals;
      st->best_index = sbest_idx; st->val_no_change = vnc; st->draws_consumed = src.total;
      st->best_valid = !sbest.empty(); st->stop_reason = stop_reason; st->std_err = se;
    }
    if (!out) return;
    if (out->x_best && !sbest.empty()) std::memcpy(out->x_best, sbest.data(), d * sizeof(T));
    if (out->positions) std::memcpy(out->positions, X.data(), P * d * sizeof(T));
    if (out->velocities && vanilla) std::memcpy(out->velocities, V.data(), P * d * sizeof(T));
    if (out->pbest_values) std::memcpy(out->pbest_values, pbest.data(), P * sizeof(T));
    if (out->last_values) std::memcpy(out->last_values, last.data(), P * sizeof(T));
  }
};

template <class T, class Src>
void pso_run(const orc_pso_cfg &c, const T *lower, const T *upper, Src &src, const orc_pso_out *out,
             orc_status *st) {
  PSOLoop<T, Src> loop(c, lower, upper, src);
  auto update_best = [&]() {                        /* update_best_positions, :2716-2741 */
    T v = 0; size_t i = 0;
    const bool any = loop.evaluate(&v, &i);
    loop.adopt(any, v, c.particle_offset + i, any ? &loop.X[i * loop.d] : nullptr, loop.P);
  };
  update_best();                                    /* solve, :2592-2624 */
  while (!loop.stop_test(std_err(loop.pbest))) {
    loop.move();
    update_best();
    loop.iter++;
  }
  loop.report(out, st);
  src = loop.src;
}

/* ---------------------------------------------------------------- SANN ----------------------------------- */
/* One chain of SANN::solve (nlsolver.h:2778-2815).  `src.begin(e, chain)` opens epoch e (see oracle_abi.h); with the
 * sequential xorshift source it is a no-op.  For T = float the reference's unqualified exp() is the double overload
 * and the comparison with the float draw runs in double, as do `difference <= 0.0` and `1.0 / temperature_max`. */
template <class T, class Src>
void sann_chain(const orc_sann_cfg &c, u64 chain, const T *x0, Src &src, T *x, T *p_out, T *f_best,
                uint32_t *n_acc_out, uint32_t *n_imp_out, u64 *iters_out, u64 *evals_out) {
  const size_t d = c.dim;
  const T fm = c.minimize ? static_cast<T>(1.0) : static_cast<T>(-1.0);
  const T e_minus_1 = static_cast<T>(1.7182818), tmax = static_cast<T>(c.temperature_max);
  std::vector<T> best(x0, x0 + d), p = best, ptry = best;
  T best_val = fm * objective<T>(c.objective, best.data(), d);          /* :2781 */
  u64 evals = 1, steps = 0, iter = 0;
  uint32_t n_acc = 0, n_imp = 0;
  src.begin(0, chain);
  const T scale = static_cast<T>(1.0 / tmax);                           /* :2782 */
  bool cut = false;
  while (iter < c.max_iter && !cut) {                                   /* :2786-2790 */
    const T t = tmax / std::log(static_cast<T>(iter) + e_minus_1);      /* :2792-2793 */
    for (size_t j = 1; j < c.temperature_iter; j++) {                   /* :2794 */
      if (c.max_steps && steps >= c.max_steps) { cut = true; break; }
      const T current_scale = t * scale;
      for (size_t i = 0; i < d; i++) {                                  /* :2797-2800 */
        const T u_log = unit<T>(src.next()), u_cos = unit<T>(src.next());
        ptry[i] = p[i] + current_scale * rnorm<T>(u_log, u_cos);
      }
      const T current_val = fm * objective<T>(c.objective, ptry.data(), d);
      evals++; steps++;
      src.begin(evals - 1, chain);
      const T difference = current_val - best_val;                      /* NB: against best_val, not f(p) */
      bool accept = difference <= 0.0;                                  /* :2804, short-circuit: no draw if true */
      if (!accept) accept = static_cast<double>(unit<T>(src.next())) < std::exp(static_cast<double>(-difference / t));
      if (accept) {
        p = ptry; n_acc++;
        if (current_val <= best_val) { best = p; best_val = current_val; n_imp++; }
      }
    }
    if (!cut) iter++;
  }
  if (x) std::memcpy(x, best.data(), d * sizeof(T));
  if (p_out) std::memcpy(p_out, p.data(), d * sizeof(T));
  *f_best = best_val; *n_acc_out = n_acc; *n_imp_out = n_imp; *iters_out = iter; *evals_out = evals;
}

template <class T>
int sann_dispatch(const orc_sann_cfg *c, const void *x0v, const orc_sann_out *out, orc_status *st) {
  if (c->n_chains < 1 || c->dim < 1 || (c->x0_count != 1 && c->x0_count != c->n_chains)) return -1;
  const T *x0 = static_cast<const T *>(x0v);
  const size_t d = c->dim;
  SeqSource seq(c->xs_state);
  T best_f = 0; u64 best_chain = 0, iters = 0, evals_total = 0, draws_total = 0;
  std::vector<T> xb(d), pc(d);
  for (u64 ch = 0; ch < c->n_chains; ch++) {
    const T *start = x0 + (c->x0_count == 1 ? 0 : ch * d);
    T f; uint32_t na, ni; u64 it, ev, draws;
    if (c->rng_mode == ORC_RNG_TAPE) {
      TapeSource src(c->seed, c->chain_offset);
      sann_chain<T>(*c, ch, start, src, xb.data(), pc.data(), &f, &na, &ni, &it, &ev);
      draws = src.total;
    } else {
      const u64 before = seq.total;
      sann_chain<T>(*c, ch, start, seq, xb.data(), pc.data(), &f, &na, &ni, &it, &ev);
      draws = seq.total - before;
    }
    if (ch == 0 || f < best_f) { best_f = f; best_chain = ch; }
    iters = it; evals_total += ev; draws_total += draws;
    if (!out) continue;
    if (out->x_best) std::memcpy(static_cast<T *>(out->x_best) + ch * d, xb.data(), d * sizeof(T));
    if (out->p_cur) std::memcpy(static_cast<T *>(out->p_cur) + ch * d, pc.data(), d * sizeof(T));
    if (out->f_best) static_cast<T *>(out->f_best)[ch] = f;
    if (out->n_accepted) out->n_accepted[ch] = na;
    if (out->n_improved) out->n_improved[ch] = ni;
    if (out->draws) out->draws[ch] = draws;
    if (out->iterations) out->iterations[ch] = it;
    if (out->function_calls) out->function_calls[ch] = ev;
  }
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->f_value = best_f; st->iterations = iters; st->function_calls = evals_total; st->best_index = best_chain;
    st->draws_consumed = draws_total; st->best_valid = 1; st->stop_reason = 1;
  }
  return 0;
}

/* ---------------------------------------------------------------- NelderMeadPSO -------------------------- */
/* One NelderMeadPSO::solve<minimize, false> (nlsolver.h:3615-3920) with the accidents of the reference kept:
 *   - init_solver_state (:3679-3729): the nested loops share `i`, so the body runs once; simplex vertex v (1 <= v < n)
 *     is x with coordinate v raised by `scale` (coordinate 0 is never raised on its own), vertex n is x itself (the
 *     store to [n][n] is out of bounds and ignored here), vertex 0 is x + (1 - sqrt(n + 1)) / n * scale in every
 *     coordinate; PSO particles draw position and velocity interleaved;
 *   - best_val is set once from particle 0 and never updated (:3650), so best_val_no_change only counts iterations
 *     whose best value EQUALS that initial value;
 *   - apply_pso (:3824-3868): `velocity` and `pairwise_best` are by-value copies (only the first declarator of that
 *     line is a reference), so velocities never change; the pair's reference particle is current_order[i + 1] from
 *     the second pair on (the worse of the pair);
 *   - std::sort is unstable: `ties` reports whether two compared values were ever equal. */
template <class T, class Src>
void nmpso_solve(const orc_nmpso_cfg &c, u64 solver, const T *x0, Src &src, T *x_out, T *f_out, u64 *iters_out,
                 u64 *evals_out, uint8_t *ties_out) {
  const size_t n = c.dim, ns = n + 1, np = 2 * n, N = ns + np;
  const T fm = c.minimize ? static_cast<T>(1.0) : static_cast<T>(-1.0);
  const T alpha = static_cast<T>(c.alpha), gamma = static_cast<T>(c.gamma), rho = static_cast<T>(c.rho),
          sigma = static_cast<T>(c.sigma), inertia = static_cast<T>(c.inertia), cog = static_cast<T>(c.cognitive_coef),
          soc = static_cast<T>(c.social_coef), eps = static_cast<T>(c.eps);
  std::vector<T> x(x0, x0 + n), lower(n), upper(n);
  for (size_t i = 0; i < n; i++) {                                     /* minimize(x), :3583-3593 */
    const T temp = std::abs(2.5 * x[i]);
    lower[i] = -temp; upper[i] = temp;
  }
  std::vector<std::vector<T>> pos(N, std::vector<T>(n)), vel(N, std::vector<T>(n, 0.0));
  std::vector<T> val(N);
  u64 evals = 0;
  uint8_t ties = 0;
  pos[0] = x;                                                          /* :3691 */
  if (ns > 1) {                                                        /* :3692-3708, executed once */
    T x_inf_norm = std::abs(x[0]);                                     /* max_abs_vec, :1894-1904 */
    for (size_t i = 1; i < n; i++) { const T t = std::abs(x[i]); if (x_inf_norm < t) x_inf_norm = t; }
    const T a = x_inf_norm < 1.0 ? 1.0 : x_inf_norm;
    const T scale = a < 10 ? a : 10;
    for (size_t i = 1; i < ns; i++) {
      pos[i] = x;
      if (i < n) pos[i][i] = x[i] + scale;                             /* i == n: out of bounds in the reference */
    }
    const T nn = static_cast<T>(n);
    for (size_t i = 0; i < n; i++) pos[0][i] = x[i] + ((1.0 - sqrt(nn + 1.0)) / nn * scale);
  }
  src.begin(0, solver);
  for (size_t i = ns; i < N; i++)                                      /* :3710-3721 */
    for (size_t j = 0; j < n; j++) {
      const T temp = std::abs(upper[j] - lower[j]);
      pos[i][j] = lower[j] + ((upper[j] - lower[j]) * unit<T>(src.next()));
      vel[i][j] = -temp + (unit<T>(src.next()) * temp);
    }
  for (size_t i = 0; i < N; i++) { val[i] = fm * objective<T>(c.objective, pos[i].data(), n); evals++; }   /* :3723-3727 */
  std::vector<T> centroid(n), t_ref(n), t_exp(n), t_con(n);
  std::vector<size_t> order(N);
  for (size_t i = 0; i < N; i++) order[i] = i;
  auto sort_order = [&]() {
    std::sort(order.begin(), order.end(), [&](size_t l, size_t r) { if (l != r && val[l] == val[r]) ties = 1; return val[l] < val[r]; });
  };
  auto transform = [&](const std::vector<T> &point, std::vector<T> &result, T coef, bool reflect) {   /* :1988-2010 */
    for (size_t i = 0; i < n; i++)
      result[i] = reflect ? centroid[i] + coef * (centroid[i] - point[i]) : centroid[i] + coef * (point[i] - centroid[i]);
  };
  u64 iter = 0, no_change = 0;
  const T best_val = val[0];                                           /* :3650, never updated */
  while (true) {
    sort_order();                                                      /* :3654-3658 */
    const bool same = best_val == val[order[0]];
    no_change += same; no_change *= same;                              /* :3663-3664 */
    bool stop = iter >= c.max_iter || no_change >= c.no_change_best_iter;
    if (!stop) {                                                       /* simplex_std_err, :3901-3918 */
      T mean_val = 0, result = 0;
      for (size_t i = 0; i < ns; i++) mean_val += val[order[i]];
      mean_val /= static_cast<T>(ns);
      for (size_t i = 0; i < ns; i++) result += pow(val[order[i]] - mean_val, 2);
      result /= (T)(ns - 1);
      stop = static_cast<T>(sqrt(result)) < eps;
    }
    if (stop) break;
    /* ---- apply_simplex, :3731-3822 */
    {
      const T best_score = val[order[0]];
      const size_t worst = order[ns - 1], second = order[ns - 2];
      std::fill(centroid.begin(), centroid.end(), 0.0);                /* update_centroid, :3869-3885 */
      for (size_t i = 0; i < ns - 1; i++)
        for (size_t j = 0; j < n; j++) centroid[j] += pos[order[i]][j];
      for (auto &v : centroid) v /= static_cast<T>(ns - 1);
      transform(pos[worst], t_ref, alpha, true);
      const T ref_score = fm * objective<T>(c.objective, t_ref.data(), n); evals++;
      if (ref_score >= best_score && ref_score < val[second]) {
        pos[worst] = t_ref; val[worst] = ref_score;
      } else if (ref_score < best_score) {
        transform(t_ref, t_exp, gamma, false);
        const T exp_score = fm * objective<T>(c.objective, t_exp.data(), n); evals++;
        pos[worst] = exp_score < ref_score ? t_exp : t_ref;
        val[worst] = exp_score < ref_score ? exp_score : ref_score;
      } else {
        const T worst_score = val[worst];
        transform(ref_score < worst_score ? t_ref : pos[worst], t_con, rho, false);
        const T cont_score = fm * objective<T>(c.objective, t_con.data(), n); evals++;
        if (cont_score < std::min(ref_score, worst_score)) {
          pos[worst] = t_con; val[worst] = cont_score;
        } else {
          const std::vector<T> &best = pos[order[0]];                  /* shrink, :3886-3900 */
          for (size_t i = 1; i < ns; i++) {
            std::vector<T> &cur = pos[order[i]];
            for (size_t j = 0; j < n; j++) cur[j] = best[j] + sigma * (cur[j] - best[j]);
          }
          for (size_t i = 1; i < ns; i++) val[order[i]] = fm * objective<T>(c.objective, pos[order[i]].data(), n);
          evals += ns - 1;
          sort_order();
        }
      }
    }
    /* ---- apply_pso, :3824-3868 */
    {
      src.begin(iter + 1, solver);
      bool order_flip = false;
      size_t best_in_pair = order[ns];
      const std::vector<T> &best = pos[order[0]];
      for (size_t i = ns; i < N; i++) {
        const size_t id = order[i];
        if (order_flip) best_in_pair = order[i + 1];
        order_flip = static_cast<bool>((i - ns) % 2);
        std::vector<T> &particle = pos[id];
        const std::vector<T> velocity = vel[id], pairwise_best = pos[best_in_pair];   /* copies */
        for (size_t j = 0; j < n; j++) {
          const T r_p = unit<T>(src.next()), r_g = unit<T>(src.next());
          const T temp = (inertia * velocity[j]) + cog * r_p * (pairwise_best[j] - particle[j]) +
                         soc * r_g * (best[j] - particle[j]);
          particle[j] += temp;
        }
        val[id] = fm * objective<T>(c.objective, particle.data(), n); evals++;
      }
    }
    iter++;
  }
  std::memcpy(x_out, pos[order[0]].data(), n * sizeof(T));
  *f_out = val[order[0]]; *iters_out = iter; *evals_out = evals; *ties_out = ties;
}

template <class T>
int nmpso_dispatch(const orc_nmpso_cfg *c, const void *x0v, const orc_nmpso_out *out, orc_status *st) {
  if (c->n_solvers < 1 || c->dim < 2 || (c->x0_count != 1 && c->x0_count != c->n_solvers)) return -1;
  const T *x0 = static_cast<const T *>(x0v);
  const size_t d = c->dim;
  SeqSource seq(c->xs_state);
  T best_f = 0; u64 best_solver = 0, iters = 0, evals_total = 0, draws_total = 0;
  std::vector<T> xb(d);
  for (u64 s = 0; s < c->n_solvers; s++) {
    const T *start = x0 + (c->x0_count == 1 ? 0 : s * d);
    T f; u64 it, ev, draws; uint8_t ties;
    if (c->rng_mode == ORC_RNG_TAPE) {
      TapeSource src(c->seed, c->solver_offset);
      nmpso_solve<T>(*c, s, start, src, xb.data(), &f, &it, &ev, &ties);
      draws = src.total;
    } else {
      const u64 before = seq.total;
      nmpso_solve<T>(*c, s, start, seq, xb.data(), &f, &it, &ev, &ties);
      draws = seq.total - before;
    }
    if (s == 0 || f < best_f) { best_f = f; best_solver = s; }
    iters = it; evals_total += ev; draws_total += draws;
    if (!out) continue;
    if (out->x_best) std::memcpy(static_cast<T *>(out->x_best) + s * d, xb.data(), d * sizeof(T));
    if (out->f_best) static_cast<T *>(out->f_best)[s] = f;
    if (out->iterations) out->iterations[s] = it;
    if (out->function_calls) out->function_calls[s] = ev;
    if (out->draws) out->draws[s] = draws;
    if (out->ties) out->ties[s] = ties;
  }
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->f_value = best_f; st->iterations = iters; st->function_calls = evals_total; st->best_index = best_solver;
    st->draws_consumed = draws_total; st->best_valid = 1;
  }
  return 0;
}

template <class T>
int de_dispatch(const orc_de_cfg *c, const void *x0, const orc_de_out *out, orc_status *st) {
  if (c->pop_size < 4 || c->dim < 1) return -1;     /* the reference loops forever for pop < 4 (:2344-2354) */
  if (c->rng_mode == ORC_RNG_TAPE) { TapeSource s(c->seed, c->agent_offset); de_run<T>(*c, static_cast<const T *>(x0), s, out, st); }
  else { SeqSource s(c->xs_state); de_run<T>(*c, static_cast<const T *>(x0), s, out, st); }
  return 0;
}
template <class T>
int pso_dispatch(const orc_pso_cfg *c, const void *lo, const void *up, const orc_pso_out *out, orc_status *st) {
  if (c->n_particles < 1 || c->dim < 1) return -1;
  if (c->pso_type == ORC_PSO_VANILLA && !c->social_index_j && c->n_particles > c->dim) return -2;  /* reference UB */
  if (c->rng_mode == ORC_RNG_TAPE) { TapeSource s(c->seed, c->particle_offset); pso_run<T>(*c, static_cast<const T *>(lo), static_cast<const T *>(up), s, out, st); }
  else { SeqSource s(c->xs_state); pso_run<T>(*c, static_cast<const T *>(lo), static_cast<const T *>(up), s, out, st); }
  return 0;
}

}  // namespace

extern "C" {

int oracle_de_run(const orc_de_cfg *c, const void *x0, const orc_de_out *out, orc_status *st) {
  return c->dtype == ORC_F64 ? de_dispatch<double>(c, x0, out, st) : de_dispatch<float>(c, x0, out, st);
}
int oracle_pso_run(const orc_pso_cfg *c, const void *lower, const void *upper, const orc_pso_out *out,
                   orc_status *st) {
  return c->dtype == ORC_F64 ? pso_dispatch<double>(c, lower, upper, out, st)
                             : pso_dispatch<float>(c, lower, upper, out, st);
}
/* a batch of SANN chains; the batch status reports the best chain (lowest value, lowest chain index on ties) */
int oracle_nmpso_run(const orc_nmpso_cfg *c, const void *x0, const orc_nmpso_out *out, orc_status *st) {
  return c->dtype == ORC_F64 ? nmpso_dispatch<double>(c, x0, out, st) : nmpso_dispatch<float>(c, x0, out, st);
}
int oracle_sann_run(const orc_sann_cfg *c, const void *x0, const orc_sann_out *out, orc_status *st) {
  return c->dtype == ORC_F64 ? sann_dispatch<double>(c, x0, out, st) : sann_dispatch<float>(c, x0, out, st);
}
/* PSO::minimize(x) without bounds derives lower = -|x|, upper = |x| (nlsolver.h:2553-2563); callers do that. */

/* ---- stepwise handles (fp64 tape only): what the world_size-2 CPU tests drive through nlsolver_b200.distributed ---- */
typedef DELoop<double, TapeSource> DEHandle;
typedef PSOLoop<double, TapeSource> PSOHandle;

void *oracle_de_open(const orc_de_cfg *c, const double *x0) {
  if (c->dtype != ORC_F64 || c->rng_mode != ORC_RNG_TAPE || c->pop_size < 4 || c->dim < 1) return nullptr;
  DEHandle *h = new DEHandle(*c, x0, TapeSource(c->seed, c->agent_offset), false);
  h->scan();
  return h;
}
/* n generations, each followed by the next loop-top scan; stops early when a stop rule has fired */
void oracle_de_advance(void *p, uint64_t n) {
  DEHandle *h = static_cast<DEHandle *>(p);
  for (uint64_t g = 0; g < n && !h->stop_reason; g++) { h->generation(); h->scan(); }
}
void oracle_de_report(void *p, const orc_de_out *out, orc_status *st) { static_cast<DEHandle *>(p)->report(out, st); }
void oracle_de_export_top(void *p, uint64_t k, double *rows, double *scores) { static_cast<DEHandle *>(p)->export_top(k, rows, scores); }
void oracle_de_import_migrants(void *p, uint64_t k, const double *rows, const double *scores) { static_cast<DEHandle *>(p)->import_migrants(k, rows, scores); }
void oracle_de_close(void *p) { delete static_cast<DEHandle *>(p); }

void *oracle_pso_open(const orc_pso_cfg *c, const double *lower, const double *upper) {
  if (c->dtype != ORC_F64 || c->rng_mode != ORC_RNG_TAPE) return nullptr;
  return new PSOHandle(*c, lower, upper, TapeSource(c->seed, c->particle_offset));
}
/* evaluate this shard; returns 1 and the candidate (value, local index) if it has one */
int oracle_pso_evaluate(void *p, double *value, uint64_t *local_index) {
  size_t i = 0;
  const bool any = static_cast<PSOHandle *>(p)->evaluate(value, &i);
  *local_index = i;
  return any;
}
const double *oracle_pso_row(void *p, uint64_t local_index) { PSOHandle *h = static_cast<PSOHandle *>(p); return &h->X[local_index * h->d]; }
const double *oracle_pso_pbest(void *p) { return static_cast<PSOHandle *>(p)->pbest.data(); }
/* swarm-level update with the winner over all shards, then iter++ (unless initial) and the stop test */
int oracle_pso_adopt(void *p, int have, double value, uint64_t global_index, const double *row, uint64_t n_global,
                     double std_err_all, int initial) {
  PSOHandle *h = static_cast<PSOHandle *>(p);
  h->adopt(have != 0, value, global_index, row, n_global);
  if (!initial) h->iter++;
  return h->stop_test(std_err_all);
}
void oracle_pso_move(void *p) { static_cast<PSOHandle *>(p)->move(); }
void oracle_pso_report(void *p, const orc_pso_out *out, orc_status *st) { static_cast<PSOHandle *>(p)->report(out, st); }
void oracle_pso_close(void *p) { delete static_cast<PSOHandle *>(p); }

void oracle_set_custom_objective(int pairwise, double lane0_seed, orc_term_fn term, orc_finish_fn finish) {
  g_custom.pairwise = pairwise; g_custom.seed = lane0_seed; g_custom.term = term; g_custom.finish = finish;
  g_custom.full = nullptr;
}
void oracle_set_custom_full(orc_full_fn full) { g_custom.full = full; }

double oracle_objective(int dtype, int id, const void *x, uint64_t d) {
  return dtype == ORC_F64 ? objective<double>(id, static_cast<const double *>(x), d)
                          : static_cast<double>(objective<float>(id, static_cast<const float *>(x), d));
}
uint64_t oracle_tape_key(uint64_t seed, uint64_t gen, uint64_t agent) { return tape_key(seed, gen, agent); }
uint64_t oracle_tape_draw(uint64_t key, uint64_t k) { return tape_draw(key, k); }
double oracle_unit_f64(uint64_t u) { return unit<double>(u); }
float oracle_unit_f32(uint64_t u) { return unit<float>(u); }
/* xorshift128+ known-answer access: writes the default state and the first n raw sums t+s */
void oracle_xorshift_default(uint64_t state[2], uint64_t *raw, uint64_t n) {
  XorShift g; state[0] = g.a; state[1] = g.b;
  for (uint64_t i = 0; i < n; i++) raw[i] = g.next();
}
double oracle_std_err_f64(const double *x, uint64_t n) { return std_err(std::vector<double>(x, x + n)); }

}  /* extern "C" */
