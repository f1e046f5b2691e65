/*
 * ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference templates nlsolver::DE / nlsolver::PSO (#included from /root/reference where
 * they lie; nothing of the reference is copied into this repo) so that
 *   (a) oracle/popsolve_oracle.cpp (the restatement) can be validated bit-for-bit against the real thing, and
 *   (b) bench.py can time the reference's own CPU path (`cpu_baseline.kind = "reference"`, `--impl reference`).
 * Built by oracle/Makefile into oracle/_ref/libnls_ref.so (git-ignored, travels to the GPU box as a binary).
 *
 * How the reference is fed: `RNG` and `Callable` are template parameters held by reference
 * (nlsolver.h:2383-2384, 2503-2504), so a replay generator (TapeRNG) and a notifying objective (Hook) need no
 * change to the reference.  DE calls f exactly once per agent after that agent's draws (nlsolver.h:2463), which is
 * what lets the tape switch to the next (generation, agent) stream.  Decisions (donors, dim, mask, accept) are
 * reconstructed from the draws handed out and checked against the trial vector the reference actually built.
 */
#include <chrono>
#include <cstring>
#include <vector>

#include "nlsolver.h"        /* -I/root/reference */
#include "test_functions.h"  /* the reference's 2-D objectives, for ref_objective_2d only */

#include "oracle_abi.h"

extern "C" double oracle_objective(int dtype, int id, const void *x, uint64_t d);
extern "C" uint64_t oracle_tape_key(uint64_t seed, uint64_t gen, uint64_t agent);
extern "C" uint64_t oracle_tape_draw(uint64_t key, uint64_t k);

namespace {
typedef uint64_t u64;
thread_local u64 g_inconsistencies = 0;   /* reconstructed-decision vs. actual-trial mismatches of the last DE run */

template <class T> constexpr int dtype_of() { return sizeof(T) == 8 ? ORC_F64 : ORC_F32; }
template <class T> T to_unit(u64 u) { return static_cast<T>(u / static_cast<T>(18446744073709551615U)); }

/* plain N-D objective functor (the same N-D forms the device uses), for the timing runs */
template <class T>
struct PlainObjective {
  int id;
  T operator()(std::vector<T> &x) { return static_cast<T>(oracle_objective(dtype_of<T>(), id, x.data(), x.size())); }
};

/* ------------------------------------------------------------------ DE probe ----------------------------- */
template <class T>
struct DEProbe {
  const orc_de_cfg &c;
  const size_t P, d;
  bool tape;
  /* tape position */
  bool init_phase = true;
  size_t init_draws = 0, init_evals = 0, agent = 0;
  u64 gen = 0, key = 0, k = 0, total = 0;
  std::vector<u64> cur;   /* raw draws of the agent in flight */
  /* mirror of what the reference holds privately */
  std::vector<T> A, scores, tscores;
  size_t best_id = 0;
  std::vector<uint32_t> donors, dimv, rej;
  std::vector<uint8_t> acc, masks;
  size_t inconsistencies = 0;

  DEProbe(const orc_de_cfg &cfg, bool use_tape)
      : c(cfg), P(cfg.pop_size), d(cfg.dim), tape(use_tape), A(P * d), scores(P), tscores(P), donors(P * 3), dimv(P),
        rej(P), acc(P), masks(P * d) {}

  T draw() {
    total++;
    u64 raw;
    if (init_phase) {
      const size_t a = init_draws / d, j = init_draws % d;
      init_draws++;
      raw = oracle_tape_draw(oracle_tape_key(c.seed, 0, c.agent_offset + a), j);
    } else {
      if (k == 0) key = oracle_tape_key(c.seed, gen, c.agent_offset + agent);
      raw = oracle_tape_draw(key, k++);
      cur.push_back(raw);
    }
    return to_unit<T>(raw);
  }
  void scan_best() {   /* harness-level mirror of the best scan, needed only for the `best` exclusion index */
    for (size_t i = 0; i < P; i++) if (scores[i] < scores[best_id]) best_id = i;
  }
  void evaluated(const std::vector<T> &x, T raw_value) {
    const T fm = c.minimize ? static_cast<T>(1.0) : static_cast<T>(-1.0);
    const T s = fm * raw_value;
    if (init_phase) {
      std::memcpy(&A[init_evals * d], x.data(), d * sizeof(T));
      scores[init_evals] = s;
      if (++init_evals == P) { init_phase = false; gen = 1; agent = 0; k = 0; scan_best(); }
      return;
    }
    const size_t i = agent;
    if (tape) reconstruct(i, x);
    tscores[i] = s;
    acc[i] = s < scores[i];
    if (acc[i]) { std::memcpy(&A[i * d], x.data(), d * sizeof(T)); scores[i] = s; }
    cur.clear(); k = 0;
    if (++agent == P) { agent = 0; gen++; scan_best(); }
  }
  void reconstruct(size_t i, const std::vector<T> &trial) {
    const size_t n_idx = cur.size() - 1 - d;
    const size_t fixed = (c.strategy == ORC_DE_RANDOM) ? i : best_id;
    size_t ids[4] = {fixed, 0, 0, 0};
    uint32_t n = 1, rejected = 0;
    for (size_t q = 0; q < n_idx && n < 4; q++) {
      const size_t prop = static_cast<size_t>(to_unit<T>(cur[q]) * P);
      bool used = false;
      for (uint32_t r = 0; r < n; r++) used |= ids[r] == prop;
      if (used) rejected++; else ids[n++] = prop;
    }
    if (n != 4 || rejected + 3 != n_idx) inconsistencies++;
    const size_t dim = static_cast<size_t>(to_unit<T>(cur[n_idx]) * d);
    const T CR = static_cast<T>(c.crossover_prob), F = static_cast<T>(c.differential_weight);
    for (size_t j = 0; j < d; j++) {
      const bool mut = to_unit<T>(cur[n_idx + 1 + j]) < CR || j == dim;
      masks[i * d + j] = mut;
      const T expect = mut ? A[ids[1] * d + j] + F * (A[ids[2] * d + j] - A[ids[3] * d + j]) : A[ids[0] * d + j];
      if (std::memcmp(&expect, &trial[j], sizeof(T)) != 0) inconsistencies++;
    }
    donors[i * 3] = ids[1]; donors[i * 3 + 1] = ids[2]; donors[i * 3 + 2] = ids[3];
    dimv[i] = dim; rej[i] = rejected;
  }
};

template <class T> struct DETapeRNG { DEProbe<T> &p; T operator()() { return p.draw(); } };
template <class T> struct DEHook {
  DEProbe<T> &p;
  T operator()(std::vector<T> &x) {
    const T v = static_cast<T>(oracle_objective(dtype_of<T>(), p.c.objective, x.data(), x.size()));
    p.evaluated(x, v);
    return v;
  }
};

template <class T, class RNG>
nlsolver::solver_status<T> de_solve(const orc_de_cfg &c, DEHook<T> &f, RNG &g, std::vector<T> &x) {
  const T CR = static_cast<T>(c.crossover_prob), F = static_cast<T>(c.differential_weight), eps = static_cast<T>(c.eps);
  if (c.strategy == ORC_DE_RANDOM) {
    nlsolver::DE<DEHook<T>, RNG, T, nlsolver::RecombinationStrategy::random> s(f, g, CR, F, eps, c.pop_size,
                                                                              c.max_iter, c.best_val_no_change);
    return c.minimize ? s.minimize(x) : s.maximize(x);
  }
  nlsolver::DE<DEHook<T>, RNG, T, nlsolver::RecombinationStrategy::best> s(f, g, CR, F, eps, c.pop_size, c.max_iter,
                                                                          c.best_val_no_change);
  return c.minimize ? s.minimize(x) : s.maximize(x);
}

template <class T>
int de_run(const orc_de_cfg *c, const void *x0, const orc_de_out *out, orc_status *st) {
  if (c->pop_size < 4 || c->dim < 1) return -1;
  const bool tape = c->rng_mode == ORC_RNG_TAPE;
  DEProbe<T> probe(*c, tape);
  DEHook<T> hook{probe};
  std::vector<T> x(static_cast<const T *>(x0), static_cast<const T *>(x0) + c->dim);
  u64 draws = 0;
  auto status = [&]() {
    if (tape) { DETapeRNG<T> g{probe}; auto r = de_solve<T>(*c, hook, g, x); draws = probe.total; return r; }
    nlsolver::rng::xorshift<T> g;
    if (c->xs_state[0] | c->xs_state[1]) g.set_state(c->xs_state[0], c->xs_state[1]);
    return de_solve<T>(*c, hook, g, x);
  }();
  const auto sum = status.get_summary();
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->f_value = std::get<2>(sum); st->iterations = std::get<1>(sum); st->function_calls = std::get<0>(sum);
    st->best_index = probe.best_id; st->draws_consumed = draws; st->best_valid = 1;
  }
  g_inconsistencies = probe.inconsistencies;
  if (!out) return 0;
  const size_t P = c->pop_size, d = c->dim;
  if (out->x_best) std::memcpy(out->x_best, x.data(), d * sizeof(T));
  if (out->rows) std::memcpy(out->rows, probe.A.data(), P * d * sizeof(T));
  if (out->scores) std::memcpy(out->scores, probe.scores.data(), P * sizeof(T));
  if (out->trial_scores) std::memcpy(out->trial_scores, probe.tscores.data(), P * sizeof(T));
  if (out->donors) std::memcpy(out->donors, probe.donors.data(), P * 3 * sizeof(uint32_t));
  if (out->dim_idx) std::memcpy(out->dim_idx, probe.dimv.data(), P * sizeof(uint32_t));
  if (out->rejects) std::memcpy(out->rejects, probe.rej.data(), P * sizeof(uint32_t));
  if (out->accepted) std::memcpy(out->accepted, probe.acc.data(), P);
  if (out->masks) std::memcpy(out->masks, probe.masks.data(), P * d);
  return 0;
}

/* ------------------------------------------------------------------ PSO probe ---------------------------- */
template <class T>
struct PSOProbe {
  const orc_pso_cfg &c;
  const size_t P, d, init_per_particle;
  u64 n = 0;       /* draws handed out */
  size_t evals = 0;
  std::vector<T> X, last, pbest;
  PSOProbe(const orc_pso_cfg &cfg)
      : c(cfg), P(cfg.n_particles), d(cfg.dim), init_per_particle(cfg.pso_type == ORC_PSO_VANILLA ? 2 * d : d),
        X(P * d), last(P), pbest(P, static_cast<T>(10000)) {}
  T draw() {
    /* fixed draw counts: init d or 2d per particle (nlsolver.h:2642-2653), then 2d per particle per generation */
    const u64 init_total = static_cast<u64>(P) * init_per_particle;
    u64 gen, i, k;
    if (n < init_total) { gen = 0; i = n / init_per_particle; k = n % init_per_particle; }
    else { const u64 m = n - init_total, per_gen = static_cast<u64>(P) * 2 * d; gen = 1 + m / per_gen; i = (m % per_gen) / (2 * d); k = m % (2 * d); }
    n++;
    return to_unit<T>(oracle_tape_draw(oracle_tape_key(c.seed, gen, c.particle_offset + i), k));
  }
  void evaluated(const std::vector<T> &x, T raw) {
    const size_t i = evals++ % P;
    const T s = (c.minimize ? static_cast<T>(1.0) : static_cast<T>(-1.0)) * raw;
    std::memcpy(&X[i * d], x.data(), d * sizeof(T));
    last[i] = s;
    if (s < pbest[i]) pbest[i] = s;
  }
};
template <class T> struct PSOTapeRNG { PSOProbe<T> &p; T operator()() { return p.draw(); } };
template <class T> struct PSOHook {
  PSOProbe<T> &p;
  T operator()(std::vector<T> &x) {
    const T v = static_cast<T>(oracle_objective(dtype_of<T>(), p.c.objective, x.data(), x.size()));
    p.evaluated(x, v);
    return v;
  }
};

template <class T, class RNG, nlsolver::PSOType Type>
nlsolver::solver_status<T> pso_solve_t(const orc_pso_cfg &c, PSOHook<T> &f, RNG &g, std::vector<T> &x,
                                       const std::vector<T> &lo, const std::vector<T> &up) {
  nlsolver::PSO<PSOHook<T>, RNG, T, Type> s(f, g, static_cast<T>(c.inertia), static_cast<T>(c.cognitive_coef),
                                            static_cast<T>(c.social_coef), c.n_particles, c.max_iter,
                                            c.best_val_no_change, static_cast<T>(c.eps));
  if (c.constrained) return c.minimize ? s.minimize(x, lo, up) : s.maximize(x, lo, up);
  x = up;   /* unbounded overloads derive lower = -|x|, upper = |x| from x (nlsolver.h:2553-2575) */
  return c.minimize ? s.minimize(x) : s.maximize(x);
}

template <class T>
int pso_run(const orc_pso_cfg *c, const void *lower, const void *upper, const orc_pso_out *out, orc_status *st) {
  if (c->n_particles < 1 || c->dim < 1) return -1;
  /* the reference reads swarm_best_position[i] out of bounds when P > d (nlsolver.h:2674) — not runnable */
  if (c->pso_type == ORC_PSO_VANILLA && c->n_particles > c->dim) return -2;
  if (c->pso_type == ORC_PSO_VANILLA && c->social_index_j) return -3;   /* the reference has no corrected mode */
  const size_t d = c->dim;
  PSOProbe<T> probe(*c);
  PSOHook<T> hook{probe};
  std::vector<T> lo(static_cast<const T *>(lower), static_cast<const T *>(lower) + d);
  std::vector<T> up(static_cast<const T *>(upper), static_cast<const T *>(upper) + d), x(d);
  const bool tape = c->rng_mode == ORC_RNG_TAPE;
  auto run = [&](auto &g) {
    return c->pso_type == ORC_PSO_VANILLA
               ? pso_solve_t<T, std::remove_reference_t<decltype(g)>, nlsolver::PSOType::Vanilla>(*c, hook, g, x, lo, up)
               : pso_solve_t<T, std::remove_reference_t<decltype(g)>, nlsolver::PSOType::Accelerated>(*c, hook, g, x, lo, up);
  };
  auto status = [&]() {
    if (tape) { PSOTapeRNG<T> g{probe}; return run(g); }
    nlsolver::rng::xorshift<T> g;
    if (c->xs_state[0] | c->xs_state[1]) g.set_state(c->xs_state[0], c->xs_state[1]);
    return run(g);
  }();
  const auto sum = status.get_summary();
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->f_value = std::get<2>(sum); st->iterations = std::get<1>(sum); st->function_calls = std::get<0>(sum);
    st->draws_consumed = probe.n; st->best_valid = x.size() == d;
  }
  if (!out) return 0;
  if (out->x_best && x.size() == d) std::memcpy(out->x_best, x.data(), d * sizeof(T));
  if (out->positions) std::memcpy(out->positions, probe.X.data(), probe.X.size() * sizeof(T));
  if (out->pbest_values) std::memcpy(out->pbest_values, probe.pbest.data(), probe.pbest.size() * sizeof(T));
  if (out->last_values) std::memcpy(out->last_values, probe.last.data(), probe.last.size() * sizeof(T));
  return 0;
}

/* ------------------------------------------------------------------ SANN probe --------------------------- */
/* nlsolver::SANN (nlsolver.h:2744-2815) is one chain; a batch is that solver run once per chain.  The tape opens a new
 * epoch at every objective call (oracle_abi.h), which needs no knowledge of whether the Metropolis draw happened. */
template <class T>
struct SANNProbe {
  const orc_sann_cfg &c;
  u64 chain = 0, calls = 0, k = 0, total = 0;
  explicit SANNProbe(const orc_sann_cfg &cfg) : c(cfg) {}
  T draw() {
    total++;
    return to_unit<T>(oracle_tape_draw(oracle_tape_key(c.seed, calls - 1, c.chain_offset + chain), k++));
  }
  void evaluated() { calls++; k = 0; }
};
template <class T> struct SANNTapeRNG { SANNProbe<T> &p; T operator()() { return p.draw(); } };
template <class T> struct SANNHook {
  SANNProbe<T> &p;
  T operator()(std::vector<T> &x) {
    const T v = static_cast<T>(oracle_objective(dtype_of<T>(), p.c.objective, x.data(), x.size()));
    p.evaluated();
    return v;
  }
};

template <class T>
int sann_run(const orc_sann_cfg *c, const void *x0v, const orc_sann_out *out, orc_status *st) {
  if (c->n_chains < 1 || c->dim < 1 || (c->x0_count != 1 && c->x0_count != c->n_chains)) return -1;
  if (c->max_steps) return -3;   /* the reference cannot stop inside a solve */
  const size_t d = c->dim;
  const T *x0 = static_cast<const T *>(x0v);
  const bool tape = c->rng_mode == ORC_RNG_TAPE;
  nlsolver::rng::xorshift<T> seq;
  if (c->xs_state[0] | c->xs_state[1]) seq.set_state(c->xs_state[0], c->xs_state[1]);
  T best_f = 0; u64 best_chain = 0, evals_total = 0, draws_total = 0, iters = 0;
  for (u64 ch = 0; ch < c->n_chains; ch++) {
    const T *start = x0 + (c->x0_count == 1 ? 0 : ch * d);
    std::vector<T> x(start, start + d);
    SANNProbe<T> probe(*c);
    probe.chain = ch;
    SANNHook<T> hook{probe};
    auto status = [&]() {
      if (tape) {
        SANNTapeRNG<T> g{probe};
        nlsolver::SANN<SANNHook<T>, SANNTapeRNG<T>, T> s(hook, g, c->max_iter, c->temperature_iter,
                                                        static_cast<T>(c->temperature_max));
        return c->minimize ? s.minimize(x) : s.maximize(x);
      }
      nlsolver::SANN<SANNHook<T>, nlsolver::rng::xorshift<T>, T> s(hook, seq, c->max_iter, c->temperature_iter,
                                                                  static_cast<T>(c->temperature_max));
      return c->minimize ? s.minimize(x) : s.maximize(x);
    }();
    const auto sum = status.get_summary();
    const T f = std::get<2>(sum);
    if (ch == 0 || f < best_f) { best_f = f; best_chain = ch; }
    iters = std::get<1>(sum); evals_total += std::get<0>(sum); draws_total += probe.total;
    if (!out) continue;
    if (out->x_best) std::memcpy(static_cast<T *>(out->x_best) + ch * d, x.data(), d * sizeof(T));
    if (out->f_best) static_cast<T *>(out->f_best)[ch] = f;
    if (out->draws) out->draws[ch] = probe.total;
    if (out->iterations) out->iterations[ch] = std::get<1>(sum);
    if (out->function_calls) out->function_calls[ch] = std::get<0>(sum);
  }
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->f_value = best_f; st->iterations = iters; st->function_calls = evals_total; st->best_index = best_chain;
    st->draws_consumed = draws_total; st->best_valid = 1; st->stop_reason = 1;
  }
  return 0;
}

/* the reference SANN, its own xorshift<T>, a plain N-D objective: n_chains solves one after the other */
template <class T>
int sann_time(const orc_sann_cfg *c, const void *x0v, double *seconds, orc_status *st) {
  if (c->n_chains < 1 || c->dim < 1 || (c->x0_count != 1 && c->x0_count != c->n_chains)) return -1;
  const size_t d = c->dim;
  const T *x0 = static_cast<const T *>(x0v);
  PlainObjective<T> f{c->objective};
  nlsolver::rng::xorshift<T> g;
  T best_f = 0; u64 best_chain = 0, evals_total = 0, iters = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (u64 ch = 0; ch < c->n_chains; ch++) {
    const T *start = x0 + (c->x0_count == 1 ? 0 : ch * d);
    std::vector<T> x(start, start + d);
    nlsolver::SANN<PlainObjective<T>, nlsolver::rng::xorshift<T>, T> s(f, g, c->max_iter, c->temperature_iter,
                                                                      static_cast<T>(c->temperature_max));
    const auto sum = (c->minimize ? s.minimize(x) : s.maximize(x)).get_summary();
    if (ch == 0 || std::get<2>(sum) < best_f) { best_f = std::get<2>(sum); best_chain = ch; }
    iters = std::get<1>(sum); evals_total += std::get<0>(sum);
  }
  *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (st) { std::memset(st, 0, sizeof(*st)); st->f_value = best_f; st->iterations = iters; st->function_calls = evals_total; st->best_index = best_chain; }
  return 0;
}

/* ------------------------------------------------------------------ timing runs -------------------------- */
template <class T>
int de_time(const orc_de_cfg *c, const void *x0, double *seconds, orc_status *st) {
  PlainObjective<T> f{c->objective};
  nlsolver::rng::xorshift<T> g;
  std::vector<T> x(static_cast<const T *>(x0), static_cast<const T *>(x0) + c->dim);
  const T CR = static_cast<T>(c->crossover_prob), F = static_cast<T>(c->differential_weight), eps = static_cast<T>(c->eps);
  const auto t0 = std::chrono::steady_clock::now();
  auto run = [&]() {
    if (c->strategy == ORC_DE_RANDOM) {
      nlsolver::DE<PlainObjective<T>, nlsolver::rng::xorshift<T>, T, nlsolver::RecombinationStrategy::random> s(
          f, g, CR, F, eps, c->pop_size, c->max_iter, c->best_val_no_change);
      return c->minimize ? s.minimize(x) : s.maximize(x);
    }
    nlsolver::DE<PlainObjective<T>, nlsolver::rng::xorshift<T>, T, nlsolver::RecombinationStrategy::best> s(
        f, g, CR, F, eps, c->pop_size, c->max_iter, c->best_val_no_change);
    return c->minimize ? s.minimize(x) : s.maximize(x);
  };
  const auto status = run();
  *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const auto sum = status.get_summary();
  if (st) { std::memset(st, 0, sizeof(*st)); st->f_value = std::get<2>(sum); st->iterations = std::get<1>(sum); st->function_calls = std::get<0>(sum); }
  return 0;
}
template <class T>
int pso_time(const orc_pso_cfg *c, const void *upper, double *seconds, orc_status *st) {
  if (c->pso_type == ORC_PSO_VANILLA && c->n_particles > c->dim) return -2;
  PlainObjective<T> f{c->objective};
  nlsolver::rng::xorshift<T> g;
  std::vector<T> x(static_cast<const T *>(upper), static_cast<const T *>(upper) + c->dim);
  const auto t0 = std::chrono::steady_clock::now();
  auto run = [&]() {
    if (c->pso_type == ORC_PSO_VANILLA) {
      nlsolver::PSO<PlainObjective<T>, nlsolver::rng::xorshift<T>, T, nlsolver::PSOType::Vanilla> s(
          f, g, static_cast<T>(c->inertia), static_cast<T>(c->cognitive_coef), static_cast<T>(c->social_coef),
          c->n_particles, c->max_iter, c->best_val_no_change, static_cast<T>(c->eps));
      return c->minimize ? s.minimize(x) : s.maximize(x);
    }
    nlsolver::PSO<PlainObjective<T>, nlsolver::rng::xorshift<T>, T, nlsolver::PSOType::Accelerated> s(
        f, g, static_cast<T>(c->inertia), static_cast<T>(c->cognitive_coef), static_cast<T>(c->social_coef),
        c->n_particles, c->max_iter, c->best_val_no_change, static_cast<T>(c->eps));
    return c->minimize ? s.minimize(x) : s.maximize(x);
  };
  const auto status = run();
  *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const auto sum = status.get_summary();
  if (st) { std::memset(st, 0, sizeof(*st)); st->f_value = std::get<2>(sum); st->iterations = std::get<1>(sum); st->function_calls = std::get<0>(sum); }
  return 0;
}

/* ------------------------------------------------------------------ NelderMeadPSO ------------------------ */
/* nlsolver::NelderMeadPSO (nlsolver.h:3546-3920), one solver per start point.  Both draw phases consume exactly
 * 4 n^2 draws (init_solver_state, then every apply_pso), so the tape position follows from the draw count. */
template <class T>
struct NMPSOTapeRNG {
  const orc_nmpso_cfg &c;
  u64 solver, count = 0;
  T operator()() {
    const u64 per = 4 * c.dim * c.dim, tag = count / per, k = count % per;
    count++;
    return to_unit<T>(oracle_tape_draw(oracle_tape_key(c.seed, tag, c.solver_offset + solver), k));
  }
};
template <class T>
int nmpso_run(const orc_nmpso_cfg *c, const void *x0v, const orc_nmpso_out *out, orc_status *st) {
  /* The reference stores one element past the end of a vector<T>(n) (nlsolver.h:3697-3700).  glibc hands out chunks of
   * max(32, (bytes + 8 + 15) & ~15) bytes of which all but 8 are usable: the stray store is harmless iff it still fits
   * (fp64: even n; fp32: n = 2, 4, 8, 12, 16, ...), otherwise it overwrites the next chunk's size field and the process
   * aborts in free().  The harness refuses the shapes that would corrupt its own heap. */
  {
    const size_t bytes = c->dim * sizeof(T);
    size_t chunk = (bytes + 8 + 15) & ~size_t(15);
    if (chunk < 32) chunk = 32;
    if (c->n_solvers < 1 || c->dim < 2 || chunk - 8 < bytes + sizeof(T)) return -1;
  }
  const T *x0 = static_cast<const T *>(x0v);
  const size_t d = c->dim;
  PlainObjective<T> f{c->objective};
  nlsolver::rng::xorshift<T> seq;
  if (c->xs_state[0] | c->xs_state[1]) seq.set_state(c->xs_state[0], c->xs_state[1]);
  T best_f = 0; u64 best_solver = 0, iters = 0, evals_total = 0, draws_total = 0;
  for (u64 s = 0; s < c->n_solvers; s++) {
    std::vector<T> x(x0 + (c->x0_count == 1 ? 0 : s * d), x0 + (c->x0_count == 1 ? 0 : s * d) + d);
    u64 draws = 0;
    auto run = [&](auto &gen) {
      nlsolver::NelderMeadPSO<PlainObjective<T>, std::remove_reference_t<decltype(gen)>, T> solver(
          f, gen, c->alpha, c->gamma, c->rho, c->sigma, c->inertia, c->cognitive_coef, c->social_coef, c->eps,
          c->max_iter, c->no_change_best_iter);
      return c->minimize ? solver.minimize(x) : solver.maximize(x);
    };
    T fv; u64 it, ev;
    if (c->rng_mode == ORC_RNG_TAPE) {
      NMPSOTapeRNG<T> g{*c, s};
      const auto sum = run(g).get_summary();
      fv = std::get<2>(sum); it = std::get<1>(sum); ev = std::get<0>(sum); draws = g.count;
    } else {
      const auto sum = run(seq).get_summary();
      fv = std::get<2>(sum); it = std::get<1>(sum); ev = std::get<0>(sum);
    }
    if (s == 0 || fv < best_f) { best_f = fv; best_solver = s; }
    iters = it; evals_total += ev; draws_total += draws;
    if (!out) continue;
    if (out->x_best) std::memcpy(static_cast<T *>(out->x_best) + s * d, x.data(), d * sizeof(T));
    if (out->f_best) static_cast<T *>(out->f_best)[s] = fv;
    if (out->iterations) out->iterations[s] = it;
    if (out->function_calls) out->function_calls[s] = ev;
    if (out->draws) out->draws[s] = draws;
    if (out->ties) out->ties[s] = 0;
  }
  if (st) {
    std::memset(st, 0, sizeof(*st));
    st->f_value = best_f; st->iterations = iters; st->function_calls = evals_total; st->best_index = best_solver;
    st->draws_consumed = draws_total; st->best_valid = 1;
  }
  return 0;
}

}  // namespace

extern "C" {

int ref_de_run(const orc_de_cfg *c, const void *x0, const orc_de_out *out, orc_status *st) {
  return c->dtype == ORC_F64 ? de_run<double>(c, x0, out, st) : de_run<float>(c, x0, out, st);
}
int ref_pso_run(const orc_pso_cfg *c, const void *lower, const void *upper, const orc_pso_out *out, orc_status *st) {
  return c->dtype == ORC_F64 ? pso_run<double>(c, lower, upper, out, st) : pso_run<float>(c, lower, upper, out, st);
}
int ref_sann_run(const orc_sann_cfg *c, const void *x0, const orc_sann_out *out, orc_status *st) {
  return c->dtype == ORC_F64 ? sann_run<double>(c, x0, out, st) : sann_run<float>(c, x0, out, st);
}
int ref_nmpso_run(const orc_nmpso_cfg *c, const void *x0, const orc_nmpso_out *out, orc_status *st) {
  return c->dtype == ORC_F64 ? nmpso_run<double>(c, x0, out, st) : nmpso_run<float>(c, x0, out, st);
}
int ref_sann_time(const orc_sann_cfg *c, const void *x0, double *seconds, orc_status *st) {
  return c->dtype == ORC_F64 ? sann_time<double>(c, x0, seconds, st) : sann_time<float>(c, x0, seconds, st);
}
/* the reference solver, its own xorshift<T>, a plain N-D objective, steady_clock around minimize() */
int ref_de_time(const orc_de_cfg *c, const void *x0, double *seconds, orc_status *st) {
  return c->dtype == ORC_F64 ? de_time<double>(c, x0, seconds, st) : de_time<float>(c, x0, seconds, st);
}
int ref_pso_time(const orc_pso_cfg *c, const void *upper, double *seconds, orc_status *st) {
  return c->dtype == ORC_F64 ? pso_time<double>(c, upper, seconds, st) : pso_time<float>(c, upper, seconds, st);
}
/* the reference's own 2-D objectives (test_functions.h:51-92), double */
/* every objective the reference's test driver uses, evaluated by the reference's own functor (double) */
double ref_objective_nd(int id, const double *xs, uint64_t d) {
  std::vector<double> x(xs, xs + d);
  namespace tf = nlsolver::test_functions;
  switch (id) {
    case ORC_SPHERE: return tf::Sphere<double>()(x);
    case ORC_ROSENBROCK: return tf::Rosenbrock<double>()(x);
    case ORC_RASTRIGIN: return tf::Rastrigin<double>()(x);
    case ORC_ACKLEY: return tf::Ackley<double>()(x);
    case ORC_BEALE: return tf::Beale<double>()(x);
    case ORC_GOLDSTEIN_PRICE: return tf::Goldstein_Price<double>()(x);
    case ORC_THREE_HUMP_CAMEL: return tf::ThreeHumpCamel<double>()(x);
    case ORC_MCCORMICK: return tf::McCormick<double>()(x);
    case ORC_SCHAFFER_N2: return tf::SchafferN2<double>()(x);
    case ORC_STYBLINSKI_TANG: return tf::StyblinskiTang<double>()(x);
    case ORC_SHEKEL: return tf::Shekel<double>()(x);
    case ORC_BOOTH: return tf::Booth<double>()(x);
    case ORC_BUKIN_N6: return tf::BukinN6<double>()(x);
    case ORC_MATYAS: return tf::Matyas<double>()(x);
    case ORC_LEVI_N13: return tf::LeviN13<double>()(x);
  }
  return 0.0 / 0.0;
}
double ref_objective_2d(int id, double x0, double x1) {
  std::vector<double> x = {x0, x1};
  switch (id) {
    case ORC_SPHERE: return nlsolver::test_functions::Sphere<double>()(x);
    case ORC_ROSENBROCK: return nlsolver::test_functions::Rosenbrock<double>()(x);
    case ORC_RASTRIGIN: return nlsolver::test_functions::Rastrigin<double>()(x);
    case ORC_ACKLEY: return nlsolver::test_functions::Ackley<double>()(x);
  }
  return 0.0 / 0.0;
}
/* first n draws of a default-constructed nlsolver::rng::xorshift<T>, widened to double */
void ref_xorshift_draws(int dtype, uint64_t n, double *out) {
  if (dtype == ORC_F64) { nlsolver::rng::xorshift<double> g; for (uint64_t i = 0; i < n; i++) out[i] = g(); }
  else { nlsolver::rng::xorshift<float> g; for (uint64_t i = 0; i < n; i++) out[i] = g(); }
}
/* how many coordinates / index sequences of the last ref_de_run (this thread) disagreed with the reconstruction */
uint64_t ref_last_inconsistencies(void) { return g_inconsistencies; }
double ref_std_err_f64(const double *x, uint64_t n) { return nlsolver::std_err(std::vector<double>(x, x + n)); }

}  /* extern "C" */
