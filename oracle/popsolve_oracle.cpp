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
#include <cmath>
#include <cstdint>
#include <cstring>
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

template <class T>
T objective(int id, const T *x, size_t d) {
  const T two_pi = static_cast<T>(2 * M_PI);
  switch (id) {
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

template <class T, class Src>
void de_run(const orc_de_cfg &c, const T *x0, Src &src, const orc_de_out *out, orc_status *st) {
  const size_t P = c.pop_size, d = c.dim;
  const T CR = static_cast<T>(c.crossover_prob), F = static_cast<T>(c.differential_weight);
  const T eps = static_cast<T>(c.eps);
  const T fm = c.minimize ? static_cast<T>(1.0) : static_cast<T>(-1.0);   /* nlsolver.h:2418 */
  std::vector<T> A(P * d), scores(P), trial(d);

  /* init_agents / generate_sequence, nlsolver.h:2302-2323: agent[j] = (g() - 0.5) * x0[j], agent-major */
  for (size_t i = 0; i < P; i++) {
    src.begin(0, i);
    for (size_t j = 0; j < d; j++) A[i * d + j] = static_cast<T>((unit<T>(src.next()) - 0.5) * x0[j]);
  }
  for (size_t i = 0; i < P; i++) scores[i] = fm * objective<T>(c.objective, &A[i * d], d);   /* :2423-2425 */

  u64 fcalls = P, iter = 0, best_id = 0, vnc = 0;
  int stop_reason = 0;
  T se = 0;
  std::vector<uint32_t> donors(P * 3), dimv(P), rej(P);
  std::vector<uint8_t> acc(P), masks(out && out->masks ? P * d : 0);
  std::vector<T> tscores(P);

  while (true) {
    bool not_updated = true;                      /* :2430-2439 */
    for (size_t i = 0; i < P; i++)
      if (scores[i] < scores[best_id]) { best_id = i; not_updated = false; }
    vnc = not_updated * (vnc + 1);
    if (iter >= c.max_iter) stop_reason = 1;      /* :2441-2447, short-circuit order preserved */
    else if (vnc >= c.best_val_no_change) stop_reason = 2;
    else { se = std_err(scores); if (se < eps) stop_reason = 3; }
    if (stop_reason) break;

    for (size_t i = 0; i < P; i++) {              /* :2449-2472, sequential and IN PLACE */
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
  if (out->masks) std::memcpy(out->masks, masks.data(), P * d);
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

template <class T, class Src>
void pso_run(const orc_pso_cfg &c, const T *lower, const T *upper, Src &src, const orc_pso_out *out,
             orc_status *st) {
  const size_t P = c.n_particles, d = c.dim;
  const bool vanilla = c.pso_type == ORC_PSO_VANILLA;
  const T init_inertia = static_cast<T>(c.inertia);
  T inertia = init_inertia;
  const T cog = static_cast<T>(c.cognitive_coef), soc = static_cast<T>(c.social_coef);
  const T eps = static_cast<T>(c.eps);
  const T fm = c.minimize ? static_cast<T>(1.0) : static_cast<T>(-1.0);
  std::vector<T> X(P * d), V(vanilla ? P * d : 0), pbest(P, static_cast<T>(10000)), last(P), sbest;
  T sbest_val = static_cast<T>(100000.0);          /* init_solver_state, nlsolver.h:2626-2657 */
  u64 f_evals = 0, vnc = 0, iter = 0, sbest_idx = 0;
  for (size_t i = 0; i < P; i++) {
    src.begin(0, i);
    for (size_t j = 0; j < d; j++) {
      const T temp = std::abs(upper[j] - lower[j]);
      X[i * d + j] = lower[j] + ((upper[j] - lower[j]) * unit<T>(src.next()));
      if (vanilla) V[i * d + j] = -temp + (unit<T>(src.next()) * temp);
    }
  }
  auto update_best = [&]() {                        /* update_best_positions, :2716-2741 */
    size_t best_index = 0;
    bool update_happened = false;
    for (size_t i = 0; i < P; i++) {
      const T temp = fm * objective<T>(c.objective, &X[i * d], d);
      last[i] = temp;
      if (temp < sbest_val) { sbest_val = temp; best_index = i; update_happened = true; }
      if (temp < pbest[i]) pbest[i] = temp;
    }
    f_evals += P;
    if (update_happened) { sbest.assign(&X[best_index * d], &X[best_index * d] + d); sbest_idx = best_index; }
    vnc = (best_index == 0) * (vnc + 1);
  };
  update_best();                                    /* solve, :2592-2624 */
  int stop_reason = 0;
  T se = 0;
  while (true) {
    if (iter >= c.max_iter) stop_reason = 1;
    else if (vnc >= c.best_val_no_change) stop_reason = 2;
    else { se = std_err(pbest); if (se < eps) stop_reason = 3; }
    if (stop_reason) break;
    if (vanilla) {                                  /* update_velocities, :2658-2677 (both quirks kept) */
      for (size_t i = 0; i < P; i++) {
        src.begin(iter + 1, i);
        for (size_t j = 0; j < d; j++) {
          const T r_p = unit<T>(src.next()), r_g = unit<T>(src.next());
          const T x = X[i * d + j];
          const T sb = sbest.empty() ? static_cast<T>(0) : sbest[c.social_index_j ? j : i];
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
    update_best();
    iter++;
  }
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
/* PSO::minimize(x) without bounds derives lower = -|x|, upper = |x| (nlsolver.h:2553-2563); callers do that. */

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
