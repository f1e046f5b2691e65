"""ctypes loader for the CPU checkers under oracle/ — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference legs) may import this module.
The product package (nlsolver_b200/) never does: it has no CPU path at all.

Two libraries share one call surface (oracle/oracle_abi.h):
  * ``oracle()``    -> oracle/_build/liboracle.so : the restatement (popsolve_oracle.cpp), prefix ``oracle_``
  * ``reference()`` -> oracle/_ref/libnls_ref.so  : the UNMODIFIED reference templates behind ref_harness.cpp,
                       prefix ``ref_``; ``None`` when it was never built (reference tree absent and no prebuilt .so)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

F32, F64 = 0, 1
SPHERE, ROSENBROCK, RASTRIGIN, ACKLEY, ROSENBROCK_EX = range(5)
(BEALE, GOLDSTEIN_PRICE, THREE_HUMP_CAMEL, MCCORMICK, SCHAFFER_N2, STYBLINSKI_TANG, SHEKEL, BOOTH, BUKIN_N6, MATYAS,
 LEVI_N13) = range(5, 16)
CUSTOM = 100
DE_BEST, DE_RANDOM = 0, 1
PSO_VANILLA, PSO_ACCELERATED = 0, 1
RNG_TAPE, RNG_XORSHIFT = 0, 1

u64, i32, f64 = C.c_uint64, C.c_int32, C.c_double


class DECfg(C.Structure):
    _fields_ = [("dtype", i32), ("objective", i32), ("strategy", i32), ("minimize", i32),
                ("pop_size", u64), ("dim", u64),
                ("crossover_prob", f64), ("differential_weight", f64), ("eps", f64),
                ("max_iter", u64), ("best_val_no_change", u64),
                ("rng_mode", i32), ("_pad", i32),
                ("seed", u64), ("agent_offset", u64), ("xs_state", u64 * 2)]


class PSOCfg(C.Structure):
    _fields_ = [("dtype", i32), ("objective", i32), ("pso_type", i32), ("minimize", i32),
                ("n_particles", u64), ("dim", u64),
                ("inertia", f64), ("cognitive_coef", f64), ("social_coef", f64), ("eps", f64),
                ("max_iter", u64), ("best_val_no_change", u64),
                ("constrained", i32), ("social_index_j", i32), ("rng_mode", i32), ("_pad", i32),
                ("seed", u64), ("particle_offset", u64), ("n_particles_global", u64), ("xs_state", u64 * 2)]


class Status(C.Structure):
    _fields_ = [("f_value", f64), ("iterations", u64), ("function_calls", u64), ("best_index", u64),
                ("val_no_change", u64), ("draws_consumed", u64), ("best_valid", i32), ("stop_reason", i32),
                ("std_err", f64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class DEOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("x_best", "rows", "scores", "trial_scores", "donors", "dim_idx", "rejects", "accepted", "masks")]


class PSOOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x_best", "positions", "velocities", "pbest_values", "last_values")]


class SANNCfg(C.Structure):
    _fields_ = [("dtype", i32), ("objective", i32), ("minimize", i32), ("rng_mode", i32),
                ("n_chains", u64), ("dim", u64), ("max_iter", u64), ("temperature_iter", u64),
                ("temperature_max", f64), ("seed", u64), ("chain_offset", u64), ("xs_state", u64 * 2),
                ("x0_count", u64), ("max_steps", u64)]


class SANNOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("x_best", "f_best", "p_cur", "n_accepted", "n_improved", "draws", "iterations", "function_calls")]


class NMPSOCfg(C.Structure):
    _fields_ = [("dtype", i32), ("objective", i32), ("minimize", i32), ("rng_mode", i32),
                ("n_solvers", u64), ("dim", u64),
                ("alpha", f64), ("gamma", f64), ("rho", f64), ("sigma", f64), ("inertia", f64), ("cognitive_coef", f64),
                ("social_coef", f64), ("eps", f64), ("max_iter", u64), ("no_change_best_iter", u64),
                ("seed", u64), ("solver_offset", u64), ("xs_state", u64 * 2), ("x0_count", u64)]


class NMPSOOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("x_best", "f_best", "iterations", "function_calls", "draws", "ties")]


def np_dtype(dtype):
    return np.float64 if dtype == F64 else np.float32


def build(ref=True):
    """Compile the checkers (g++; seconds). The reference harness is built only where /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_cache = {}


def _load(path, prefix, with_extras):
    lib = C.CDLL(path)
    de, pso = getattr(lib, prefix + "de_run"), getattr(lib, prefix + "pso_run")
    de.argtypes = [C.POINTER(DECfg), C.c_void_p, C.POINTER(DEOut), C.POINTER(Status)]
    pso.argtypes = [C.POINTER(PSOCfg), C.c_void_p, C.c_void_p, C.POINTER(PSOOut), C.POINTER(Status)]
    de.restype = pso.restype = C.c_int
    sann = getattr(lib, prefix + "sann_run")
    sann.argtypes = [C.POINTER(SANNCfg), C.c_void_p, C.POINTER(SANNOut), C.POINTER(Status)]
    sann.restype = C.c_int
    lib.oracle_objective.argtypes = [C.c_int, C.c_int, C.c_void_p, u64]
    lib.oracle_objective.restype = f64
    lib.oracle_tape_key.argtypes = [u64, u64, u64]
    lib.oracle_tape_key.restype = u64
    lib.oracle_tape_draw.argtypes = [u64, u64]
    lib.oracle_tape_draw.restype = u64
    lib.oracle_unit_f64.argtypes = [u64]
    lib.oracle_unit_f64.restype = f64
    lib.oracle_unit_f32.argtypes = [u64]
    lib.oracle_unit_f32.restype = C.c_float
    lib.oracle_xorshift_default.argtypes = [C.POINTER(u64), C.POINTER(u64), u64]
    lib.oracle_std_err_f64.argtypes = [C.c_void_p, u64]
    lib.oracle_std_err_f64.restype = f64
    if with_extras:
        lib.ref_de_time.argtypes = [C.POINTER(DECfg), C.c_void_p, C.POINTER(f64), C.POINTER(Status)]
        lib.ref_pso_time.argtypes = [C.POINTER(PSOCfg), C.c_void_p, C.POINTER(f64), C.POINTER(Status)]
        lib.ref_sann_time.argtypes = [C.POINTER(SANNCfg), C.c_void_p, C.POINTER(f64), C.POINTER(Status)]
        lib.ref_objective_nd.argtypes = [C.c_int, C.c_void_p, u64]
        lib.ref_objective_nd.restype = f64
        lib.ref_objective_2d.argtypes = [C.c_int, f64, f64]
        lib.ref_objective_2d.restype = f64
        lib.ref_xorshift_draws.argtypes = [C.c_int, u64, C.c_void_p]
        lib.ref_last_inconsistencies.restype = u64
        lib.ref_std_err_f64.argtypes = [C.c_void_p, u64]
        lib.ref_std_err_f64.restype = f64
    return lib


def oracle():
    if "oracle" not in _cache:
        path = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _cache["oracle"] = _load(path, "oracle_", False)
    return _cache["oracle"]


def reference():
    if "ref" not in _cache:
        path = os.path.join(HERE, "_ref", "libnls_ref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/nlsolver.h"):
            build(ref=True)
        _cache["ref"] = _load(path, "ref_", True) if os.path.exists(path) else None
    return _cache["ref"]


def de_cfg(dtype=F64, objective=SPHERE, strategy=DE_RANDOM, minimize=True, pop_size=50, dim=2, crossover_prob=0.9,
           differential_weight=0.8, eps=10e-4, max_iter=1000, best_val_no_change=50, rng_mode=RNG_TAPE, seed=0,
           agent_offset=0, xs_state=(0, 0)):
    return DECfg(dtype, objective, strategy, int(minimize), pop_size, dim, crossover_prob, differential_weight, eps,
                 max_iter, best_val_no_change, rng_mode, 0, seed, agent_offset, (u64 * 2)(*xs_state))


def pso_cfg(dtype=F64, objective=SPHERE, pso_type=PSO_VANILLA, minimize=True, n_particles=10, dim=2, inertia=0.8,
            cognitive_coef=1.8, social_coef=1.8, eps=10e-4, max_iter=5000, best_val_no_change=50, constrained=False,
            social_index_j=False, rng_mode=RNG_TAPE, seed=0, particle_offset=0, n_particles_global=0, xs_state=(0, 0)):
    return PSOCfg(dtype, objective, pso_type, int(minimize), n_particles, dim, inertia, cognitive_coef, social_coef,
                  eps, max_iter, best_val_no_change, int(constrained), int(social_index_j), rng_mode, 0, seed,
                  particle_offset, n_particles_global or n_particles, (u64 * 2)(*xs_state))


def de_run(lib, cfg, x0, masks=False, prefix=None):
    """Run DE to its stop rule; returns (status dict, arrays dict)."""
    prefix = prefix or ("ref_" if hasattr(lib, "ref_de_run") else "oracle_")
    dt = np_dtype(cfg.dtype)
    P, d = cfg.pop_size, cfg.dim
    x0 = np.ascontiguousarray(x0, dtype=dt)
    a = {"x_best": np.zeros(d, dt), "rows": np.zeros((P, d), dt), "scores": np.zeros(P, dt),
         "trial_scores": np.zeros(P, dt), "donors": np.zeros((P, 3), np.uint32), "dim_idx": np.zeros(P, np.uint32),
         "rejects": np.zeros(P, np.uint32), "accepted": np.zeros(P, np.uint8)}
    if masks:
        a["masks"] = np.zeros((P, d), np.uint8)
    out = DEOut(**{k: v.ctypes.data for k, v in a.items()})
    st = Status()
    rc = getattr(lib, prefix + "de_run")(C.byref(cfg), x0.ctypes.data, C.byref(out), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"{prefix}de_run failed: {rc}")
    return st.as_dict(), a


def pso_run(lib, cfg, lower, upper, prefix=None):
    prefix = prefix or ("ref_" if hasattr(lib, "ref_pso_run") else "oracle_")
    dt = np_dtype(cfg.dtype)
    P, d = cfg.n_particles, cfg.dim
    lower = np.ascontiguousarray(lower, dtype=dt)
    upper = np.ascontiguousarray(upper, dtype=dt)
    a = {"x_best": np.zeros(d, dt), "positions": np.zeros((P, d), dt), "velocities": np.zeros((P, d), dt),
         "pbest_values": np.zeros(P, dt), "last_values": np.zeros(P, dt)}
    out = PSOOut(**{k: v.ctypes.data for k, v in a.items()})
    st = Status()
    rc = getattr(lib, prefix + "pso_run")(C.byref(cfg), lower.ctypes.data, upper.ctypes.data, C.byref(out),
                                          C.byref(st))
    if rc != 0:
        raise RuntimeError(f"{prefix}pso_run failed: {rc}")
    return st.as_dict(), a


def sann_cfg(dtype=F64, objective=SPHERE, minimize=True, n_chains=1, dim=2, max_iter=5000, temperature_iter=10,
             temperature_max=10.0, rng_mode=RNG_TAPE, seed=0, chain_offset=0, xs_state=(0, 0), x0_count=1,
             max_steps=0):
    return SANNCfg(dtype, objective, int(minimize), rng_mode, n_chains, dim, max_iter, temperature_iter,
                   temperature_max, seed, chain_offset, (u64 * 2)(*xs_state), x0_count, max_steps)


def sann_run(lib, cfg, x0, prefix=None):
    """Run a batch of SANN chains; x0 is [dim] (shared start) or [n_chains, dim]. Returns (status, arrays)."""
    prefix = prefix or ("ref_" if hasattr(lib, "ref_sann_run") else "oracle_")
    dt = np_dtype(cfg.dtype)
    n, d = cfg.n_chains, cfg.dim
    x0 = np.ascontiguousarray(x0, dtype=dt)
    cfg.x0_count = 1 if x0.ndim == 1 else n
    assert x0.size == cfg.x0_count * d
    a = {"x_best": np.zeros((n, d), dt), "f_best": np.zeros(n, dt), "p_cur": np.zeros((n, d), dt),
         "n_accepted": np.zeros(n, np.uint32), "n_improved": np.zeros(n, np.uint32), "draws": np.zeros(n, np.uint64),
         "iterations": np.zeros(n, np.uint64), "function_calls": np.zeros(n, np.uint64)}
    out = SANNOut(**{k: v.ctypes.data for k, v in a.items()})
    st = Status()
    rc = getattr(lib, prefix + "sann_run")(C.byref(cfg), x0.ctypes.data, C.byref(out), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"{prefix}sann_run failed: {rc}")
    return st.as_dict(), a


def nmpso_cfg(dtype=F64, objective=SPHERE, minimize=True, n_solvers=1, dim=2, alpha=1.0, gamma=2.0, rho=0.5, sigma=0.5,
              inertia=0.8, cognitive_coef=1.8, social_coef=1.8, eps=1e-6, max_iter=1000, no_change_best_iter=20,
              rng_mode=RNG_TAPE, seed=0, solver_offset=0, xs_state=(0, 0), x0_count=1):
    return NMPSOCfg(dtype, objective, int(minimize), rng_mode, n_solvers, dim, alpha, gamma, rho, sigma, inertia,
                    cognitive_coef, social_coef, eps, max_iter, no_change_best_iter, seed, solver_offset,
                    (u64 * 2)(*xs_state), x0_count)


def nmpso_run(lib, cfg, x0, prefix=None):
    """A batch of NelderMeadPSO solvers; x0 is [dim] or [n_solvers, dim].  Returns (status, arrays); status is None when
    the reference harness refuses the shape (the reference would corrupt the heap, oracle_abi.h)."""
    prefix = prefix or ("ref_" if hasattr(lib, "ref_de_run") else "oracle_")
    dt = np_dtype(cfg.dtype)
    n, d = cfg.n_solvers, cfg.dim
    x0 = np.ascontiguousarray(x0, dtype=dt)
    cfg.x0_count = 1 if x0.ndim == 1 else n
    a = {"x_best": np.zeros((n, d), dt), "f_best": np.zeros(n, dt), "iterations": np.zeros(n, np.uint64),
         "function_calls": np.zeros(n, np.uint64), "draws": np.zeros(n, np.uint64), "ties": np.zeros(n, np.uint8)}
    out = NMPSOOut(**{k: v.ctypes.data for k, v in a.items()})
    st = Status()
    fn = getattr(lib, prefix + "nmpso_run")
    fn.argtypes = [C.POINTER(NMPSOCfg), C.c_void_p, C.POINTER(NMPSOOut), C.POINTER(Status)]
    fn.restype = C.c_int
    rc = fn(C.byref(cfg), x0.ctypes.data, C.byref(out), C.byref(st))
    if rc != 0:
        return None, a
    return st.as_dict(), a


TERM_FN = C.CFUNCTYPE(f64, f64, f64, u64, u64)
FINISH_FN = C.CFUNCTYPE(f64, f64, u64)
_custom_keepalive = []


def set_custom_objective(term, finish, pairwise=False, lane0_seed=0.0):
    """Install Python callables as objective CUSTOM of the restatement (small cases only: one call per coordinate)."""
    lib = oracle()
    t, f = TERM_FN(term), FINISH_FN(finish)
    _custom_keepalive[:] = [t, f]
    lib.oracle_set_custom_objective.argtypes = [C.c_int, f64, TERM_FN, FINISH_FN]
    lib.oracle_set_custom_objective(int(pairwise), lane0_seed, t, f)


FULL_FN = C.CFUNCTYPE(f64, C.POINTER(f64), u64)


def set_custom_full(fn):
    """Install a Python callable f(list of coordinates) -> value as objective CUSTOM (closed forms, small d)."""
    lib = oracle()
    cb = FULL_FN(lambda p, d: float(fn([p[k] for k in range(d)])))
    _custom_keepalive[:] = [cb]
    lib.oracle_set_custom_full.argtypes = [FULL_FN]
    lib.oracle_set_custom_full(cb)


def objective(dtype, obj_id, x):
    x = np.ascontiguousarray(x, dtype=np_dtype(dtype))
    return oracle().oracle_objective(dtype, obj_id, x.ctypes.data, x.size)


# ------------------------------------------------------------------ stepwise handles (fp64, tape) -----------------
def _step_api(lib):
    if getattr(lib, "_step_api_ready", False):
        return lib
    P = C.c_void_p
    lib.oracle_de_open.argtypes = [C.POINTER(DECfg), P]
    lib.oracle_de_open.restype = P
    lib.oracle_de_advance.argtypes = [P, u64]
    lib.oracle_de_report.argtypes = [P, C.POINTER(DEOut), C.POINTER(Status)]
    lib.oracle_de_export_top.argtypes = [P, u64, P, P]
    lib.oracle_de_import_migrants.argtypes = [P, u64, P, P]
    lib.oracle_de_close.argtypes = [P]
    lib.oracle_pso_open.argtypes = [C.POINTER(PSOCfg), P, P]
    lib.oracle_pso_open.restype = P
    lib.oracle_pso_evaluate.argtypes = [P, C.POINTER(f64), C.POINTER(u64)]
    lib.oracle_pso_evaluate.restype = C.c_int
    lib.oracle_pso_row.argtypes = [P, u64]
    lib.oracle_pso_row.restype = C.POINTER(f64)
    lib.oracle_pso_pbest.argtypes = [P]
    lib.oracle_pso_pbest.restype = C.POINTER(f64)
    lib.oracle_pso_adopt.argtypes = [P, C.c_int, f64, u64, P, u64, f64, C.c_int]
    lib.oracle_pso_adopt.restype = C.c_int
    lib.oracle_pso_move.argtypes = [P]
    lib.oracle_pso_report.argtypes = [P, C.POINTER(PSOOut), C.POINTER(Status)]
    lib.oracle_pso_close.argtypes = [P]
    lib._step_api_ready = True
    return lib


class DEStepper:
    """Generation-by-generation DE on the oracle (island tests): advance / report / export_top / import_migrants."""

    def __init__(self, cfg, x0):
        self.lib, self.cfg = _step_api(oracle()), cfg
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.h = self.lib.oracle_de_open(C.byref(cfg), x0.ctypes.data)
        assert self.h, "oracle_de_open: fp64 tape configurations only"

    def advance(self, n=1):
        self.lib.oracle_de_advance(self.h, n)

    def report(self):
        P, d = self.cfg.pop_size, self.cfg.dim
        a = {"x_best": np.zeros(d), "rows": np.zeros((P, d)), "scores": np.zeros(P)}
        out = DEOut(**{k: v.ctypes.data for k, v in a.items()})
        st = Status()
        self.lib.oracle_de_report(self.h, C.byref(out), C.byref(st))
        return st.as_dict(), a

    def export_top(self, k):
        rows, scores = np.zeros((k, self.cfg.dim)), np.zeros(k)
        self.lib.oracle_de_export_top(self.h, k, rows.ctypes.data, scores.ctypes.data)
        return rows, scores

    def import_migrants(self, rows, scores):
        rows, scores = np.ascontiguousarray(rows, np.float64), np.ascontiguousarray(scores, np.float64)
        self.lib.oracle_de_import_migrants(self.h, scores.size, rows.ctypes.data, scores.ctypes.data)

    def close(self):
        if self.h:
            self.lib.oracle_de_close(self.h)
            self.h = None


class PSOStepper:
    """Phase-by-phase PSO shard on the oracle: evaluate -> (exchange) -> adopt -> move."""

    def __init__(self, cfg, lower, upper):
        self.lib, self.cfg = _step_api(oracle()), cfg
        lower = np.ascontiguousarray(lower, dtype=np.float64)
        upper = np.ascontiguousarray(upper, dtype=np.float64)
        self.h = self.lib.oracle_pso_open(C.byref(cfg), lower.ctypes.data, upper.ctypes.data)
        assert self.h, "oracle_pso_open: fp64 tape configurations only"

    def evaluate(self):
        """-> (has_candidate, value, local_index, row copy, pbest copy)."""
        v, i = f64(), u64()
        any_ = self.lib.oracle_pso_evaluate(self.h, C.byref(v), C.byref(i))
        d, P = self.cfg.dim, self.cfg.n_particles
        row = np.ctypeslib.as_array(self.lib.oracle_pso_row(self.h, i.value), shape=(d,)).copy()
        pbest = np.ctypeslib.as_array(self.lib.oracle_pso_pbest(self.h), shape=(P,)).copy()
        return bool(any_), v.value, i.value, row, pbest

    def adopt(self, have, value, global_index, row, n_global, std_err_all, initial):
        row = np.ascontiguousarray(row, np.float64)
        return bool(self.lib.oracle_pso_adopt(self.h, int(have), value, global_index, row.ctypes.data, n_global,
                                              std_err_all, int(initial)))

    def move(self):
        self.lib.oracle_pso_move(self.h)

    def report(self):
        P, d = self.cfg.n_particles, self.cfg.dim
        a = {"x_best": np.zeros(d), "positions": np.zeros((P, d)), "pbest_values": np.zeros(P),
             "last_values": np.zeros(P)}
        out = PSOOut(**{k: v.ctypes.data for k, v in a.items()})
        st = Status()
        self.lib.oracle_pso_report(self.h, C.byref(out), C.byref(st))
        return st.as_dict(), a

    def close(self):
        if self.h:
            self.lib.oracle_pso_close(self.h)
            self.h = None
