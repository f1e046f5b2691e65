"""ctypes binding of libnls_b200.so (C ABI: include/nls_b200.h).

The shared library is the product; this module only marshals arguments.  There is no CPU path: if the library is
missing or no CUDA device is present the calls raise."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NLS_B200_LIB selects an alternative build of the same library (kernel tuning experiments, tools/build_variants.sh)
LIB_PATH = os.environ.get("NLS_B200_LIB") or os.path.join(HERE, "libnls_b200.so")

F32, F64 = 0, 1
SPHERE, ROSENBROCK, RASTRIGIN, ACKLEY, ROSENBROCK_EX = range(5)
(BEALE, GOLDSTEIN_PRICE, THREE_HUMP_CAMEL, MCCORMICK, SCHAFFER_N2, STYBLINSKI_TANG, SHEKEL, BOOTH, BUKIN_N6, MATYAS,
 LEVI_N13) = range(5, 16)
DE_BEST, DE_RANDOM = 0, 1
PSO_VANILLA, PSO_ACCELERATED = 0, 1
FLAG_RECORD_MASKS, FLAG_SOCIAL_INDEX_J = 1, 2
XCHG_HANDLE_BYTES = 64

u64, i32, u32, f64 = C.c_uint64, C.c_int32, C.c_uint32, C.c_double


class DECfg(C.Structure):
    _fields_ = [("dtype", i32), ("objective", i32), ("strategy", i32), ("minimize", i32),
                ("pop_size", u64), ("dim", u64),
                ("crossover_prob", f64), ("differential_weight", f64), ("eps", f64),
                ("max_iter", u64), ("best_val_no_change", u64),
                ("seed", u64), ("agent_offset", u64), ("flags", u32), ("_reserved", u32)]


class PSOCfg(C.Structure):
    _fields_ = [("dtype", i32), ("objective", i32), ("pso_type", i32), ("minimize", i32),
                ("n_particles", u64), ("dim", u64),
                ("inertia", f64), ("cognitive_coef", f64), ("social_coef", f64), ("eps", f64),
                ("max_iter", u64), ("best_val_no_change", u64),
                ("constrained", i32), ("flags", u32),
                ("seed", u64), ("particle_offset", u64), ("n_particles_global", u64)]


class SANNCfg(C.Structure):
    _fields_ = [("dtype", i32), ("objective", i32), ("minimize", i32), ("flags", u32),
                ("n_chains", u64), ("dim", u64), ("max_iter", u64), ("temperature_iter", u64),
                ("temperature_max", f64), ("seed", u64), ("chain_offset", u64)]


class NMPSOCfg(C.Structure):
    _fields_ = [("dtype", i32), ("objective", i32), ("minimize", i32), ("flags", u32),
                ("n_solvers", u64), ("dim", u64),
                ("alpha", f64), ("gamma", f64), ("rho", f64), ("sigma", f64), ("inertia", f64), ("cognitive_coef", f64),
                ("social_coef", f64), ("eps", f64), ("max_iter", u64), ("no_change_best_iter", u64),
                ("seed", u64), ("solver_offset", u64)]


class Status(C.Structure):
    _fields_ = [("f_value", f64), ("iterations", u64), ("function_calls", u64), ("best_index", u64),
                ("val_no_change", u64), ("stopped", i32), ("stop_reason", i32), ("best_valid", i32),
                ("_reserved", i32), ("std_err", f64), ("repair_reruns", u64), ("repair_rounds", u64),
                ("accepted_total", u64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


class NlsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libnls_b200 error {code}: {message}")
        self.code = code


# every symbol include/nls_b200.h declares: name -> (restype, argtypes)
P = C.c_void_p
SYMBOLS = {
    "nls_last_error": (C.c_char_p, []),
    "nls_version": (C.c_int, []),
    "nls_ctx_create": (C.c_int, [C.c_int, P, C.POINTER(P)]),
    "nls_ctx_destroy": (C.c_int, [P]),
    "nls_ctx_trim": (C.c_int, [P]),
    "nls_ctx_set_pool_limit": (C.c_int, [P, u64]),
    "nls_ctx_pool_bytes": (u64, [P]),
    "nls_debug_guard_violations": (C.c_ulonglong, []),
    "nls_ctx_device": (C.c_int, [P]),
    "nls_ctx_sm_count": (C.c_int, [P]),
    "nls_de_solve": (C.c_int, [P, C.POINTER(DECfg), P, P, C.POINTER(Status)]),
    "nls_pso_solve": (C.c_int, [P, C.POINTER(PSOCfg), P, P, P, C.POINTER(Status)]),
    "nls_de_create": (C.c_int, [P, C.POINTER(DECfg), P, C.POINTER(P)]),
    "nls_de_step": (C.c_int, [P, u64]),
    "nls_de_sync": (C.c_int, [P, C.POINTER(Status)]),
    "nls_de_read_best": (C.c_int, [P, P]),
    "nls_de_read_population": (C.c_int, [P, P]),
    "nls_de_read_scores": (C.c_int, [P, P]),
    "nls_de_read_rows": (C.c_int, [P, u64, u64, P]),
    "nls_de_read_decisions": (C.c_int, [P, P, P, P, P, P, P]),
    "nls_de_destroy": (C.c_int, [P]),
    "nls_de_enable_kernel_timing": (C.c_int, [P, C.c_int]),
    "nls_de_kernel_times": (C.c_int, [P, C.POINTER(f64), C.POINTER(u64)]),
    "nls_record_bytes": (u64, [i32, u64]),
    "nls_de_export_best": (C.c_int, [P, P]),
    "nls_de_export_top": (C.c_int, [P, u64, P, P]),
    "nls_de_import_migrants": (C.c_int, [P, u64, P, P]),
    "nls_pso_create": (C.c_int, [P, C.POINTER(PSOCfg), P, P, C.POINTER(P)]),
    "nls_pso_step": (C.c_int, [P, u64]),
    "nls_pso_sync": (C.c_int, [P, C.POINTER(Status)]),
    "nls_pso_read_best": (C.c_int, [P, P]),
    "nls_pso_read_positions": (C.c_int, [P, P]),
    "nls_pso_read_velocities": (C.c_int, [P, P]),
    "nls_pso_read_pbest_values": (C.c_int, [P, P]),
    "nls_pso_read_last_values": (C.c_int, [P, P]),
    "nls_pso_destroy": (C.c_int, [P]),
    "nls_pso_step_local": (C.c_int, [P, P]),
    "nls_pso_export_candidate": (C.c_int, [P, P]),
    "nls_pso_apply_candidates": (C.c_int, [P, P, u64]),
    "nls_load_objective": (C.c_int, [C.c_char_p, C.POINTER(i32)]),
    "nls_xchg_create": (C.c_int, [P, u64, C.c_int, C.c_int, C.POINTER(P)]),
    "nls_xchg_get_handle": (C.c_int, [P, P]),
    "nls_xchg_open_peers": (C.c_int, [P, P]),
    "nls_xchg_destroy": (C.c_int, [P]),
    "nls_pso_attach_exchange": (C.c_int, [P, P]),
    "nls_pso_step_fused": (C.c_int, [P, u64]),
    "nls_de_attach_exchange": (C.c_int, [P, P]),
    "nls_de_read_exchange": (C.c_int, [P, P]),
    "nls_nmpso_solve": (C.c_int, [P, C.POINTER(NMPSOCfg), P, u64, P, P, P, P, C.POINTER(Status)]),
    "nls_group_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(P)]),
    "nls_group_destroy": (C.c_int, [P]),
    "nls_group_size": (C.c_int, [P]),
    "nls_pso_sharded_create": (C.c_int, [P, C.POINTER(PSOCfg), P, P, C.POINTER(P)]),
    "nls_pso_sharded_step": (C.c_int, [P, u64]),
    "nls_pso_sharded_sync": (C.c_int, [P, C.POINTER(Status)]),
    "nls_pso_sharded_read_best": (C.c_int, [P, P]),
    "nls_pso_sharded_shard": (C.c_int, [P, C.c_int, C.POINTER(P)]),
    "nls_pso_sharded_destroy": (C.c_int, [P]),
    "nls_pso_solve_sharded": (C.c_int, [P, C.POINTER(PSOCfg), P, P, P, C.POINTER(Status)]),
    "nls_de_islands_create": (C.c_int, [P, C.POINTER(DECfg), P, u64, u64, C.POINTER(P)]),
    "nls_de_islands_step": (C.c_int, [P, u64]),
    "nls_de_islands_sync": (C.c_int, [P, C.POINTER(Status)]),
    "nls_de_islands_read_best": (C.c_int, [P, P]),
    "nls_de_islands_island": (C.c_int, [P, C.c_int, C.POINTER(P)]),
    "nls_de_islands_destroy": (C.c_int, [P]),
    "nls_de_solve_islands": (C.c_int, [P, C.POINTER(DECfg), P, u64, u64, P, C.POINTER(Status)]),
    "nls_sann_create": (C.c_int, [P, C.POINTER(SANNCfg), P, u64, C.POINTER(P)]),
    "nls_sann_step": (C.c_int, [P, u64]),
    "nls_sann_sync": (C.c_int, [P, C.POINTER(Status)]),
    "nls_sann_read_best": (C.c_int, [P, P]),
    "nls_sann_read_chains": (C.c_int, [P, P, P, P, P, P]),
    "nls_sann_destroy": (C.c_int, [P]),
    "nls_sann_solve": (C.c_int, [P, C.POINTER(SANNCfg), P, u64, P, C.POINTER(Status)]),
}

_lib = None


def lib():
    """Load libnls_b200.so (built in-tree by `__graft_entry__.build()` / `make -C nlsolver_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C nlsolver_b200/csrc` "
                              "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                if os.environ.get("NLS_B200_LIB"):   # an older experimental build (A/B timing) may lack newer symbols
                    continue
                raise
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise NlsError(rc, lib().nls_last_error().decode())
