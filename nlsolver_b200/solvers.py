"""Host-side mirror of the reference's solver interface for the DE / PSO population loop, over the C ABI.

Names, argument order, defaults and behaviour follow nlsolver::DE (nlsolver.h:2379-2410), nlsolver::PSO
(nlsolver.h:2498-2589), nlsolver::SANN (nlsolver.h:2744-2815) and nlsolver::solver_status (nlsolver.h:2054-2097): `minimize(x)` / `maximize(x)` overwrite
`x` with the best point and return a `SolverStatus`.  The objective is one of the device functors (`Sphere`,
`Rosenbrock`, `Rastrigin`, `Ackley`, `RosenbrockExample`); the random generator is any callable returning floats in
[0, 1] — two draws are taken from it per solve to seed the device draw tape, so it advances deterministically.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib as L

Sphere, Rosenbrock, Rastrigin, Ackley, RosenbrockExample = (L.SPHERE, L.ROSENBROCK, L.RASTRIGIN, L.ACKLEY,
                                                             L.ROSENBROCK_EX)
# the other problems of the reference's test driver (test_functions.h:94-318)
(Beale, Goldstein_Price, ThreeHumpCamel, McCormick, SchafferN2, StyblinskiTang, Shekel, Booth, BukinN6, Matyas,
 LeviN13) = range(5, 16)


class RecombinationStrategy:  # nlsolver.h:2377
    best, random = L.DE_BEST, L.DE_RANDOM


class PSOType:  # nlsolver.h:2496
    Vanilla, Accelerated = L.PSO_VANILLA, L.PSO_ACCELERATED


def np_dtype(dtype):
    return np.float64 if dtype == L.F64 else np.float32


def nls_dtype(scalar_t):
    return L.F64 if np.dtype(scalar_t) == np.float64 else L.F32


def seed_from_generator(generator):
    """Two draws -> 64-bit tape seed (the same rule as include/nlsolver_b200.hpp)."""
    hi = min(int(float(generator()) * 4294967296.0), 0xFFFFFFFF)
    lo = min(int(float(generator()) * 4294967296.0), 0xFFFFFFFF)
    return (hi << 32) | lo


class SolverStatus:
    """solver_status<scalar_t> (nlsolver.h:2054-2097)."""

    def __init__(self, f_value, iteration, function_calls_used, gradient_evals_used=0, hessian_evals_used=0):
        self.f_value = f_value
        self.iteration = iteration
        self.function_calls_used = function_calls_used
        self.gradient_evals_used = gradient_evals_used
        self.hessian_evals_used = hessian_evals_used

    def print(self):
        print(f"Function calls used: {self.function_calls_used}")
        print(f"Algorithm iterations used: {self.iteration}")
        if self.gradient_evals_used > 0:
            print(f"Gradient evaluations used: {self.gradient_evals_used}")
        if self.hessian_evals_used > 0:
            print(f"Hessian evaluations used: {self.hessian_evals_used}")
        print(f"With final function value of {self.f_value:g}")

    def get_summary(self):
        return (self.function_calls_used, self.iteration, self.f_value, self.gradient_evals_used,
                self.hessian_evals_used)

    def add(self, other):
        calls, it, f, g, h = other.get_summary()
        self.function_calls_used += calls
        self.iteration += it
        self.f_value = f
        self.gradient_evals_used += g
        self.hessian_evals_used += h


class Context:
    """One per (process, GPU). `stream` is a raw cudaStream_t (int) or None for a library-owned stream."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        self._children = weakref.WeakSet()   # live solver handles: they must be destroyed before the context
        L.check(L.lib().nls_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))

    def _adopt(self, child):
        self._children.add(child)

    @property
    def handle(self):
        return self._h

    @property
    def sm_count(self):
        return L.lib().nls_ctx_sm_count(self._h)

    def trim(self):
        """Return the cached device buffers of finished solves to the driver."""
        L.check(L.lib().nls_ctx_trim(self._h))

    def set_pool_limit(self, n_bytes):
        """Bound the cache of released device buffers (0: the default, the largest single handle released so far)."""
        L.check(L.lib().nls_ctx_set_pool_limit(self._h, n_bytes))

    @property
    def pool_bytes(self):
        return L.lib().nls_ctx_pool_bytes(self._h)

    def close(self):
        if self._h:
            for child in list(self._children):
                child.close()
            L.lib().nls_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DEPopulation:
    """Stepwise handle (nls_de_*): the DE loop cut at generation boundaries, population resident in HBM."""

    def __init__(self, ctx, cfg, x0):
        self.ctx, self.cfg = ctx, cfg
        self.dt = np_dtype(cfg.dtype)
        x0 = np.ascontiguousarray(x0, dtype=self.dt)
        assert x0.size == cfg.dim
        self._h = C.c_void_p()
        L.check(L.lib().nls_de_create(ctx.handle, C.byref(cfg), x0.ctypes.data, C.byref(self._h)))
        ctx._adopt(self)

    def step(self, n=1):
        L.check(L.lib().nls_de_step(self._h, n))

    def sync(self):
        st = L.Status()
        L.check(L.lib().nls_de_sync(self._h, C.byref(st)))
        return st.as_dict()

    def best(self):
        x = np.zeros(self.cfg.dim, self.dt)
        L.check(L.lib().nls_de_read_best(self._h, x.ctypes.data))
        return x

    def population(self):
        rows = np.zeros((self.cfg.pop_size, self.cfg.dim), self.dt)
        L.check(L.lib().nls_de_read_population(self._h, rows.ctypes.data))
        return rows

    def rows(self, first, count):
        out = np.zeros((count, self.cfg.dim), self.dt)
        L.check(L.lib().nls_de_read_rows(self._h, first, count, out.ctypes.data))
        return out

    def scores(self):
        s = np.zeros(self.cfg.pop_size, self.dt)
        L.check(L.lib().nls_de_read_scores(self._h, s.ctypes.data))
        return s

    def decisions(self, masks=False):
        P, d = self.cfg.pop_size, self.cfg.dim
        a = {"donors": np.zeros((P, 3), np.uint32), "dim_idx": np.zeros(P, np.uint32),
             "rejects": np.zeros(P, np.uint32), "accepted": np.zeros(P, np.uint8),
             "trial_scores": np.zeros(P, self.dt)}
        m = np.zeros((P, d), np.uint8) if masks else None
        L.check(L.lib().nls_de_read_decisions(self._h, a["donors"].ctypes.data, a["dim_idx"].ctypes.data,
                                              a["rejects"].ctypes.data, a["accepted"].ctypes.data,
                                              a["trial_scores"].ctypes.data, m.ctypes.data if masks else None))
        if masks:
            a["masks"] = m
        return a

    def enable_kernel_timing(self, enable=True):
        L.check(L.lib().nls_de_enable_kernel_timing(self._h, int(enable)))

    def kernel_times(self):
        """(ms per kernel [K2 generation pass, K2r repair, K3 commit+reduce], generations) since the last call."""
        ms, n = (L.f64 * 3)(), L.u64()
        L.check(L.lib().nls_de_kernel_times(self._h, ms, C.byref(n)))
        return list(ms), n.value

    def export_best(self, record_ptr):
        L.check(L.lib().nls_de_export_best(self._h, C.c_void_p(record_ptr)))

    def export_top(self, k, rows_ptr, scores_ptr):
        L.check(L.lib().nls_de_export_top(self._h, k, C.c_void_p(rows_ptr), C.c_void_p(scores_ptr)))

    def import_migrants(self, k, rows_ptr, scores_ptr):
        L.check(L.lib().nls_de_import_migrants(self._h, k, C.c_void_p(rows_ptr), C.c_void_p(scores_ptr)))

    def attach_exchange(self, window):
        """From here on the commit kernel of every generation stores the island's record into every peer's window."""
        L.check(L.lib().nls_de_attach_exchange(self._h, window.handle))
        self._window = window

    def read_exchange(self, world):
        """Newest record of every island out of this island's own window: uint8 [world, record_bytes]."""
        rb = L.lib().nls_record_bytes(self.cfg.dtype, self.cfg.dim)
        out = np.zeros((world, rb), np.uint8)
        L.check(L.lib().nls_de_read_exchange(self._h, out.ctypes.data))
        return out

    def close(self):
        if self._h:
            L.lib().nls_de_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PSOSwarm:
    """Stepwise handle (nls_pso_*)."""

    def __init__(self, ctx, cfg, lower, upper):
        self.ctx, self.cfg = ctx, cfg
        self.dt = np_dtype(cfg.dtype)
        lower = np.ascontiguousarray(lower, dtype=self.dt)
        upper = np.ascontiguousarray(upper, dtype=self.dt)
        assert lower.size == cfg.dim and upper.size == cfg.dim
        self._h = C.c_void_p()
        L.check(L.lib().nls_pso_create(ctx.handle, C.byref(cfg), lower.ctypes.data, upper.ctypes.data,
                                       C.byref(self._h)))
        ctx._adopt(self)

    def step(self, n=1):
        L.check(L.lib().nls_pso_step(self._h, n))

    def step_local(self, record_ptr=None):
        L.check(L.lib().nls_pso_step_local(self._h, C.c_void_p(record_ptr) if record_ptr else None))

    def export_candidate(self, record_ptr):
        L.check(L.lib().nls_pso_export_candidate(self._h, C.c_void_p(record_ptr)))

    def apply_candidates(self, records_ptr, n):
        L.check(L.lib().nls_pso_apply_candidates(self._h, C.c_void_p(records_ptr), n))

    def attach_exchange(self, window):
        L.check(L.lib().nls_pso_attach_exchange(self._h, window.handle))

    def step_fused(self, n=1):
        L.check(L.lib().nls_pso_step_fused(self._h, n))

    def sync(self):
        st = L.Status()
        L.check(L.lib().nls_pso_sync(self._h, C.byref(st)))
        return st.as_dict()

    def _read(self, fn, shape):
        a = np.zeros(shape, self.dt)
        L.check(fn(self._h, a.ctypes.data))
        return a

    def best(self):
        return self._read(L.lib().nls_pso_read_best, self.cfg.dim)

    def positions(self):
        return self._read(L.lib().nls_pso_read_positions, (self.cfg.n_particles, self.cfg.dim))

    def velocities(self):
        return self._read(L.lib().nls_pso_read_velocities, (self.cfg.n_particles, self.cfg.dim))

    def pbest_values(self):
        return self._read(L.lib().nls_pso_read_pbest_values, self.cfg.n_particles)

    def last_values(self):
        return self._read(L.lib().nls_pso_read_last_values, self.cfg.n_particles)

    def close(self):
        if self._h:
            L.lib().nls_pso_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SANNChains:
    """Stepwise handle (nls_sann_*): a batch of independent annealing chains resident in HBM."""

    def __init__(self, ctx, cfg, x0):
        self.ctx, self.cfg = ctx, cfg
        self.dt = np_dtype(cfg.dtype)
        x0 = np.ascontiguousarray(x0, dtype=self.dt)
        count = 1 if x0.ndim == 1 else x0.shape[0]
        assert x0.size == count * cfg.dim
        self._h = C.c_void_p()
        L.check(L.lib().nls_sann_create(ctx.handle, C.byref(cfg), x0.ctypes.data, count, C.byref(self._h)))
        ctx._adopt(self)

    def step(self, n=1):
        """Enqueue n candidates per chain (clamped to what is left of max_iter * (temperature_iter - 1))."""
        L.check(L.lib().nls_sann_step(self._h, n))

    def run(self):
        L.check(L.lib().nls_sann_step(self._h, 0xFFFFFFFFFFFFFFFF))

    def sync(self):
        st = L.Status()
        L.check(L.lib().nls_sann_sync(self._h, C.byref(st)))
        return st.as_dict()

    def best(self):
        x = np.zeros(self.cfg.dim, self.dt)
        L.check(L.lib().nls_sann_read_best(self._h, x.ctypes.data))
        return x

    def chains(self):
        """Per-chain results: x_best / p_cur [n_chains, dim], f_best, n_accepted, n_improved [n_chains]."""
        n, d = self.cfg.n_chains, self.cfg.dim
        a = {"x_best": np.zeros((n, d), self.dt), "f_best": np.zeros(n, self.dt), "p_cur": np.zeros((n, d), self.dt),
             "n_accepted": np.zeros(n, np.uint32), "n_improved": np.zeros(n, np.uint32)}
        L.check(L.lib().nls_sann_read_chains(self._h, a["x_best"].ctypes.data, a["f_best"].ctypes.data,
                                             a["p_cur"].ctypes.data, a["n_accepted"].ctypes.data,
                                             a["n_improved"].ctypes.data))
        return a

    def close(self):
        if self._h:
            L.lib().nls_sann_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ExchangeWindow:
    """Peer-memory exchange window (nls_xchg_*): records and flags in this rank's HBM, mapped by every peer via IPC."""

    def __init__(self, ctx, record_bytes, world, rank):
        self._h = C.c_void_p()
        L.check(L.lib().nls_xchg_create(ctx.handle, record_bytes, world, rank, C.byref(self._h)))
        ctx._adopt(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def ipc_handle(self):
        buf = (C.c_ubyte * L.XCHG_HANDLE_BYTES)()
        L.check(L.lib().nls_xchg_get_handle(self._h, buf))
        return bytes(buf)

    def open_peers(self, handles):
        """handles: rank-ordered list of the bytes objects returned by every rank's ipc_handle()."""
        blob = b"".join(handles)
        L.check(L.lib().nls_xchg_open_peers(self._h, blob))

    def close(self):
        if self._h:
            L.lib().nls_xchg_destroy(self._h)
            self._h = C.c_void_p()


class DeviceGroup:
    """Several GPUs driven from ONE process (nls_group_*): one context per device, peer access enabled between them."""

    def __init__(self, devices):
        devices = list(range(devices)) if isinstance(devices, int) else list(devices)
        arr = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        self.devices = devices
        L.check(L.lib().nls_group_create(len(devices), arr, C.byref(self._h)))
        self._children = weakref.WeakSet()

    @property
    def handle(self):
        return self._h

    def __len__(self):
        return len(self.devices)

    def close(self):
        if self._h:
            for child in list(self._children):
                child.close()
            L.lib().nls_group_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _ShardView(PSOSwarm):
    """A shard of a ShardedSwarm (owned by it): only the read-back calls are meaningful."""

    def __init__(self, handle, cfg):
        self._h, self.cfg, self.dt = handle, cfg, np_dtype(cfg.dtype)

    def close(self):
        self._h = C.c_void_p()


class ShardedSwarm:
    """One swarm over the devices of a group (nls_pso_sharded_*): identical to the same swarm on one GPU."""

    def __init__(self, group, cfg, lower, upper):
        self.group, self.cfg = group, cfg
        self.dt = np_dtype(cfg.dtype)
        lower = np.ascontiguousarray(lower, dtype=self.dt)
        upper = np.ascontiguousarray(upper, dtype=self.dt)
        self._h = C.c_void_p()
        L.check(L.lib().nls_pso_sharded_create(group.handle, C.byref(cfg), lower.ctypes.data, upper.ctypes.data,
                                               C.byref(self._h)))
        group._children.add(self)

    def step(self, n=1):
        L.check(L.lib().nls_pso_sharded_step(self._h, n))

    def sync(self):
        st = L.Status()
        L.check(L.lib().nls_pso_sharded_sync(self._h, C.byref(st)))
        return st.as_dict()

    def best(self):
        x = np.zeros(self.cfg.dim, self.dt)
        L.check(L.lib().nls_pso_sharded_read_best(self._h, x.ctypes.data))
        return x

    def positions(self):
        """particle_positions of the whole swarm, shard after shard (global particle order)."""
        parts, world, n = [], len(self.group), self.cfg.n_particles
        for r in range(world):
            h = C.c_void_p()
            L.check(L.lib().nls_pso_sharded_shard(self._h, r, C.byref(h)))
            base, extra = divmod(n, world)
            local = pso_cfg(self.cfg.dtype, self.cfg.objective, self.cfg.pso_type, True, base + (1 if r < extra else 0),
                            self.cfg.dim)
            parts.append(_ShardView(h, local).positions())
        return np.concatenate(parts)

    def close(self):
        if self._h:
            L.lib().nls_pso_sharded_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _IslandView(DEPopulation):
    def __init__(self, handle, cfg):
        self._h, self.cfg, self.dt = handle, cfg, np_dtype(cfg.dtype)

    def close(self):
        self._h = C.c_void_p()


class DEIslands:
    """One DE island per device of a group with ring migration over NVLink (nls_de_islands_*)."""

    def __init__(self, group, cfg, x0, migrate_every=10, migrants=64):
        self.group, self.cfg = group, cfg
        self.dt = np_dtype(cfg.dtype)
        x0 = np.ascontiguousarray(x0, dtype=self.dt)
        self._h = C.c_void_p()
        L.check(L.lib().nls_de_islands_create(group.handle, C.byref(cfg), x0.ctypes.data, migrate_every, migrants,
                                              C.byref(self._h)))
        group._children.add(self)

    def step(self, n=1):
        L.check(L.lib().nls_de_islands_step(self._h, n))

    def sync(self):
        st = L.Status()
        L.check(L.lib().nls_de_islands_sync(self._h, C.byref(st)))
        return st.as_dict()

    def best(self):
        x = np.zeros(self.cfg.dim, self.dt)
        L.check(L.lib().nls_de_islands_read_best(self._h, x.ctypes.data))
        return x

    def island(self, rank):
        h = C.c_void_p()
        L.check(L.lib().nls_de_islands_island(self._h, rank, C.byref(h)))
        return _IslandView(h, self.cfg)

    def close(self):
        if self._h:
            L.lib().nls_de_islands_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def de_cfg(dtype=L.F64, objective=L.SPHERE, strategy=L.DE_RANDOM, minimize=True, pop_size=50, dim=2,
           crossover_prob=0.9, differential_weight=0.8, eps=10e-4, max_iter=1000, best_val_no_change=50, seed=0,
           agent_offset=0, flags=0):
    return L.DECfg(dtype, objective, strategy, int(minimize), pop_size, dim, crossover_prob, differential_weight, eps,
                   max_iter, best_val_no_change, seed, agent_offset, flags, 0)


def pso_cfg(dtype=L.F64, objective=L.SPHERE, pso_type=L.PSO_VANILLA, minimize=True, n_particles=10, dim=2,
            inertia=0.8, cognitive_coef=1.8, social_coef=1.8, eps=10e-4, max_iter=5000, best_val_no_change=50,
            constrained=False, flags=0, seed=0, particle_offset=0, n_particles_global=0):
    return L.PSOCfg(dtype, objective, pso_type, int(minimize), n_particles, dim, inertia, cognitive_coef, social_coef,
                    eps, max_iter, best_val_no_change, int(constrained), flags, seed, particle_offset,
                    n_particles_global)


def sann_cfg(dtype=L.F64, objective=L.SPHERE, minimize=True, n_chains=1, dim=2, max_iter=5000, temperature_iter=10,
             temperature_max=10.0, seed=0, chain_offset=0, flags=0):
    return L.SANNCfg(dtype, objective, int(minimize), flags, n_chains, dim, max_iter, temperature_iter,
                     temperature_max, seed, chain_offset)


class DE:
    """nlsolver::DE<Callable, RNG, scalar_t, RecombinationType> (nlsolver.h:2379-2410)."""

    def __init__(self, f, generator, crossover_prob=0.9, differential_weight=0.8, eps=10e-4, pop_size=50,
                 max_iter=1000, best_val_no_change=50, scalar_t=np.float64,
                 recombination=RecombinationStrategy.random, ctx=None):
        self.f, self.generator = f, generator
        self.crossover_prob, self.differential_weight, self.eps = crossover_prob, differential_weight, eps
        self.pop_size, self.max_iter, self.best_val_no_change = pop_size, max_iter, best_val_no_change
        self.scalar_t, self.recombination = scalar_t, recombination
        self.ctx = ctx

    def _solve(self, x, minimize):
        ctx = self.ctx or default_context()
        dt = np.dtype(self.scalar_t)
        cfg = de_cfg(nls_dtype(dt), self.f, self.recombination, minimize, self.pop_size, len(x), self.crossover_prob,
                     self.differential_weight, self.eps, self.max_iter, self.best_val_no_change,
                     seed_from_generator(self.generator))
        x0 = np.ascontiguousarray(x, dtype=dt)
        out = np.zeros(len(x), dt)
        st = L.Status()
        L.check(L.lib().nls_de_solve(ctx.handle, C.byref(cfg), x0.ctypes.data, out.ctypes.data, C.byref(st)))
        x[:] = out.tolist() if isinstance(x, list) else out
        return SolverStatus(dt.type(st.f_value), st.iterations, st.function_calls)

    def minimize(self, x):
        return self._solve(x, True)

    def maximize(self, x):
        return self._solve(x, False)


class PSO:
    """nlsolver::PSO<Callable, RNG, scalar_t, Type> (nlsolver.h:2498-2589)."""

    def __init__(self, f, generator, inertia=0.8, cognitive_coef=1.8, social_coef=1.8, n_particles=10, max_iter=5000,
                 best_val_no_change=50, eps=10e-4, scalar_t=np.float64, pso_type=PSOType.Vanilla, ctx=None):
        self.f, self.generator = f, generator
        self.inertia, self.cognitive_coef, self.social_coef = inertia, cognitive_coef, social_coef
        self.n_particles, self.max_iter, self.best_val_no_change, self.eps = (n_particles, max_iter,
                                                                            best_val_no_change, eps)
        self.scalar_t, self.pso_type = scalar_t, pso_type
        self.ctx = ctx

    def _solve(self, x, lower, upper, minimize):
        ctx = self.ctx or default_context()
        dt = np.dtype(self.scalar_t)
        d = len(x)
        constrained = lower is not None
        if not constrained:   # nlsolver.h:2553-2563: lower = -|x|, upper = |x|, no clamping
            upper = np.abs(np.asarray(x, dtype=dt))
            lower = -upper
        flags = 0
        if self.pso_type == PSOType.Vanilla and self.n_particles > d:
            flags |= L.FLAG_SOCIAL_INDEX_J   # the reference reads out of bounds here (nlsolver.h:2674)
        cfg = pso_cfg(nls_dtype(dt), self.f, self.pso_type, minimize, self.n_particles, d, self.inertia,
                      self.cognitive_coef, self.social_coef, self.eps, self.max_iter, self.best_val_no_change,
                      constrained, flags, seed_from_generator(self.generator))
        lo = np.ascontiguousarray(lower, dtype=dt)
        up = np.ascontiguousarray(upper, dtype=dt)
        out = np.zeros(d, dt)
        st = L.Status()
        L.check(L.lib().nls_pso_solve(ctx.handle, C.byref(cfg), lo.ctypes.data, up.ctypes.data, out.ctypes.data,
                                      C.byref(st)))
        if st.best_valid:
            x[:] = out.tolist() if isinstance(x, list) else out
        elif isinstance(x, list):
            del x[:]          # the reference assigns an empty swarm_best_position (nlsolver.h:2601)
        return SolverStatus(dt.type(st.f_value), st.iterations, st.function_calls)

    def minimize(self, x, lower=None, upper=None):
        return self._solve(x, lower, upper, True)

    def maximize(self, x, lower=None, upper=None):
        return self._solve(x, lower, upper, False)


class SANN:
    """nlsolver::SANN<Callable, RNG, scalar_t> (nlsolver.h:2744-2776): `minimize(x)` runs one chain from x.

    `minimize_batch(xs)` / `maximize_batch(xs)` are the batch form this engine adds: one independent chain per row of
    `xs` (or `n_chains` chains from one shared start), each the reference loop on its own draw stream."""

    def __init__(self, f, generator, max_iter=5000, temperature_iter=10, temperature_max=10.0, scalar_t=np.float64,
                 ctx=None):
        self.f, self.generator = f, generator
        self.max_iter, self.temperature_iter, self.temperature_max = max_iter, temperature_iter, temperature_max
        self.scalar_t, self.ctx = scalar_t, ctx
        self.f_evals = 0          # accumulates over calls, like the reference member (nlsolver.h:2751, 2783)

    def _cfg(self, n_chains, dim, minimize):
        return sann_cfg(nls_dtype(np.dtype(self.scalar_t)), self.f, minimize, n_chains, dim, self.max_iter,
                        self.temperature_iter, self.temperature_max, seed_from_generator(self.generator))

    def _solve(self, x, minimize, n_chains=1):
        ctx = self.ctx or default_context()
        dt = np.dtype(self.scalar_t)
        cfg = self._cfg(n_chains, len(x), minimize)
        x0 = np.ascontiguousarray(x, dtype=dt)
        out = np.zeros(len(x), dt)
        st = L.Status()
        L.check(L.lib().nls_sann_solve(ctx.handle, C.byref(cfg), x0.ctypes.data, 1, out.ctypes.data, C.byref(st)))
        x[:] = out.tolist() if isinstance(x, list) else out
        self.f_evals += st.function_calls
        return SolverStatus(dt.type(st.f_value), st.iterations, self.f_evals)

    def minimize(self, x):
        return self._solve(x, True)

    def maximize(self, x):
        return self._solve(x, False)

    def minimize_multistart(self, x, n_chains):
        """n_chains chains from the same start x; x receives the best chain's point."""
        return self._solve(x, True, n_chains)

    def maximize_multistart(self, x, n_chains):
        return self._solve(x, False, n_chains)

    def _batch(self, xs, n_chains, minimize):
        ctx = self.ctx or default_context()
        dt = np.dtype(self.scalar_t)
        xs = np.ascontiguousarray(xs, dtype=dt)
        n = xs.shape[0] if xs.ndim == 2 else int(n_chains)
        chains = SANNChains(ctx, self._cfg(n, xs.shape[-1], minimize), xs)
        try:
            chains.run()
            st = chains.sync()
            res = chains.chains()
        finally:
            chains.close()
        per_chain = st["function_calls"] // n
        self.f_evals += st["function_calls"]
        return res["x_best"], [SolverStatus(dt.type(f), st["iterations"], per_chain) for f in res["f_best"]]

    def minimize_batch(self, xs, n_chains=None):
        """xs: [n_chains, dim] start points, or [dim] with n_chains.  Returns (best points, list of SolverStatus)."""
        return self._batch(xs, n_chains, True)

    def maximize_batch(self, xs, n_chains=None):
        return self._batch(xs, n_chains, False)


def nmpso_cfg(dtype=L.F64, objective=L.SPHERE, minimize=True, n_solvers=1, dim=2, alpha=1.0, gamma=2.0, rho=0.5, sigma=0.5,
              inertia=0.8, cognitive_coef=1.8, social_coef=1.8, eps=1e-6, max_iter=1000, no_change_best_iter=20, seed=0,
              solver_offset=0, flags=0):
    return L.NMPSOCfg(dtype, objective, int(minimize), flags, n_solvers, dim, alpha, gamma, rho, sigma, inertia,
                      cognitive_coef, social_coef, eps, max_iter, no_change_best_iter, seed, solver_offset)


def nmpso_solve(ctx, cfg, x0):
    """nls_nmpso_solve: x0 is [dim] (every solver starts there) or [n_solvers, dim].
    Returns (status dict, {"x_best" [n, dim], "f_best" [n], "iterations" [n], "function_calls" [n]})."""
    dt = np_dtype(cfg.dtype)
    x0 = np.ascontiguousarray(x0, dtype=dt)
    count = 1 if x0.ndim == 1 else x0.shape[0]
    n, d = cfg.n_solvers, cfg.dim
    a = {"x_best": np.zeros((n, d), dt), "f_best": np.zeros(n, dt), "iterations": np.zeros(n, np.uint64),
         "function_calls": np.zeros(n, np.uint64)}
    st = L.Status()
    L.check(L.lib().nls_nmpso_solve(ctx.handle, C.byref(cfg), x0.ctypes.data, count, a["x_best"].ctypes.data,
                                    a["f_best"].ctypes.data, a["iterations"].ctypes.data,
                                    a["function_calls"].ctypes.data, C.byref(st)))
    return st.as_dict(), a


class NelderMeadPSO:
    """nlsolver::NelderMeadPSO<Callable, RNG, scalar_t> (nlsolver.h:3546-3614): `minimize(x)` / `maximize(x)` run one
    solver from x.  `minimize_batch(xs)` is the batch form this engine adds: one independent solver per row of xs (or
    n_solvers solvers from one point), each the reference's loop on its own draw stream.  The bounded overloads of the
    reference read its bounds out of range (nlsolver.h:3859) and are not offered."""

    def __init__(self, f, generator, alpha=1.0, gamma=2.0, rho=0.5, sigma=0.5, inertia=0.8, cognitive_coef=1.8,
                 social_coef=1.8, eps=1e-6, max_iter=1000, no_change_best_iter=20, scalar_t=np.float64, ctx=None):
        self.f, self.generator = f, generator
        self.p = dict(alpha=alpha, gamma=gamma, rho=rho, sigma=sigma, inertia=inertia, cognitive_coef=cognitive_coef,
                      social_coef=social_coef, eps=eps, max_iter=max_iter, no_change_best_iter=no_change_best_iter)
        self.scalar_t, self.ctx = scalar_t, ctx

    def _run(self, xs, n_solvers, minimize):
        ctx = self.ctx or default_context()
        dt = np.dtype(self.scalar_t)
        xs = np.ascontiguousarray(xs, dtype=dt)
        n = xs.shape[0] if xs.ndim == 2 else int(n_solvers)
        cfg = nmpso_cfg(nls_dtype(dt), self.f, minimize, n, xs.shape[-1], seed=seed_from_generator(self.generator),
                        **self.p)
        return nmpso_solve(ctx, cfg, xs), dt

    def _solve(self, x, minimize):
        (st, a), dt = self._run(np.asarray(x, dtype=np.dtype(self.scalar_t)), 1, minimize)
        x[:] = a["x_best"][0].tolist() if isinstance(x, list) else a["x_best"][0]
        return SolverStatus(dt.type(a["f_best"][0]), int(a["iterations"][0]), int(a["function_calls"][0]))

    def minimize(self, x):
        return self._solve(x, True)

    def maximize(self, x):
        return self._solve(x, False)

    def minimize_batch(self, xs, n_solvers=None, minimize=True):
        """xs: [n_solvers, dim] start points, or [dim] with n_solvers.  Returns (best points, list of SolverStatus)."""
        (st, a), dt = self._run(xs, n_solvers, minimize)
        return a["x_best"], [SolverStatus(dt.type(f), int(i), int(c))
                             for f, i, c in zip(a["f_best"], a["iterations"], a["function_calls"])]

    def maximize_batch(self, xs, n_solvers=None):
        return self.minimize_batch(xs, n_solvers, minimize=False)


DESolver, PSOSolver = DE, PSO   # the names README.md:80,99 uses
