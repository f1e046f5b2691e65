"""Shared helpers for the -m gpu parity tests: everything goes through the C ABI (nlsolver_b200._lib)."""
import numpy as np

import nlsolver_b200 as nb
from oracle import binding as B


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32}[a.dtype.itemsize]) if a.dtype.kind == "f" else a


def rel_close(a, b, tol):
    """|a - b| <= tol * scale.  Rows (2-D): scale is the agent's own max-norm — north_star's "per-agent positions
    within 1e-12 relative"; sums with cancellation put single coordinates near zero, where an elementwise relative
    test is meaningless.  Values (0-D / 1-D): elementwise relative with a floor of 1 (objectives cancel near optima)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if b.ndim == 2:
        scale = np.maximum(np.max(np.abs(b), axis=1, keepdims=True), 1e-300)
    else:
        scale = np.maximum(np.abs(b), 1.0)
    return bool(np.all(np.abs(a - b) <= tol * scale))


def tolerance(dtype, objective):
    """Sphere / Rosenbrock use only + - * in the canonical order: bit-exact (tol 0).  Rastrigin / Ackley call cos /
    exp / sqrt, where CUDA libdevice and glibc differ in the last ulp: north_star's 1e-12 relative in fp64."""
    exact = objective in (B.SPHERE, B.ROSENBROCK, B.ROSENBROCK_EX)
    if dtype == B.F64:
        return 0.0 if exact else 1e-12
    return 0.0 if exact else 1e-5


def oracle_de(lib, dtype, obj, strategy, minimize, P, d, g, seed, x0, cr=0.9, f=0.8, offset=0, masks=True):
    cfg = B.de_cfg(dtype=dtype, objective=obj, strategy=strategy, minimize=minimize, pop_size=P, dim=d,
                   crossover_prob=cr, differential_weight=f, eps=0.0, max_iter=g, best_val_no_change=1 << 40,
                   seed=seed, agent_offset=offset)
    return B.de_run(lib, cfg, x0, masks=masks)


def gpu_de(ctx, dtype, obj, strategy, minimize, P, d, seed, x0, cr=0.9, f=0.8, offset=0, masks=True, max_iter=1 << 40,
           eps=0.0, vnc=1 << 40):
    cfg = nb.de_cfg(dtype=dtype, objective=obj, strategy=strategy, minimize=minimize, pop_size=P, dim=d,
                    crossover_prob=cr, differential_weight=f, eps=eps, max_iter=max_iter, best_val_no_change=vnc,
                    seed=seed, agent_offset=offset, flags=nb.FLAG_RECORD_MASKS if masks else 0)
    return nb.DEPopulation(ctx, cfg, x0)
