"""Loads the committed reference fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

from oracle import binding as B

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files(kind):
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if os.path.basename(p).startswith(kind))


def load_de(path):
    z = np.load(path)
    dtype, obj, strat, mini, P, d, G, seed = (int(v) for v in z["cfg"])
    cfg = B.de_cfg(dtype=dtype, objective=obj, strategy=strat, minimize=bool(mini), pop_size=P, dim=d, eps=0.0,
                   max_iter=G, best_val_no_change=1 << 40, seed=seed)
    return cfg, z["x0"], z


def load_pso(path):
    z = np.load(path)
    dtype, obj, ptype, mini, con, P, d, G, seed = (int(v) for v in z["cfg"])
    cfg = B.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, minimize=bool(mini), n_particles=P, dim=d, eps=0.0,
                    max_iter=G, best_val_no_change=1 << 40, constrained=bool(con), seed=seed)
    return cfg, z["upper"], z


def load_sann(path):
    z = np.load(path)
    dtype, obj, mini, n, d, it, ti, seed = (int(v) for v in z["cfg"])
    cfg = B.sann_cfg(dtype=dtype, objective=obj, minimize=bool(mini), n_chains=n, dim=d, max_iter=it,
                     temperature_iter=ti, temperature_max=float(z["tmax"]), seed=seed)
    return cfg, z["x0"], z


def load_nmpso(path):
    z = np.load(path)
    dtype, obj, mini, n, d, it, seed = (int(v) for v in z["cfg"])
    cfg = B.nmpso_cfg(dtype=dtype, objective=obj, minimize=bool(mini), n_solvers=n, dim=d, max_iter=it, seed=seed)
    return cfg, z["x0"], z
