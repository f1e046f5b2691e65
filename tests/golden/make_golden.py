"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libnls_ref.so, built from /root/reference
by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py
Each fixture stores the configuration, x0 / bounds and everything the reference produced on the draw tape, so the
oracle and the CUDA path can be checked against the reference where the reference itself cannot travel."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import binding as B  # noqa: E402

DE_GOLDEN = {
    # name: (dtype, objective, strategy, minimize, P, d, G, scale, seed)
    "de_rand_sphere_f64": (B.F64, B.SPHERE, B.DE_RANDOM, True, 96, 10, 8, 10.24, 11),
    "de_best_rosenbrock_f64": (B.F64, B.ROSENBROCK, B.DE_BEST, True, 64, 8, 25, 4.096, 12),
    "de_rand_rastrigin_f64": (B.F64, B.RASTRIGIN, B.DE_RANDOM, True, 130, 67, 6, 10.24, 13),
    "de_rand_ackley_max_f64": (B.F64, B.ACKLEY, B.DE_RANDOM, False, 48, 5, 10, 65.536, 14),
    "de_rand_sphere_f32": (B.F32, B.SPHERE, B.DE_RANDOM, True, 80, 13, 8, 10.24, 15),
}
PSO_GOLDEN = {
    # name: (dtype, objective, type, minimize, constrained, P, d, G, bound, seed)
    "pso_vanilla_sphere_f64": (B.F64, B.SPHERE, B.PSO_VANILLA, True, False, 16, 16, 20, 10.24, 21),
    "pso_vanilla_bounded_rosenbrock_f64": (B.F64, B.ROSENBROCK, B.PSO_VANILLA, True, True, 6, 8, 30, 4.096, 22),
    "pso_accel_ackley_f64": (B.F64, B.ACKLEY, B.PSO_ACCELERATED, True, False, 200, 32, 10, 32.768, 23),
    "pso_accel_bounded_rastrigin_f64": (B.F64, B.RASTRIGIN, B.PSO_ACCELERATED, True, True, 40, 8, 40, 5.12, 24),
    "pso_accel_sphere_f32": (B.F32, B.SPHERE, B.PSO_ACCELERATED, True, False, 64, 16, 10, 10.24, 25),
}
SANN_GOLDEN = {
    # name: (dtype, objective, minimize, n_chains, d, max_iter, temperature_iter, temperature_max, start, seed)
    "sann_rosenbrock_ex_f64": (B.F64, B.ROSENBROCK_EX, True, 12, 2, 400, 10, 10.0, 5.0, 31),
    "sann_rastrigin_f64": (B.F64, B.RASTRIGIN, True, 9, 37, 120, 6, 4.0, 2.5, 32),
    "sann_ackley_max_f64": (B.F64, B.ACKLEY, False, 7, 11, 150, 10, 10.0, 1.0, 33),
    "sann_sphere_f32": (B.F32, B.SPHERE, True, 10, 21, 150, 10, 10.0, 3.0, 34),
}


NMPSO_GOLDEN = {
    # name: (dtype, objective, minimize, n_solvers, d, max_iter, spread, seed)   d: shapes the reference survives (oracle_abi.h)
    "nmpso_rosenbrock_ex_f64": (B.F64, B.ROSENBROCK_EX, True, 24, 2, 1000, 3.0, 41),
    "nmpso_rastrigin_f64": (B.F64, B.RASTRIGIN, True, 10, 6, 400, 3.0, 42),
    "nmpso_sphere_max_f64": (B.F64, B.SPHERE, False, 8, 4, 60, 2.0, 43),
    "nmpso_rosenbrock_f64": (B.F64, B.ROSENBROCK, True, 6, 16, 300, 1.5, 44),
    "nmpso_sphere_f32": (B.F32, B.SPHERE, True, 12, 8, 300, 2.0, 45),
}


def nmpso_start(n, d, spread, seed, dtype):
    return np.random.default_rng(seed).uniform(-spread, spread, size=(n, d)).astype(B.np_dtype(dtype))


def write_nmpso(ref):
    for name, (dtype, obj, mini, n, d, it, spread, seed) in NMPSO_GOLDEN.items():
        cfg = B.nmpso_cfg(dtype=dtype, objective=obj, minimize=mini, n_solvers=n, dim=d, max_iter=it, seed=seed)
        x0 = nmpso_start(n, d, spread, seed, dtype)
        st, a = B.nmpso_run(ref, cfg, x0)
        assert st is not None, name
        # ties are detected by the restatement (the reference cannot report them): fixtures must be tie-free
        so, ao = B.nmpso_run(B.oracle(), B.nmpso_cfg(dtype=dtype, objective=obj, minimize=mini, n_solvers=n, dim=d,
                                                     max_iter=it, seed=seed), x0)
        assert not ao["ties"].any(), (name, "a sort compared equal values: pick another seed")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind="nmpso",
                            cfg=np.array([dtype, obj, int(mini), n, d, it, seed], dtype=np.int64), x0=x0,
                            f_value=st["f_value"], best_index=st["best_index"], x_best=a["x_best"], f_best=a["f_best"],
                            draws=a["draws"], iterations=a["iterations"], function_calls=a["function_calls"])
    print("wrote", len(NMPSO_GOLDEN), "NelderMeadPSO fixtures to", HERE)


def sann_start(n_chains, d, start, dtype):
    # one start point per chain: start * (1 + chain / 8) on even coordinates, -start on odd ones
    x0 = np.empty((n_chains, d), dtype=B.np_dtype(dtype))
    x0[:, 0::2] = (start * (1 + np.arange(n_chains) / 8.0))[:, None]
    x0[:, 1::2] = -start
    return x0


def main():
    ref = B.reference()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    if "--only-nmpso" in sys.argv:      # leaves the other fixtures' files untouched
        write_nmpso(ref)
        return
    for name, (dtype, obj, strat, mini, P, d, G, scale, seed) in DE_GOLDEN.items():
        cfg = B.de_cfg(dtype=dtype, objective=obj, strategy=strat, minimize=mini, pop_size=P, dim=d, eps=0.0,
                       max_iter=G, best_val_no_change=1 << 40, seed=seed)
        x0 = np.full(d, scale, dtype=B.np_dtype(dtype))
        st, a = B.de_run(ref, cfg, x0, masks=True)
        assert ref.ref_last_inconsistencies() == 0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind="de",
                            cfg=np.array([dtype, obj, strat, int(mini), P, d, G, seed], dtype=np.int64), x0=x0,
                            f_value=st["f_value"], iterations=st["iterations"], function_calls=st["function_calls"],
                            draws_consumed=st["draws_consumed"], best_index=st["best_index"], **a)
    for name, (dtype, obj, ptype, mini, con, P, d, G, bound, seed) in PSO_GOLDEN.items():
        cfg = B.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, minimize=mini, n_particles=P, dim=d, eps=0.0,
                        max_iter=G, best_val_no_change=1 << 40, constrained=con, seed=seed)
        up = np.full(d, bound, dtype=B.np_dtype(dtype))
        st, a = B.pso_run(ref, cfg, -up, up)
        a.pop("velocities")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind="pso",
                            cfg=np.array([dtype, obj, ptype, int(mini), int(con), P, d, G, seed], dtype=np.int64),
                            upper=up, f_value=st["f_value"], iterations=st["iterations"],
                            function_calls=st["function_calls"], draws_consumed=st["draws_consumed"],
                            best_valid=st["best_valid"], **a)
    for name, (dtype, obj, mini, n, d, it, ti, tmax, start, seed) in SANN_GOLDEN.items():
        cfg = B.sann_cfg(dtype=dtype, objective=obj, minimize=mini, n_chains=n, dim=d, max_iter=it, temperature_iter=ti,
                         temperature_max=tmax, seed=seed)
        x0 = sann_start(n, d, start, dtype)
        st, a = B.sann_run(ref, cfg, x0)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), kind="sann",
                            cfg=np.array([dtype, obj, int(mini), n, d, it, ti, seed], dtype=np.int64), tmax=tmax,
                            x0=x0, f_value=st["f_value"], best_index=st["best_index"],
                            function_calls_total=st["function_calls"], x_best=a["x_best"], f_best=a["f_best"],
                            draws=a["draws"], iterations=a["iterations"], function_calls=a["function_calls"])
    print("wrote", len(DE_GOLDEN) + len(PSO_GOLDEN) + len(SANN_GOLDEN), "fixtures to", HERE)
    write_nmpso(ref)


if __name__ == "__main__":
    main()
