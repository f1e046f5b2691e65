"""Small populations: the one-launch path (all generations of a step in one kernel on one thread-block cluster) against
the graph-replay path, and BASELINE.json configs[0] (the reference's own CPU-runnable case: DE-random on the README's
2-D Rosenbrock, default population 50, x0 = {5, 7}) end to end against the reference's CPU time on the same host.

    python tools/bench_small.py            -> one JSON line
"""
import ctypes as C
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nlsolver_b200 as nb  # noqa: E402

NEVER = 1 << 40


def per_generation_us(make, gens, reps=5):
    """median device microseconds per generation of `gens` generations enqueued in ONE step call (CUDA events)."""
    stream = torch.cuda.Stream()
    ctx = nb.Context(0, stream.cuda_stream)
    h = make(ctx)
    h.step(gens)
    h.sync()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h.step(gens)
        e1.record(stream)
        torch.cuda.synchronize()
        out.append(1e3 * e0.elapsed_time(e1) / gens)
    st = h.sync()
    h.close()
    ctx.close()
    return statistics.median(out), st["iterations"]


def sweep_point(solver, dtype, P, d=64, gens=512):
    x0 = np.full(d, 10.24)

    def make(ctx):
        if solver == "DE-random":
            return nb.DEPopulation(ctx, nb.de_cfg(dtype=dtype, objective=nb.SPHERE, pop_size=P, dim=d, eps=0.0,
                                                  max_iter=NEVER, best_val_no_change=NEVER, seed=1), x0)
        ptype = nb.PSO_VANILLA if solver == "PSO-vanilla" else nb.PSO_ACCELERATED
        return nb.PSOSwarm(ctx, nb.pso_cfg(dtype=dtype, objective=nb.SPHERE, pso_type=ptype, n_particles=P, dim=d,
                                           eps=0.0, max_iter=NEVER, best_val_no_change=NEVER, seed=1,
                                           flags=nb.FLAG_SOCIAL_INDEX_J), -x0, x0)
    rec = {"solver": solver, "dtype": "fp64" if dtype == nb.F64 else "fp32", "pop": P, "dim": d}
    for name, env in (("one_launch_us_per_generation", "1"), ("graph_replay_us_per_generation", "0")):
        os.environ["NLS_DE_ONE_LAUNCH"] = env
        rec[name], _ = per_generation_us(make, gens)
    os.environ.pop("NLS_DE_ONE_LAUNCH", None)
    rec["agent_evals_per_sec"] = P / (rec["one_launch_us_per_generation"] * 1e-6)
    return rec


def config1():
    """DE<RosenbrockExample, xorshift<double>>, defaults (pop 50, CR 0.9, F 0.8, eps 1e-3, max_iter 1000, no-change 50),
    x0 = {5, 7} through the Python mirror of the header: wall time of one minimize() call, warm context."""
    class XorShift:   # nlsolver::rng::xorshift<double> (nlsolver.h:1343-1381), host side: only seeds the draw tape
        def __init__(self):
            self.x = [0x7c26ca28fb68bc1b, 0x7c26ca28]

        def __call__(self):
            t, s = self.x
            self.x[0] = s
            t ^= (t << 23) & 0xFFFFFFFFFFFFFFFF
            t ^= t >> 18
            t ^= s ^ (s >> 5)
            self.x[1] = t
            return ((t + s) & 0xFFFFFFFFFFFFFFFF) / 18446744073709551615.0
    ctx = nb.Context(0)
    times, last = [], None
    for k in range(int(os.environ.get("NLS_BENCH_SMALL_CALLS", "60"))):
        x = [5.0, 7.0]
        solver = nb.DE(nb.RosenbrockExample, XorShift(), ctx=ctx)
        t0 = time.perf_counter()
        st = solver.minimize(x)
        times.append(time.perf_counter() - t0)
        last = (st, x)
    ctx.close()
    st, x = last
    out = {"gpu_wall_us_per_minimize_median": 1e6 * statistics.median(times[10:]),
           "gpu_wall_us_per_minimize_first": 1e6 * times[0],
           "iterations": st.iteration, "function_calls": st.function_calls_used, "f_value": float(st.f_value),
           "x": [float(v) for v in x]}
    try:
        from oracle import binding as B
        ref = B.reference()
        if ref is not None:
            cfg = B.de_cfg(objective=B.ROSENBROCK_EX, strategy=B.DE_RANDOM, pop_size=50, dim=2)
            x0 = np.array([5.0, 7.0])
            secs = []
            for _ in range(50):
                sec, rst = C.c_double(), B.Status()
                ref.ref_de_time(C.byref(cfg), x0.ctypes.data, C.byref(sec), C.byref(rst))
                secs.append(sec.value)
            out["reference_cpu_us_per_minimize_median"] = 1e6 * statistics.median(secs)
            out["reference_iterations"] = rst.iterations
            out["reference_function_calls"] = rst.function_calls
    except Exception as exc:      # the CPU leg is optional
        out["reference_cpu_error"] = str(exc)[:200]
    return out


def main():
    res = {"config1_readme_de": config1(), "sweep_d64_small": []}
    os.environ["NLS_DE_TINY"] = "0"          # the same call through the general kernels (one-launch cluster path)
    res["config1_readme_de_general_kernels"] = {k: v for k, v in config1().items() if k.startswith("gpu_") or k == "iterations"}
    os.environ.pop("NLS_DE_TINY")
    for P, d in ((50, 2),):
        res["sweep_d64_small"].append(sweep_point("DE-random", nb.F64, P, d=d, gens=256))
    for solver in ("DE-random", "PSO-vanilla", "PSO-accelerated"):
        for P in (1 << 10, 1 << 12):
            res["sweep_d64_small"].append(sweep_point(solver, nb.F64, P))
    res["sweep_d64_small"].append(sweep_point("DE-random", nb.F32, 1 << 10))
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
