"""Every kernel of libnls_b200.so once, at small awkward shapes (rows that end inside a vector, populations that do not
fill a block, lane groups of every width) — meant to run under compute-sanitizer where that is available:

    compute-sanitizer --tool memcheck --error-exitcode 3 python tests/tools/sanitize_driver.py

and, where it is not (it is closed on the B200 pool this was developed on), with the library's own guard zones:

    NLS_B200_GUARD=1 python tests/tools/sanitize_driver.py      # exits 3 if any buffer's guard zone was overwritten

No torch import (keeps the instrumented process small); device scratch for the island hooks comes from cudaMalloc
through ctypes."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nlsolver_b200 as nb  # noqa: E402

rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]
NEVER = 1 << 40


def dmalloc(nbytes):
    p = C.c_void_p()
    assert rt.cudaMalloc(C.byref(p), nbytes) == 0
    return p


ctx = nb.Context(0)
n_launch_groups = 0

# ---- DE: both strategies, both dtypes, short and multi-sweep rows, high acceptance (repair rounds), masks ----
for dtype, obj, strat, P, d, gens in ((nb.F64, nb.RASTRIGIN, nb.DE_RANDOM, 300, 37, 3),
                                      (nb.F32, nb.ROSENBROCK, nb.DE_BEST, 129, 5, 4),
                                      (nb.F64, nb.SPHERE, nb.DE_RANDOM, 1000, 3, 6),
                                      (nb.F64, nb.ACKLEY, nb.DE_BEST, 77, 131, 2),
                                      (nb.F32, nb.SPHERE, nb.DE_RANDOM, 513, 70, 2),
                                      (nb.F64, nb.BEALE, nb.DE_RANDOM, 50, 2, 5),
                                      (nb.F64, nb.SHEKEL, nb.DE_BEST, 50, 4, 5)):
    cfg = nb.de_cfg(dtype=dtype, objective=obj, strategy=strat, pop_size=P, dim=d, eps=0.0, max_iter=NEVER,
                    best_val_no_change=NEVER, seed=7, flags=nb.FLAG_RECORD_MASKS)
    pop = nb.DEPopulation(ctx, cfg, np.full(d, 4.0))
    pop.step(gens)
    st = pop.sync()
    assert st["iterations"] == gens
    pop.population(), pop.scores(), pop.best(), pop.decisions(masks=True), pop.rows(P // 2, 3)
    es = 8 if dtype == nb.F64 else 4
    k = 5
    rows, scores, rec = dmalloc(k * d * es), dmalloc(k * es), dmalloc(nb.lib().nls_record_bytes(dtype, d))
    pop.export_best(rec.value)
    pop.export_top(k, rows.value, scores.value)
    pop.import_migrants(k, rows.value, scores.value)
    pop.step(1)
    pop.sync()
    pop.close()
    for p in (rows, scores, rec):
        rt.cudaFree(p)
    n_launch_groups += 1

# ---- PSO: both types, clamped / unclamped, CUDA-graph replay (>= 8 generations), sharded calls, fused exchange ----
for dtype, obj, ptype, P, d, con in ((nb.F64, nb.ACKLEY, nb.PSO_ACCELERATED, 301, 37, False),
                                     (nb.F64, nb.SPHERE, nb.PSO_VANILLA, 7, 9, True),
                                     (nb.F32, nb.RASTRIGIN, nb.PSO_ACCELERATED, 130, 5, True),
                                     (nb.F32, nb.SPHERE, nb.PSO_VANILLA, 200, 131, False),
                                     (nb.F64, nb.BOOTH, nb.PSO_ACCELERATED, 40, 2, False)):
    up = np.full(d, 3.0)
    cfg = nb.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, n_particles=P, dim=d, eps=0.0, max_iter=NEVER,
                     best_val_no_change=NEVER, constrained=con, seed=3, flags=nb.FLAG_SOCIAL_INDEX_J)
    sw = nb.PSOSwarm(ctx, cfg, -up, up)
    sw.step(11)
    st = sw.sync()
    assert st["iterations"] == 11
    sw.positions(), sw.pbest_values(), sw.last_values(), sw.best()
    rec = dmalloc(nb.lib().nls_record_bytes(dtype, d))
    sw.step_local(rec.value)
    sw.apply_candidates(rec.value, 1)
    win = nb.ExchangeWindow(ctx, nb.lib().nls_record_bytes(dtype, d), 1, 0)
    sw.attach_exchange(win)
    sw.step_fused(2)
    assert sw.sync()["iterations"] == 14
    sw.close()
    win.close()
    rt.cudaFree(rec)
    n_launch_groups += 1

# ---- SANN chains: every lane-group width, multi-sweep rows, closed forms, launch cut points ----
for dtype, obj, n, d in ((nb.F64, nb.RASTRIGIN, 33, 3), (nb.F64, nb.SPHERE, 10, 13), (nb.F32, nb.ACKLEY, 21, 50),
                         (nb.F64, nb.ROSENBROCK, 9, 64), (nb.F64, nb.RASTRIGIN, 5, 131), (nb.F32, nb.ROSENBROCK_EX, 300, 7),
                         (nb.F64, nb.MATYAS, 12, 2)):
    x0 = np.random.default_rng(d).uniform(-2, 2, size=(n, d))
    ch = nb.SANNChains(ctx, nb.sann_cfg(dtype=dtype, objective=obj, n_chains=n, dim=d, max_iter=12, seed=5), x0)
    ch.step(7)
    ch.sync()
    ch.run()
    st = ch.sync()
    assert st["stopped"] == 1 and st["function_calls"] == n * (1 + 12 * 9)
    ch.chains(), ch.best()
    ch.close()
    n_launch_groups += 1

# ---- DE with an exchange window (the commit kernel publishes the island record), high acceptance (repair iterations
#      behind the coarse bitmaps), long rows (TMA-staged K2, four-step re-evaluation), one-shot solves (tiny solver) ----
for dtype, obj, strat, P, d, F in ((nb.F64, nb.SPHERE, nb.DE_RANDOM, 3001, 37, 0.3), (nb.F32, nb.SPHERE, nb.DE_BEST, 1500, 300, 0.3),
                                   (nb.F64, nb.RASTRIGIN, nb.DE_RANDOM, 700, 1000, 0.2)):
    cfg = nb.de_cfg(dtype=dtype, objective=obj, strategy=strat, pop_size=P, dim=d, differential_weight=F, eps=0.0,
                    max_iter=NEVER, best_val_no_change=NEVER, seed=11)
    pop = nb.DEPopulation(ctx, cfg, np.full(d, 3.0))
    win = nb.ExchangeWindow(ctx, nb.lib().nls_record_bytes(dtype, d), 1, 0)
    pop.attach_exchange(win)
    pop.step(1)
    pop.step(12)
    st = pop.sync()
    assert st["iterations"] == 13 and st["accepted_total"] > 0
    pop.read_exchange(1)
    pop.close()
    win.close()
    n_launch_groups += 1
for dtype in (nb.F64, nb.F32):
    x = np.array([5.0, 7.0])
    st = nb.DE(nb.RosenbrockExample, lambda: 0.5, scalar_t=np.float64 if dtype == nb.F64 else np.float32, ctx=ctx).minimize(x)
    assert st.iteration > 0
    n_launch_groups += 1

# ---- NelderMeadPSO batches: one warp per solver ----
for dtype, obj, n, d in ((nb.F64, nb.SPHERE, 33, 5), (nb.F32, nb.ROSENBROCK, 10, 12), (nb.F64, nb.RASTRIGIN, 7, 100)):
    xs = np.random.default_rng(d).uniform(-2, 2, size=(n, d))
    st, a = nb.nmpso_solve(ctx, nb.nmpso_cfg(dtype=dtype, objective=obj, n_solvers=n, dim=d, max_iter=40, seed=9), xs)
    assert a["x_best"].shape == (n, d)
    n_launch_groups += 1

ctx.close()
bad = nb.lib().nls_debug_guard_violations()
print(f"sanitize driver: {n_launch_groups} solver configurations completed; guard mode "
      f"{'on' if os.environ.get('NLS_B200_GUARD') == '1' else 'off'}, {bad} guard-zone violations")
sys.exit(3 if bad else 0)
