"""DE on the B200 (through the C ABI) against the oracle on the same draw tape, generation by generation.

Bit-exact: donor ids, dim, rejected proposals, crossover masks, accept flags, iteration / call counters, best index.
Values: bit-exact for Sphere / Rosenbrock; <= 1e-12 relative (fp64) where libm is involved."""
import os

import numpy as np
import pytest

import nlsolver_b200 as nb
from oracle import binding as B
from tests.golden_util import golden_files, load_de
from tests.gpu_util import bits, gpu_de, oracle_de, rel_close, tolerance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context(0)
    yield c
    c.close()


def compare_generation(pop, st, so, ao, tol, g, check_decisions):
    assert st["iterations"] == so["iterations"] == g
    assert st["function_calls"] == so["function_calls"]
    if check_decisions:
        dec = pop.decisions(masks=True)
        for k in ("donors", "dim_idx", "rejects", "masks"):
            assert np.array_equal(dec[k], ao[k]), (g, k)
        flips = np.nonzero(dec["accepted"] != ao["accepted"])[0]
        assert flips.size == 0, (g, "accept flags differ at", flips[:10], dec["trial_scores"][flips[:10]],
                                 ao["trial_scores"][flips[:10]])
        if tol == 0.0:
            assert np.array_equal(bits(dec["trial_scores"]), bits(ao["trial_scores"])), (g, "trial_scores")
        else:
            assert rel_close(dec["trial_scores"], ao["trial_scores"], tol), (g, "trial_scores")
    rows, scores = pop.population(), pop.scores()
    if tol == 0.0:
        assert np.array_equal(bits(rows), bits(ao["rows"])), (g, "rows")
        assert np.array_equal(bits(scores), bits(ao["scores"])), (g, "scores")
        assert st["f_value"] == so["f_value"]
    else:
        assert rel_close(rows, ao["rows"], tol), (g, "rows")
        assert rel_close(scores, ao["scores"], tol), (g, "scores")
        assert rel_close(st["f_value"], so["f_value"], tol)
    assert st["best_index"] == so["best_index"], g
    assert np.array_equal(bits(pop.best()), bits(rows[so["best_index"]]))


CASES = [
    # dtype, objective, strategy, minimize, P, d, G, scale
    (B.F64, B.SPHERE, B.DE_RANDOM, True, 50, 2, 12, 5.0),
    (B.F64, B.SPHERE, B.DE_RANDOM, True, 1000, 16, 8, 10.24),
    (B.F64, B.ROSENBROCK, B.DE_BEST, True, 64, 8, 30, 4.096),
    (B.F64, B.ROSENBROCK, B.DE_BEST, True, 3000, 130, 5, 4.096),
    (B.F64, B.ROSENBROCK_EX, B.DE_RANDOM, False, 40, 5, 6, 3.0),
    (B.F64, B.RASTRIGIN, B.DE_RANDOM, True, 257, 33, 6, 10.24),
    (B.F64, B.RASTRIGIN, B.DE_RANDOM, True, 512, 1000, 3, 10.24),     # the north-star row shape, small population
    (B.F64, B.ACKLEY, B.DE_BEST, True, 128, 65, 6, 65.536),
    (B.F64, B.RASTRIGIN, B.DE_RANDOM, True, 4, 1, 10, 10.24),         # smallest legal population, d = 1
    (B.F64, B.SPHERE, B.DE_BEST, True, 5, 3, 10, 1.0),
    (B.F64, B.SPHERE, B.DE_RANDOM, True, 20000, 8, 4, 10.24),         # deep "donor r < i" DAG, many repair rounds
    (B.F64, B.SPHERE, B.DE_RANDOM, True, 333, 64, 5, 10.24),
    (B.F32, B.SPHERE, B.DE_RANDOM, True, 300, 17, 6, 10.24),
    (B.F32, B.ROSENBROCK, B.DE_BEST, True, 100, 12, 6, 4.096),
    (B.F32, B.RASTRIGIN, B.DE_RANDOM, True, 64, 9, 5, 10.24),
    (B.F32, B.SPHERE, B.DE_RANDOM, True, 2048, 64, 4, 10.24),
]


@pytest.mark.parametrize("dtype,obj,strategy,minimize,P,d,G,scale", CASES)
def test_de_matches_oracle_every_generation(ctx, oracle_lib, dtype, obj, strategy, minimize, P, d, G, scale):
    seed = 0xABCDEF0123456789 ^ (P * 1000003 + d)
    x0 = np.full(d, scale)
    tol = tolerance(dtype, obj)
    pop = gpu_de(ctx, dtype, obj, strategy, minimize, P, d, seed, x0)
    for g in range(G + 1):
        if g:
            pop.step(1)
        st = pop.sync()
        so, ao = oracle_de(oracle_lib, dtype, obj, strategy, minimize, P, d, g, seed, x0)
        compare_generation(pop, st, so, ao, tol, g, check_decisions=g > 0)
    pop.close()


@pytest.mark.parametrize("path", golden_files("de_"), ids=os.path.basename)
def test_de_matches_reference_fixture(ctx, path):
    """Against what the UNMODIFIED reference produced (tests/golden, written by tests/golden/make_golden.py)."""
    cfg, x0, z = load_de(path)
    tol = tolerance(cfg.dtype, cfg.objective)
    pop = gpu_de(ctx, cfg.dtype, cfg.objective, cfg.strategy, bool(cfg.minimize), cfg.pop_size, cfg.dim, cfg.seed, x0,
                 max_iter=cfg.max_iter)
    pop.step(cfg.max_iter)
    st = pop.sync()
    assert st["stopped"] and st["stop_reason"] == 1
    assert st["iterations"] == z["iterations"].item() and st["function_calls"] == z["function_calls"].item()
    assert st["best_index"] == z["best_index"].item()
    dec = pop.decisions(masks=True)
    for k in ("donors", "dim_idx", "rejects", "masks", "accepted"):
        assert np.array_equal(dec[k], z[k]), k
    if tol == 0.0:
        assert np.array_equal(bits(pop.population()), bits(z["rows"]))
        assert np.array_equal(bits(pop.best()), bits(z["x_best"]))
        assert st["f_value"] == z["f_value"].item()
    else:
        assert rel_close(pop.population(), z["rows"], tol) and rel_close(st["f_value"], z["f_value"].item(), tol)
    pop.close()


def test_de_batched_steps_equal_single_steps(ctx, oracle_lib):
    P, d, G, seed = 777, 24, 9, 5
    x0 = np.full(d, 4.096)
    a = gpu_de(ctx, B.F64, B.ROSENBROCK, B.DE_RANDOM, True, P, d, seed, x0, masks=False)
    a.step(G)
    so, ao = oracle_de(oracle_lib, B.F64, B.ROSENBROCK, B.DE_RANDOM, True, P, d, G, seed, x0, masks=False)
    st = a.sync()
    assert st["iterations"] == G and np.array_equal(bits(a.population()), bits(ao["rows"]))
    a.close()


def test_de_stop_rules_on_device(ctx, oracle_lib):
    """Default stop rules (eps = 10e-4, best_val_no_change = 50) fire on the device in the reference's iteration."""
    for obj, strategy, P, d, scale in ((B.ROSENBROCK, B.DE_BEST, 64, 8, 4.096), (B.SPHERE, B.DE_RANDOM, 50, 2, 1.0)):
        cfg = B.de_cfg(objective=obj, strategy=strategy, pop_size=P, dim=d, seed=99)
        so, ao = B.de_run(oracle_lib, cfg, np.full(d, scale))
        x = np.full(d, scale)
        ncfg = nb.de_cfg(objective=obj, strategy=strategy, pop_size=P, dim=d, seed=99)
        pop = nb.DEPopulation(ctx, ncfg, x)
        pop.step(so["iterations"] + 20)      # generations past the stop are no-ops
        st = pop.sync()
        assert st["stopped"] and st["stop_reason"] == so["stop_reason"]
        assert st["iterations"] == so["iterations"] and st["function_calls"] == so["function_calls"]
        assert st["f_value"] == so["f_value"]
        assert np.array_equal(bits(pop.best()), bits(ao["x_best"]))
        pop.close()


def test_de_solve_one_shot_matches_oracle(ctx, oracle_lib):
    import ctypes as C
    from nlsolver_b200 import _lib as L
    d = 6
    cfg = nb.de_cfg(objective=nb.ROSENBROCK, strategy=nb.DE_BEST, pop_size=80, dim=d, seed=1234)
    x0, out, st = np.full(d, 4.096), np.zeros(d), L.Status()
    L.check(L.lib().nls_de_solve(ctx.handle, C.byref(cfg), x0.ctypes.data, out.ctypes.data, C.byref(st)))
    so, ao = B.de_run(oracle_lib, B.de_cfg(objective=B.ROSENBROCK, strategy=B.DE_BEST, pop_size=80, dim=d, seed=1234), x0)
    assert (st.iterations, st.function_calls, st.f_value) == (so["iterations"], so["function_calls"], so["f_value"])
    assert np.array_equal(bits(out), bits(ao["x_best"]))


def test_de_rejects_bad_arguments(ctx):
    with pytest.raises(nb.NlsError):
        nb.DEPopulation(ctx, nb.de_cfg(pop_size=3, dim=2), np.ones(2))      # the reference would loop forever
    with pytest.raises(nb.NlsError):
        nb.DEPopulation(ctx, nb.de_cfg(pop_size=10, dim=2, objective=17), np.ones(2))


def test_de_large_population_properties(ctx):
    """Size-independent properties at a population the oracle would take minutes for: greedy selection never
    worsens a score, the reported best is the population minimum with the lowest index, counters add up."""
    P, d = 1 << 17, 64
    pop = gpu_de(ctx, B.F64, B.SPHERE, B.DE_RANDOM, True, P, d, 7, np.full(d, 10.24), masks=False)
    prev = pop.scores()
    for g in range(1, 4):
        pop.step(1)
        st = pop.sync()
        cur = pop.scores()
        assert np.all(cur <= prev) and st["iterations"] == g and st["function_calls"] == P * (g + 1)
        dec = pop.decisions()
        assert np.array_equal(dec["accepted"].astype(bool), cur < prev)
        assert np.array_equal(np.where(dec["accepted"] == 1, dec["trial_scores"], prev), cur)
        don = dec["donors"].astype(np.int64)
        idx = np.arange(P)
        assert np.all(don < P) and np.all(don != idx[:, None])
        assert np.all(don[:, 0] != don[:, 1]) and np.all(don[:, 0] != don[:, 2]) and np.all(don[:, 1] != don[:, 2])
        assert st["best_index"] == int(np.argmin(cur)) and st["f_value"] == cur.min()
        rows = pop.population()
        assert np.allclose((rows * rows).sum(1), cur, rtol=1e-12)
        prev = cur
    pop.close()


def test_std_err_stop_rule_is_exact_at_the_threshold(ctx, oracle_lib):
    """The stop statistic is a pairwise reduction on the device and sequential sums in the reference; when it lands
    within 1e-9 of eps the device recomputes it sequentially, so the stop fires in the same generation even when eps
    sits one ulp above / exactly at the reference's value."""
    P, d, seed = 200, 6, 77
    x0 = np.full(d, 2.0)
    # std_err of the scores the reference sees at the top of iteration 7
    so, ao = oracle_de(oracle_lib, B.F64, B.SPHERE, B.DE_RANDOM, True, P, d, 7, seed, x0, masks=False)
    se = oracle_lib.oracle_std_err_f64(ao["scores"].ctypes.data, P)
    for eps, stops_at in ((np.nextafter(se, np.inf), 7), (se, None)):
        cfg = B.de_cfg(objective=B.SPHERE, pop_size=P, dim=d, eps=float(eps), max_iter=40, best_val_no_change=1 << 40, seed=seed)
        want, _ = B.de_run(oracle_lib, cfg, x0)
        if stops_at is not None:
            assert want["iterations"] == stops_at and want["stop_reason"] == 3
        pop = nb.DEPopulation(ctx, nb.de_cfg(objective=nb.SPHERE, pop_size=P, dim=d, eps=float(eps), max_iter=40,
                                             best_val_no_change=1 << 40, seed=seed), x0)
        pop.step(40)
        st = pop.sync()
        assert (st["iterations"], st["stop_reason"]) == (want["iterations"], want["stop_reason"]), (eps, st, want)
        assert st["f_value"] == want["f_value"]
        pop.close()


BULK_CASES = [
    # dtype, objective, strategy, P, d, G, scale, F   (rows of 1 .. 63 chunks of 1 KB, ragged last chunks, short rows)
    (B.F64, B.RASTRIGIN, B.DE_RANDOM, 512, 1000, 3, 10.24, 0.8),
    (B.F64, B.ROSENBROCK, B.DE_BEST, 300, 777, 5, 4.096, 0.8),
    (B.F64, B.SPHERE, B.DE_RANDOM, 400, 1000, 6, 10.24, 0.2),      # trials are accepted: second sweep + repair
    (B.F64, B.ROSENBROCK, B.DE_BEST, 96, 4096, 3, 4.096, 0.8),     # the config-4 row
    (B.F64, B.ACKLEY, B.DE_BEST, 100, 130, 4, 32.768, 0.8),
    (B.F32, B.SPHERE, B.DE_RANDOM, 200, 3, 6, 10.24, 0.8),         # one 16-byte piece per row
    (B.F32, B.ROSENBROCK, B.DE_BEST, 150, 513, 4, 4.096, 0.5),
]


@pytest.mark.parametrize("dtype,obj,strategy,P,d,G,scale,F", BULK_CASES)
def test_de_tma_staged_generation_matches_oracle(ctx, oracle_lib, monkeypatch, dtype, obj, strategy, P, d, G, scale, F):
    """The generation pass with the rows staged through shared memory by bulk copies (de_generation_bulk_kernel;
    NLS_DE_BULK=3 forces it for every shape) gives the decisions / rows of the LDG pass: same oracle, same tolerances."""
    monkeypatch.setenv("NLS_DE_BULK", "3")
    seed = 0x1234ABCD ^ (P * 7919 + d)
    x0 = np.full(d, scale)
    tol = tolerance(dtype, obj)
    pop = gpu_de(ctx, dtype, obj, strategy, True, P, d, seed, x0, f=F)
    for g in range(1, G + 1):
        pop.step(1)
        st = pop.sync()
        so, ao = oracle_de(oracle_lib, dtype, obj, strategy, True, P, d, g, seed, x0, f=F)
        compare_generation(pop, st, so, ao, tol, g, check_decisions=True)
    pop.close()


@pytest.mark.parametrize("dtype,obj,strategy,P,d,scale,F", [
    (B.F64, B.SPHERE, B.DE_RANDOM, 1024, 64, 10.24, 0.5),        # 16 CTAs in the cluster
    (B.F64, B.ROSENBROCK, B.DE_BEST, 50, 2, 4.096, 0.8),         # the reference's default shape: one CTA
    (B.F32, B.RASTRIGIN, B.DE_RANDOM, 300, 40, 10.24, 0.3),
    (B.F64, B.ACKLEY, B.DE_BEST, 16, 1000, 32.768, 0.8),         # few agents, long rows
])
def test_de_one_launch_path_equals_separate_kernels(ctx, monkeypatch, dtype, obj, strategy, P, d, scale, F):
    """Populations of up to 2^16 elements run all the generations of a step in ONE launch on one thread-block cluster
    (de_persistent_kernel); with NLS_DE_ONE_LAUNCH=0 the same step goes through the separate kernels (CUDA graph).  Both
    must give the same bits: rows, scores, decisions of the last generation, counters — over several step sizes."""
    seed, x0 = 77 + P, np.full(d, scale)
    out = []
    for env in ("1", "0"):
        monkeypatch.setenv("NLS_DE_ONE_LAUNCH", env)
        pop = gpu_de(ctx, dtype, obj, strategy, True, P, d, seed, x0, f=F, masks=False)
        for n in (1, 9, 30):
            pop.step(n)
        st = pop.sync()
        out.append((st, pop.population(), pop.scores(), pop.decisions()))
        pop.close()
    (sa, ra, ca, da), (sb, rb, cb, db) = out
    assert sa["iterations"] == sb["iterations"] == 40
    for k in ("f_value", "function_calls", "best_index", "val_no_change", "std_err", "accepted_total"):
        assert sa[k] == sb[k], k
    assert np.array_equal(bits(ra), bits(rb)) and np.array_equal(bits(ca), bits(cb))
    for k in ("donors", "dim_idx", "rejects", "accepted", "trial_scores"):
        assert np.array_equal(bits(da[k]), bits(db[k])), k


@pytest.mark.parametrize("dtype,obj,strategy,P,d,scale,F", [
    (B.F64, B.SPHERE, B.DE_RANDOM, 5000, 64, 10.24, 0.4),
    (B.F32, B.ROSENBROCK, B.DE_BEST, 900, 24, 4.096, 0.8),
    (B.F64, B.RASTRIGIN, B.DE_RANDOM, 2000, 300, 10.24, 0.2),     # long rows: repair with four steps in flight
])
def test_de_commit_inside_the_repair_launch_equals_three_kernels(ctx, monkeypatch, dtype, obj, strategy, P, d, scale, F):
    """Launch-bound populations run K3 behind the repair in the same cooperative launch (two launches per generation);
    with kernel timing enabled the generation is the three separate kernels.  Same bits, same counters, same std_err."""
    monkeypatch.setenv("NLS_DE_ONE_LAUNCH", "0")
    seed, x0 = 1234 + P, np.full(d, scale)
    out = []
    for timed in (False, True):
        pop = gpu_de(ctx, dtype, obj, strategy, True, P, d, seed, x0, f=F, masks=False)
        pop.enable_kernel_timing(timed)
        for n in (1, 8, 11):
            pop.step(n)
        st = pop.sync()
        out.append((st, pop.population(), pop.scores(), pop.decisions()))
        pop.close()
    (sa, ra, ca, da), (sb, rb, cb, db) = out
    assert sa["iterations"] == sb["iterations"] == 20 and sa["accepted_total"] > 0
    for k in ("f_value", "function_calls", "best_index", "val_no_change", "std_err", "accepted_total"):
        assert sa[k] == sb[k], k
    assert np.array_equal(bits(ra), bits(rb)) and np.array_equal(bits(ca), bits(cb))
    for k in ("donors", "dim_idx", "rejects", "accepted", "trial_scores"):
        assert np.array_equal(bits(da[k]), bits(db[k])), k


TINY_CASES = [
    # dtype, objective, strategy, minimize, P, d, scale, max_iter
    (B.F64, B.ROSENBROCK_EX, B.DE_RANDOM, True, 50, 2, 5.0, 1000),       # BASELINE configs[0]: the README snippet's shape
    (B.F64, B.ROSENBROCK_EX, B.DE_BEST, True, 50, 2, 2.0, 1000),         # example.cpp:186-188
    (B.F64, B.RASTRIGIN, B.DE_RANDOM, True, 1024, 8, 10.24, 60),         # the largest shape of the one-block solver
    (B.F64, B.ACKLEY, B.DE_BEST, False, 333, 5, 32.768, 40),             # maximize
    (B.F64, B.SHEKEL, B.DE_RANDOM, True, 50, 4, -0.5, 300),
    (B.F64, B.BEALE, B.DE_BEST, True, 50, 2, -0.5, 300),
    (B.F64, B.STYBLINSKI_TANG, B.DE_RANDOM, True, 64, 7, 5.0, 80),
    (B.F64, B.SPHERE, B.DE_RANDOM, True, 4, 1, 10.24, 30),               # smallest legal population, d = 1
    (B.F32, B.ROSENBROCK, B.DE_BEST, True, 100, 8, 4.096, 50),
    (B.F32, B.SPHERE, B.DE_RANDOM, True, 50, 2, 5.0, 1000),
]


@pytest.mark.parametrize("dtype,obj,strategy,minimize,P,d,scale,max_iter", TINY_CASES)
def test_de_one_block_solver_equals_general_path_and_oracle(ctx, oracle_lib, monkeypatch, dtype, obj, strategy, minimize, P,
                                                            d, scale, max_iter):
    """nls_de_solve for pop_size <= 1024 and dim <= 8 runs the whole solve in ONE launch of ONE block with the population
    in shared memory (de_tiny_solve_kernel).  It must return what the general kernels return (NLS_DE_TINY=0) bit for bit
    — default stop rules active — and agree with the oracle like them."""
    import ctypes as C
    from nlsolver_b200 import _lib as L
    dt = np.float64 if dtype == B.F64 else np.float32
    x0 = np.full(d, scale, dt)
    cfg = nb.de_cfg(dtype=dtype, objective=obj, strategy=strategy, minimize=minimize, pop_size=P, dim=d, max_iter=max_iter,
                    seed=4242 + P)
    res = []
    for env in ("1", "0"):
        monkeypatch.setenv("NLS_DE_TINY", env)
        out, st = np.zeros(d, dt), L.Status()
        L.check(L.lib().nls_de_solve(ctx.handle, C.byref(cfg), x0.ctypes.data, out.ctypes.data, C.byref(st)))
        res.append((st.as_dict(), out))
    (sa, xa), (sb, xb) = res
    for k in ("f_value", "iterations", "function_calls", "best_index", "val_no_change", "stop_reason", "accepted_total"):
        assert sa[k] == sb[k], (k, sa[k], sb[k])
    assert sa["stopped"] == 1 and np.array_equal(bits(xa), bits(xb))
    so, ao = B.de_run(oracle_lib, B.de_cfg(dtype=dtype, objective=obj, strategy=strategy, minimize=minimize, pop_size=P, dim=d,
                                           max_iter=max_iter, seed=4242 + P), x0)
    tol = tolerance(dtype, obj)
    if tol == 0.0 or obj in (B.SHEKEL, B.BEALE):
        assert (sa["iterations"], sa["function_calls"], sa["stop_reason"]) == (so["iterations"], so["function_calls"],
                                                                             so["stop_reason"])
    if tol == 0.0:
        assert sa["f_value"] == so["f_value"] and np.array_equal(bits(xa), bits(ao["x_best"]))
