"""Pins the oracle (oracle/popsolve_oracle.cpp) to the reference.

(1) known-answer vectors derived from the reference (BASELINE.md §2 / SURVEY.md §6, §8a R1-R2);
(2) bit-for-bit agreement with the UNMODIFIED reference templates (oracle/_ref, built from /root/reference by
    oracle/Makefile) on the same draw tape: final x, f_value, iterations, function calls, draws consumed, the whole
    population, and the last generation's donors / dim / rejects / masks / accept flags.
CPU only (-m "not gpu")."""
import numpy as np
import pytest

from oracle import binding as B

OBJS = [B.SPHERE, B.ROSENBROCK, B.RASTRIGIN, B.ACKLEY, B.ROSENBROCK_EX]


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32 if a.dtype == np.float32 else a.dtype)


def same_bits(a, b):
    return np.array_equal(bits(a), bits(b))


# ---------------------------------------------------------------- known answers -----------------------------------
def test_xorshift_default_state_and_first_draws(oracle_lib):
    st, raw = (B.u64 * 2)(), (B.u64 * 4)()
    oracle_lib.oracle_xorshift_default(st, raw, 4)
    assert (st[0], st[1]) == (0x7c26ca28fb68bc1b, 0x7c26ca28)          # nlsolver.h:1345-1349
    d = [np.float64(oracle_lib.oracle_unit_f64(x)).view(np.uint64) for x in raw]
    assert d == [0x3fda16d91834b672, 0x3fea70621141f57f, 0x3fe366ccce7f386d, 0x3fdfa0ee9f259b16]
    f = [np.float32(oracle_lib.oracle_unit_f32(x)).view(np.uint32) for x in raw]
    assert f == [0x3ed0b6c9, 0x3f538311, 0x3f1b3666, 0x3efd0775]


def test_readme_de_snippet_known_answer(oracle_lib):
    # README.md:94-110 with example.cpp:41-48's Rosenbrock: DE-random, defaults, xorshift<double>, x0 = {5, 7}
    cfg = B.de_cfg(objective=B.ROSENBROCK_EX, rng_mode=B.RNG_XORSHIFT)
    st, a = B.de_run(oracle_lib, cfg, [5, 7])
    assert (st["function_calls"], st["iterations"]) == (2700, 53)
    assert np.float64(st["f_value"]).view(np.uint64) == 0x3ee000ea0efc0071
    assert a["x_best"].tolist() == [0.99754919858453095, 0.99523186613831982]


def test_example_cpp_de_best_known_answer(oracle_lib):
    # example.cpp:184-188: DE-best, xorshift<double>, x0 = {2, 7}
    cfg = B.de_cfg(objective=B.ROSENBROCK_EX, rng_mode=B.RNG_XORSHIFT, strategy=B.DE_BEST)
    st, a = B.de_run(oracle_lib, cfg, [2, 7])
    assert (st["function_calls"], st["iterations"]) == (1950, 38)
    assert "%g" % st["f_value"] == "2.07418e-05"
    assert ["%g" % v for v in a["x_best"]] == ["1.00245", "1.00529"]


# ---------------------------------------------------------------- primitives vs reference --------------------------
def test_first_draws_match_reference_generator(oracle_lib, ref_lib):
    for dtype, unit in ((B.F64, oracle_lib.oracle_unit_f64), (B.F32, oracle_lib.oracle_unit_f32)):
        n = 1000
        st, raw = (B.u64 * 2)(), (B.u64 * n)()
        oracle_lib.oracle_xorshift_default(st, raw, n)
        ref = np.zeros(n)
        ref_lib.ref_xorshift_draws(dtype, n, ref.ctypes.data)
        assert [float(unit(x)) for x in raw] == ref.tolist()


def test_objectives_reduce_to_reference_2d_forms(oracle_lib, ref_lib):
    rng = np.random.default_rng(7)
    for obj in (B.SPHERE, B.ROSENBROCK, B.RASTRIGIN, B.ACKLEY):
        for x in rng.uniform(-5, 5, size=(500, 2)):
            ours = B.objective(B.F64, obj, x)
            ref = ref_lib.ref_objective_2d(obj, x[0], x[1])
            # Rosenbrock: the reference calls pow(v, 2.0); glibc's pow is not guaranteed correctly rounded
            assert ours == ref or (obj == B.ROSENBROCK and abs(ours - ref) <= 2e-16 * abs(ref)), (obj, x)


def test_every_test_driver_objective_matches_the_reference_functor(oracle_lib, ref_lib):
    """All 15 problems the reference's test driver enables (test_functions.h:485-524), evaluated by the reference's own
    functors.  Exact where only + - * / sqrt sin are involved in the same order; ThreeHumpCamel and StyblinskiTang call
    pow(x, 4) / pow(x, 6) in the reference, restated as products (<= 1e-13 relative)."""
    rng = np.random.default_rng(11)
    for obj in range(B.BEALE, B.LEVI_N13 + 1):
        d = 4 if obj == B.SHEKEL else 2
        loose = obj in (B.THREE_HUMP_CAMEL, B.STYBLINSKI_TANG)
        for x in rng.uniform(-4.5, 4.5, size=(400, d)):
            ours = B.objective(B.F64, obj, x)
            ref = ref_lib.ref_objective_nd(obj, x.ctypes.data, d)
            assert ours == ref or (loose and abs(ours - ref) <= 1e-13 * max(abs(ref), 1.0)), (obj, x, ours, ref)


def test_std_err_matches_reference(oracle_lib, ref_lib):
    x = np.random.default_rng(3).normal(size=1001)
    assert oracle_lib.oracle_std_err_f64(x.ctypes.data, x.size) == ref_lib.ref_std_err_f64(x.ctypes.data, x.size)


# ---------------------------------------------------------------- DE on the tape -----------------------------------
DE_CASES = [
    # dtype, objective, strategy, minimize, P, d, G, x0 scale
    (B.F64, B.SPHERE, B.DE_RANDOM, True, 50, 2, 30, 5.0),
    (B.F64, B.SPHERE, B.DE_RANDOM, True, 1000, 16, 12, 10.24),
    (B.F64, B.ROSENBROCK, B.DE_BEST, True, 64, 8, 40, 4.096),
    (B.F64, B.RASTRIGIN, B.DE_RANDOM, True, 257, 33, 9, 10.24),
    (B.F64, B.ACKLEY, B.DE_BEST, True, 128, 65, 7, 65.536),
    (B.F64, B.ROSENBROCK_EX, B.DE_RANDOM, False, 40, 5, 6, 3.0),
    (B.F64, B.RASTRIGIN, B.DE_RANDOM, True, 4, 1, 25, 10.24),      # smallest legal population, d = 1
    (B.F64, B.SPHERE, B.DE_BEST, True, 5, 3, 25, 1.0),
    (B.F32, B.SPHERE, B.DE_RANDOM, True, 300, 17, 10, 10.24),
    (B.F32, B.ROSENBROCK, B.DE_BEST, True, 100, 12, 10, 4.096),
    (B.F32, B.RASTRIGIN, B.DE_RANDOM, True, 64, 9, 8, 10.24),
]


@pytest.mark.parametrize("dtype,obj,strategy,minimize,P,d,G,scale", DE_CASES)
def test_de_restatement_equals_reference_on_tape(oracle_lib, ref_lib, dtype, obj, strategy, minimize, P, d, G, scale):
    x0 = np.full(d, scale)
    for g in sorted({0, 1, 2, G}):
        cfg = B.de_cfg(dtype=dtype, objective=obj, strategy=strategy, minimize=minimize, pop_size=P, dim=d, eps=0.0,
                       max_iter=g, best_val_no_change=1 << 40, seed=0x1234_5678_9ABC_DEF0 + P)
        so, ao = B.de_run(oracle_lib, cfg, x0, masks=True)
        sr, ar = B.de_run(ref_lib, cfg, x0, masks=True)
        assert ref_lib.ref_last_inconsistencies() == 0
        for k in ("f_value", "iterations", "function_calls", "draws_consumed", "best_index"):
            assert so[k] == sr[k], (g, k, so[k], sr[k])
        assert so["iterations"] == g and so["function_calls"] == P * (g + 1)
        for k in ("x_best", "rows", "scores"):
            assert same_bits(ao[k], ar[k]), (g, k)
        if g > 0:
            for k in ("donors", "dim_idx", "rejects", "accepted", "masks", "trial_scores"):
                assert same_bits(ao[k], ar[k]), (g, k)


def test_de_stop_rules_match_reference(oracle_lib, ref_lib):
    # default stop rules (eps = 10e-4, best_val_no_change = 50): both std_err and val_no_change paths
    for obj, strategy, P, d, scale in ((B.ROSENBROCK, B.DE_BEST, 64, 8, 4.096), (B.SPHERE, B.DE_RANDOM, 50, 2, 1.0),
                                       (B.RASTRIGIN, B.DE_RANDOM, 20, 2, 0.5)):
        cfg = B.de_cfg(objective=obj, strategy=strategy, pop_size=P, dim=d, seed=99)
        so, ao = B.de_run(oracle_lib, cfg, np.full(d, scale))
        sr, ar = B.de_run(ref_lib, cfg, np.full(d, scale))
        assert so["stop_reason"] in (1, 2, 3)
        for k in ("f_value", "iterations", "function_calls", "draws_consumed"):
            assert so[k] == sr[k], k
        assert same_bits(ao["x_best"], ar["x_best"]) and same_bits(ao["rows"], ar["rows"])


def test_de_sequential_xorshift_matches_reference(oracle_lib, ref_lib):
    for dtype in (B.F64, B.F32):
        cfg = B.de_cfg(dtype=dtype, objective=B.RASTRIGIN, pop_size=200, dim=10, eps=0.0, max_iter=20,
                       best_val_no_change=1 << 40, rng_mode=B.RNG_XORSHIFT)
        so, ao = B.de_run(oracle_lib, cfg, np.full(10, 10.24))
        sr, ar = B.de_run(ref_lib, cfg, np.full(10, 10.24))
        assert so["f_value"] == sr["f_value"] and same_bits(ao["rows"], ar["rows"])


# ---------------------------------------------------------------- PSO on the tape ----------------------------------
PSO_CASES = [
    # dtype, objective, type, minimize, constrained, P, d, G, bound
    (B.F64, B.SPHERE, B.PSO_VANILLA, True, False, 8, 8, 40, 10.24),
    (B.F64, B.ROSENBROCK, B.PSO_VANILLA, True, True, 6, 8, 40, 4.096),
    (B.F64, B.RASTRIGIN, B.PSO_VANILLA, True, False, 30, 33, 15, 5.12),
    (B.F64, B.SPHERE, B.PSO_ACCELERATED, True, False, 40, 8, 30, 10.24),
    (B.F64, B.ACKLEY, B.PSO_ACCELERATED, True, False, 1000, 32, 12, 32.768),
    (B.F64, B.RASTRIGIN, B.PSO_ACCELERATED, True, True, 40, 8, 50, 5.12),
    (B.F64, B.ROSENBROCK_EX, B.PSO_ACCELERATED, False, True, 33, 5, 10, 2.0),
    (B.F32, B.SPHERE, B.PSO_ACCELERATED, True, False, 100, 16, 10, 10.24),
    (B.F32, B.SPHERE, B.PSO_VANILLA, True, True, 12, 16, 10, 10.24),
]


@pytest.mark.parametrize("dtype,obj,ptype,minimize,constrained,P,d,G,bound", PSO_CASES)
def test_pso_restatement_equals_reference_on_tape(oracle_lib, ref_lib, dtype, obj, ptype, minimize, constrained, P, d,
                                                  G, bound):
    up = np.full(d, bound)
    lo = -up
    for g in sorted({0, 1, 2, G}):
        cfg = B.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, minimize=minimize, n_particles=P, dim=d, eps=0.0,
                        max_iter=g, best_val_no_change=1 << 40, constrained=constrained, seed=42 + d)
        so, ao = B.pso_run(oracle_lib, cfg, lo, up)
        sr, ar = B.pso_run(ref_lib, cfg, lo, up)
        for k in ("f_value", "iterations", "function_calls", "draws_consumed", "best_valid"):
            assert so[k] == sr[k], (g, k, so[k], sr[k])
        for k in ("x_best", "positions", "pbest_values", "last_values"):
            assert same_bits(ao[k], ar[k]), (g, k)


def test_pso_stop_rules_match_reference(oracle_lib, ref_lib):
    for ptype, P, d in ((B.PSO_ACCELERATED, 40, 8), (B.PSO_VANILLA, 8, 8)):
        cfg = B.pso_cfg(objective=B.SPHERE, pso_type=ptype, n_particles=P, dim=d, seed=5)
        up = np.full(d, 3.0)
        so, ao = B.pso_run(oracle_lib, cfg, -up, up)
        sr, ar = B.pso_run(ref_lib, cfg, -up, up)
        for k in ("f_value", "iterations", "function_calls", "draws_consumed"):
            assert so[k] == sr[k], k
        assert same_bits(ao["x_best"], ar["x_best"])


# ---------------------------------------------------------------- SANN chains ---------------
def test_sann_default_xorshift_known_answer(oracle_lib):
    # nlsolver::SANN<Rosenbrock (example.cpp:41-48), xorshift<double>> with its defaults from {5, 5}: values produced by
    # the unmodified reference (oracle/_ref) and frozen here so that the check also runs where the reference cannot
    cfg = B.sann_cfg(objective=B.ROSENBROCK_EX, n_chains=1, dim=2, rng_mode=B.RNG_XORSHIFT)
    st, a = B.sann_run(oracle_lib, cfg, np.array([5.0, 5.0]))
    assert (st["iterations"], st["function_calls"]) == (5000, 45001)
    assert float(st["f_value"]).hex() == "0x1.18098474b13bep-12"
    assert [float(v).hex() for v in a["x_best"][0]] == ["0x1.f7ab827808a3fp-1", "0x1.ef8dddf306125p-1"]


SANN_CASES = [
    # dtype, objective, minimize, n_chains, d, max_iter, temperature_iter, temperature_max
    (B.F64, B.SPHERE, True, 6, 7, 200, 10, 10.0),
    (B.F64, B.ROSENBROCK, True, 5, 2, 300, 10, 10.0),
    (B.F64, B.RASTRIGIN, True, 4, 67, 60, 5, 3.0),
    (B.F64, B.ACKLEY, False, 6, 5, 150, 10, 10.0),
    (B.F64, B.STYBLINSKI_TANG, True, 3, 9, 100, 2, 1.0),
    (B.F64, B.BEALE, True, 4, 2, 200, 10, 10.0),
    (B.F32, B.SPHERE, True, 6, 7, 200, 10, 10.0),
    (B.F32, B.RASTRIGIN, False, 4, 16, 100, 10, 10.0),
    (B.F64, B.SHEKEL, True, 3, 4, 150, 10, 10.0),
    (B.F64, B.LEVI_N13, False, 3, 2, 150, 10, 5.0),
    (B.F32, B.ROSENBROCK_EX, True, 5, 2, 300, 10, 10.0),
    (B.F64, B.SPHERE, True, 2, 3, 50, 1, 10.0),     # temperature_iter = 1: no candidate is ever evaluated
    (B.F64, B.SPHERE, True, 2, 3, 0, 10, 10.0),     # max_iter = 0
]


@pytest.mark.parametrize("dtype,obj,minimize,n,d,it,ti,tmax", SANN_CASES)
def test_sann_restatement_equals_reference_on_tape(oracle_lib, ref_lib, dtype, obj, minimize, n, d, it, ti, tmax):
    rng = np.random.default_rng(100 + d)
    for x0 in (np.linspace(-2.0, 3.0, d), rng.uniform(-3, 3, size=(n, d))):   # shared start / one start per chain
        cfg = dict(dtype=dtype, objective=obj, minimize=minimize, n_chains=n, dim=d, max_iter=it, temperature_iter=ti,
                   temperature_max=tmax, seed=77 + d, chain_offset=5)
        so, ao = B.sann_run(oracle_lib, B.sann_cfg(**cfg), x0)
        sr, ar = B.sann_run(ref_lib, B.sann_cfg(**cfg), x0)
        for k in ("f_value", "iterations", "function_calls", "draws_consumed", "best_index"):
            assert so[k] == sr[k], (k, so[k], sr[k])
        for k in ("x_best", "f_best", "draws", "iterations", "function_calls"):
            assert same_bits(ao[k], ar[k]), k
        assert (ao["function_calls"] == 1 + it * max(ti - 1, 0)).all()


def test_sann_sequential_xorshift_matches_reference(oracle_lib, ref_lib):
    # chains run one after the other on ONE shared generator, like repeated minimize() calls on the same RNG object
    for dtype in (B.F64, B.F32):
        cfg = dict(dtype=dtype, objective=B.ROSENBROCK_EX, n_chains=3, dim=2, max_iter=800, rng_mode=B.RNG_XORSHIFT)
        so, ao = B.sann_run(oracle_lib, B.sann_cfg(**cfg), np.array([2.0, 7.0]))
        sr, ar = B.sann_run(ref_lib, B.sann_cfg(**cfg), np.array([2.0, 7.0]))
        assert so["f_value"] == sr["f_value"] and so["function_calls"] == sr["function_calls"]
        assert same_bits(ao["x_best"], ar["x_best"]) and same_bits(ao["f_best"], ar["f_best"])


def test_sann_restatement_cut_points_compose(oracle_lib):
    # max_steps cuts a chain without disturbing it: the state after k candidates is a prefix of the full run's history
    cfg = dict(objective=B.RASTRIGIN, n_chains=3, dim=9, max_iter=40, temperature_iter=10, seed=9)
    full, af = B.sann_run(oracle_lib, B.sann_cfg(**cfg), np.full(9, 2.0))
    cut, ac = B.sann_run(oracle_lib, B.sann_cfg(max_steps=360, **cfg), np.full(9, 2.0))
    assert same_bits(af["x_best"], ac["x_best"]) and same_bits(af["p_cur"], ac["p_cur"])
    part, ap = B.sann_run(oracle_lib, B.sann_cfg(max_steps=100, **cfg), np.full(9, 2.0))
    assert (ap["function_calls"] == 101).all() and (ap["n_accepted"] <= af["n_accepted"]).all()
    assert (ap["f_best"] >= af["f_best"]).all()


@pytest.mark.parametrize("dtype,obj,d,scale", [(B.F64, B.SPHERE, 2, 2.0), (B.F64, B.ROSENBROCK_EX, 2, 2.0),
                                               (B.F64, B.RASTRIGIN, 4, 3.0), (B.F64, B.ACKLEY, 6, 5.0),
                                               (B.F64, B.ROSENBROCK, 8, 1.5), (B.F64, B.STYBLINSKI_TANG, 10, 2.0),
                                               (B.F32, B.SPHERE, 2, 2.0), (B.F32, B.ROSENBROCK, 4, 1.5),
                                               (B.F32, B.RASTRIGIN, 8, 3.0)])
def test_nmpso_restatement_equals_reference(oracle_lib, ref_lib, dtype, obj, d, scale):
    """NelderMeadPSO::solve restated (oracle/popsolve_oracle.cpp: nmpso_solve) against the unmodified reference, on the
    draw tape and on the reference's own sequential xorshift, minimise and maximise: best points, values, iteration and
    call counters and the number of draws consumed, bit for bit."""
    x0 = np.random.default_rng(d).uniform(-scale, scale, size=(7, d))
    for mode in (B.RNG_TAPE, B.RNG_XORSHIFT):
        for minimize in (True, False):
            kw = dict(dtype=dtype, objective=obj, minimize=minimize, n_solvers=7, dim=d, rng_mode=mode, seed=5,
                      max_iter=1000 if minimize else 25)
            so, ao = B.nmpso_run(oracle_lib, B.nmpso_cfg(**kw), x0)
            sr, ar = B.nmpso_run(ref_lib, B.nmpso_cfg(**kw), x0)
            assert sr is not None, "the harness refused a shape this test believes the reference survives"
            for k in ("x_best", "f_best", "iterations", "function_calls"):
                assert np.array_equal(ao[k].view(np.uint8), ar[k].view(np.uint8)), (mode, minimize, k)
            if mode == B.RNG_TAPE:
                assert np.array_equal(ao["draws"], ar["draws"])
            assert so["f_value"] == sr["f_value"] and so["best_index"] == sr["best_index"]


def test_nmpso_harness_refuses_shapes_that_corrupt_the_reference_heap(ref_lib):
    """The reference stores one element past the end of a row (nlsolver.h:3697-3700); for odd n in fp64 that overwrites
    a heap chunk header.  The harness must refuse instead of aborting the test process."""
    st, _ = B.nmpso_run(ref_lib, B.nmpso_cfg(objective=B.SPHERE, n_solvers=1, dim=3), np.ones(3))
    assert st is None
