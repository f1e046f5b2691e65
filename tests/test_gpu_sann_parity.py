"""SANN chain batches on the B200 (through the C ABI) against the oracle on the same draw tape.

Accept / improve decisions are compared exactly (per-chain counters, and the chains' points, which any flipped decision
would move by a whole proposal); values within 1e-12 relative in fp64 (log / sqrt / cos / exp differ in the last ulp
between CUDA libdevice and glibc)."""
import os

import numpy as np
import pytest

import nlsolver_b200 as nb
from oracle import binding as B
from tests.golden_util import golden_files, load_sann
from tests.gpu_util import bits, rel_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context(0)
    yield c
    c.close()


def tol_of(dtype, obj=None):
    if dtype == B.F64:
        return 1e-12
    # fp32 proposals: rnorm in double, rounded once, like the reference's rnorm<float> (2 ulp of float with a + - *
    # objective); objectives that call cosf / expf keep the looser fp32 tolerance
    return 2.4e-7 if obj in (B.SPHERE, B.ROSENBROCK, B.ROSENBROCK_EX) else 2e-5


def start_points(n, d, shared, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-3, 3, size=d) if shared else rng.uniform(-3, 3, size=(n, d))


def gpu_chains(ctx, dtype, obj, minimize, n, d, it, ti, tmax, seed, x0, offset=0):
    cfg = nb.sann_cfg(dtype=dtype, objective=obj, minimize=minimize, n_chains=n, dim=d, max_iter=it,
                      temperature_iter=ti, temperature_max=tmax, seed=seed, chain_offset=offset)
    return nb.SANNChains(ctx, cfg, x0)


def oracle_chains(lib, dtype, obj, minimize, n, d, it, ti, tmax, seed, x0, offset=0, max_steps=0):
    cfg = B.sann_cfg(dtype=dtype, objective=obj, minimize=minimize, n_chains=n, dim=d, max_iter=it, temperature_iter=ti,
                     temperature_max=tmax, seed=seed, chain_offset=offset, max_steps=max_steps)
    return B.sann_run(lib, cfg, x0)


def assert_chains_match(res, ao, tol):
    assert np.array_equal(res["n_accepted"], ao["n_accepted"])
    assert np.array_equal(res["n_improved"], ao["n_improved"])
    assert rel_close(res["f_best"], ao["f_best"], tol)
    assert rel_close(res["x_best"], ao["x_best"], tol)
    assert rel_close(res["p_cur"], ao["p_cur"], tol)


CASES = [
    # dtype, objective, minimize, n_chains, d, max_iter, temperature_iter, temperature_max, shared start
    (B.F64, B.SPHERE, True, 6, 7, 200, 10, 10.0, True),          # 4-lane groups
    (B.F64, B.ROSENBROCK, True, 37, 2, 300, 10, 10.0, False),
    (B.F64, B.ROSENBROCK_EX, True, 9, 13, 150, 10, 10.0, False),  # 8-lane groups, pairwise objective
    (B.F64, B.RASTRIGIN, True, 21, 30, 100, 10, 10.0, False),     # 16-lane groups
    (B.F64, B.ACKLEY, False, 11, 64, 60, 10, 10.0, True),         # one full warp per chain
    (B.F64, B.RASTRIGIN, True, 5, 67, 60, 5, 3.0, False),         # two sweeps per candidate
    (B.F64, B.ROSENBROCK, True, 3, 200, 30, 10, 10.0, False),     # four sweeps, pairwise carry across sweeps
    (B.F64, B.STYBLINSKI_TANG, True, 300, 9, 40, 4, 2.0, False),  # more chains than one block holds
    (B.F64, B.BEALE, True, 16, 2, 200, 10, 10.0, False),          # closed form
    (B.F64, B.SHEKEL, True, 10, 4, 100, 10, 10.0, True),
    (B.F64, B.SPHERE, True, 4, 5, 30, 1, 10.0, True),             # temperature_iter = 1: no candidates
    (B.F32, B.SPHERE, True, 6, 7, 100, 10, 10.0, True),
    (B.F32, B.RASTRIGIN, True, 8, 40, 50, 10, 10.0, False),
]


@pytest.mark.parametrize("dtype,obj,minimize,n,d,it,ti,tmax,shared", CASES)
def test_sann_chains_match_oracle(ctx, oracle_lib, dtype, obj, minimize, n, d, it, ti, tmax, shared):
    seed = 1000 + 13 * d + n
    x0 = start_points(n, d, shared, seed)
    ch = gpu_chains(ctx, dtype, obj, minimize, n, d, it, ti, tmax, seed, x0, offset=3)
    ch.run()
    st = ch.sync()
    res = ch.chains()
    best = ch.best()
    ch.close()
    so, ao = oracle_chains(oracle_lib, dtype, obj, minimize, n, d, it, ti, tmax, seed, x0, offset=3)
    assert_chains_match(res, ao, tol_of(dtype, obj))
    assert st["iterations"] == it and st["stopped"] == 1 and st["stop_reason"] == 1
    assert st["function_calls"] == so["function_calls"] == n * (1 + it * max(ti - 1, 0))
    assert st["best_index"] == 3 + so["best_index"]
    assert rel_close(st["f_value"], so["f_value"], tol_of(dtype, obj))
    assert rel_close(best[None, :], ao["x_best"][so["best_index"]][None, :], tol_of(dtype, obj))


@pytest.mark.parametrize("path", golden_files("sann_"), ids=os.path.basename)
def test_sann_chains_match_reference_fixture(ctx, path):
    """Against the committed output of the UNMODIFIED reference (tests/golden, written by make_golden.py)."""
    cfg, x0, z = load_sann(path)
    ch = gpu_chains(ctx, cfg.dtype, cfg.objective, bool(cfg.minimize), cfg.n_chains, cfg.dim, cfg.max_iter,
                    cfg.temperature_iter, cfg.temperature_max, cfg.seed, x0)
    ch.run()
    st = ch.sync()
    res = ch.chains()
    ch.close()
    tol = tol_of(cfg.dtype, cfg.objective)
    assert rel_close(res["f_best"], z["f_best"], tol) and rel_close(res["x_best"], z["x_best"], tol)
    assert st["best_index"] == z["best_index"].item() and st["function_calls"] == z["function_calls_total"].item()
    # draws the reference consumed = 2d per candidate + one per Metropolis test, and a Metropolis test happens exactly
    # when a candidate is not an improvement: the device's improvement counters must reproduce the reference's draw count
    steps = cfg.max_iter * (cfg.temperature_iter - 1)
    assert np.array_equal(z["draws"], 2 * cfg.dim * steps + (steps - res["n_improved"].astype(np.uint64)))


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
@pytest.mark.parametrize("dtype,obj,d", [(B.F64, B.RASTRIGIN, 67), (B.F64, B.ROSENBROCK, 45), (B.F64, B.ACKLEY, 150),
                                         (B.F32, B.SPHERE, 200), (B.F64, B.STYBLINSKI_TANG, 20)])
def test_sann_lane_group_width_never_changes_a_decision(ctx, oracle_lib, monkeypatch, lanes, dtype, obj, d):
    """The lane-group width is a tuning choice (NLS_SANN_LANES overrides the row-size policy): narrow groups sweep a long
    row in several passes with 32 / W accumulator slots per lane, which reproduces the canonical 32-accumulator summation
    order bit for bit — so every width must give the same chains."""
    monkeypatch.setenv("NLS_SANN_LANES", str(lanes))
    n, it, seed = 19, 25, 300 + d
    x0 = start_points(n, d, False, seed)
    ch = gpu_chains(ctx, dtype, obj, True, n, d, it, 10, 10.0, seed, x0)
    ch.run()
    ch.sync()
    res = ch.chains()
    ch.close()
    so, ao = oracle_chains(oracle_lib, dtype, obj, True, n, d, it, 10, 10.0, seed, x0)
    assert_chains_match(res, ao, tol_of(dtype, obj))
    monkeypatch.setenv("NLS_SANN_LANES", "32")
    ch = gpu_chains(ctx, dtype, obj, True, n, d, it, 10, 10.0, seed, x0)
    ch.run()
    ch.sync()
    wide = ch.chains()
    ch.close()
    for k in ("x_best", "p_cur", "f_best"):
        assert np.array_equal(bits(res[k]), bits(wide[k])), k      # identical bits, not merely within tolerance


def test_sann_stepwise_equals_one_shot_and_oracle_cut_points(ctx, oracle_lib):
    dtype, obj, n, d, it, ti, tmax, seed = B.F64, B.RASTRIGIN, 13, 19, 40, 10, 10.0, 5
    x0 = start_points(n, d, False, seed)
    whole = gpu_chains(ctx, dtype, obj, True, n, d, it, ti, tmax, seed, x0)
    whole.run()
    whole.sync()
    ref = whole.chains()
    whole.close()
    ch = gpu_chains(ctx, dtype, obj, True, n, d, it, ti, tmax, seed, x0)
    done = 0
    for k in (1, 7, 100, 13, 1000):
        ch.step(k)
        done = min(done + k, it * (ti - 1))
        st = ch.sync()
        assert st["function_calls"] == n * (1 + done)
        assert st["iterations"] == (it if done == it * (ti - 1) else done // (ti - 1))
        assert st["stopped"] == int(done == it * (ti - 1))
        if done < it * (ti - 1):
            so, ao = oracle_chains(oracle_lib, dtype, obj, True, n, d, it, ti, tmax, seed, x0, max_steps=done)
            assert_chains_match(ch.chains(), ao, 1e-12)
    res = ch.chains()
    ch.close()
    for k in ("x_best", "p_cur", "f_best"):
        assert np.array_equal(bits(res[k]), bits(ref[k])), k     # launch boundaries never change a bit
    assert np.array_equal(res["n_accepted"], ref["n_accepted"])


def test_sann_shards_compose(ctx):
    """Chains are keyed by global id: two handles owning halves of a batch reproduce the whole batch bit for bit."""
    dtype, obj, n, d, it, ti, tmax, seed = B.F64, B.ACKLEY, 50, 12, 30, 10, 10.0, 77
    x0 = start_points(n, d, False, seed)
    whole = gpu_chains(ctx, dtype, obj, True, n, d, it, ti, tmax, seed, x0)
    whole.run()
    ws = whole.sync()
    ref = whole.chains()
    whole.close()
    parts, stats = [], []
    for lo, hi in ((0, 23), (23, 50)):
        ch = gpu_chains(ctx, dtype, obj, True, hi - lo, d, it, ti, tmax, seed, x0[lo:hi], offset=lo)
        ch.run()
        stats.append(ch.sync())
        parts.append(ch.chains())
        ch.close()
    for k in ("x_best", "p_cur", "f_best", "n_accepted", "n_improved"):
        assert np.array_equal(bits(np.concatenate([p[k] for p in parts])), bits(ref[k])), k
    best = min(stats, key=lambda s: (s["f_value"], s["best_index"]))
    assert (best["f_value"], best["best_index"]) == (ws["f_value"], ws["best_index"])


def test_sann_mirror_class(ctx, oracle_lib):
    """nb.SANN mirrors nlsolver::SANN: minimize(x) overwrites x, two seed draws per call, f_evals accumulates."""
    draws = iter([0.25, 0.75, 0.5, 0.125, 0.3, 0.6, 0.9])
    solver = nb.SANN(nb.RosenbrockExample, lambda: next(draws), max_iter=300, ctx=ctx)
    x = [5.0, 5.0]
    st = solver.minimize(x)
    seed = (0x40000000 << 32) | 0xC0000000
    so, ao = oracle_chains(oracle_lib, B.F64, B.ROSENBROCK_EX, True, 1, 2, 300, 10, 10.0, seed, np.array([5.0, 5.0]))
    assert st.iteration == 300 and st.function_calls_used == 2701
    assert rel_close(st.f_value, so["f_value"], 1e-12) and rel_close(np.array([x]), ao["x_best"], 1e-12)
    st2 = solver.minimize(x)
    assert st2.function_calls_used == 2 * 2701              # nlsolver.h:2751: f_evals is never reset
    xs, sts = solver.minimize_batch(np.full((8, 2), 5.0))
    assert xs.shape == (8, 2) and len(sts) == 8 and next(draws) == 0.9
    assert all(s.function_calls_used == 2701 for s in sts)


def test_sann_many_chains_find_the_minimum(ctx):
    """Multi-start: the best of 4096 chains on Rastrigin d = 4 lands in the global basin; one chain rarely does."""
    cfg = nb.sann_cfg(objective=nb.RASTRIGIN, n_chains=4096, dim=4, max_iter=300, seed=2024)
    ch = nb.SANNChains(ctx, cfg, np.full(4, 3.3))
    ch.run()
    st = ch.sync()
    res = ch.chains()
    x = ch.best()
    ch.close()
    assert st["f_value"] == res["f_best"].min() and st["best_index"] == int(np.argmin(res["f_best"]))
    assert st["f_value"] < 0.5 and np.all(np.abs(x) < 0.1)
    assert np.median(res["f_best"]) > st["f_value"]


def test_sann_invalid_arguments(ctx):
    with pytest.raises(nb.NlsError):
        nb.SANNChains(ctx, nb.sann_cfg(n_chains=0, dim=2), np.zeros(2))
    with pytest.raises(nb.NlsError):
        nb.SANNChains(ctx, nb.sann_cfg(n_chains=4, dim=2, temperature_max=0.0), np.zeros(2))
    with pytest.raises(nb.NlsError):
        nb.SANNChains(ctx, nb.sann_cfg(objective=nb.BEALE, n_chains=4, dim=3), np.zeros(3))
    with pytest.raises(nb.NlsError):
        nb.SANNChains(ctx, nb.sann_cfg(n_chains=4, dim=2), np.zeros((3, 2)))     # 3 start rows for 4 chains
