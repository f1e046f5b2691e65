"""BASELINE.json's full sizes, where the oracle would take hours: size-independent properties of the hot path plus spot
checks of sampled rows against the oracle's objective."""
import numpy as np
import pytest

import nlsolver_b200 as nb
from oracle import binding as B

pytestmark = pytest.mark.gpu


def test_config2_full_size_properties():
    """DE-random, Rastrigin, d = 1000, P = 2^20, fp64 (BASELINE configs[1])."""
    P, d, G = 1 << 20, 1000, 3
    ctx = nb.Context(0)
    pop = nb.DEPopulation(ctx, nb.de_cfg(objective=nb.RASTRIGIN, pop_size=P, dim=d, crossover_prob=0.9,
                                         differential_weight=0.8, eps=0.0, max_iter=1 << 40, best_val_no_change=1 << 40,
                                         seed=0x7c26ca28fb68bc1b), np.full(d, 10.24))
    rng = np.random.default_rng(1)
    sample = np.sort(rng.choice(P, size=64, replace=False))
    prev = pop.scores()
    # initial population: uniform in [-5.12, 5.12] (generate_sequence, nlsolver.h:2302-2312), scores = objective(rows)
    for i in sample[:16]:
        row = pop.rows(int(i), 1)[0]
        assert np.all(np.abs(row) <= 5.12) and abs(B.objective(B.F64, B.RASTRIGIN, row) - prev[i]) <= 1e-12 * prev[i]
    # E[10 + x^2 - 10 cos(2 pi x)] per coordinate for x uniform in [-5.12, 5.12]
    expect = 10 + 5.12 ** 2 / 3 - 10 * np.sin(2 * np.pi * 5.12) / (2 * np.pi * 5.12)
    assert abs(prev.mean() / d - expect) < 0.01
    for g in range(1, G + 1):
        pop.step(1)
        st = pop.sync()
        cur = pop.scores()
        dec = pop.decisions()
        assert st["iterations"] == g and st["function_calls"] == P * (g + 1)
        assert np.all(cur <= prev)                                        # greedy selection never worsens a score
        assert np.array_equal(dec["accepted"].astype(bool), cur < prev)
        assert np.array_equal(np.where(dec["accepted"] == 1, dec["trial_scores"], prev), cur)
        don = dec["donors"].astype(np.int64)
        idx = np.arange(P)
        assert don.min() >= 0 and don.max() < P and np.all(don != idx[:, None])
        assert np.all(don[:, 0] != don[:, 1]) and np.all(don[:, 0] != don[:, 2]) and np.all(don[:, 1] != don[:, 2])
        assert dec["dim_idx"].max() < d and np.all(dec["rejects"] <= 3)
        assert st["best_index"] == int(np.argmin(cur)) and st["f_value"] == cur.min()
        # donor indices are uniform: each third of the population gets a third of the picks
        assert np.all(np.abs(np.bincount(don.ravel() * 3 // P, minlength=3) / don.size - 1 / 3) < 2e-3)
        # the trial score of a sampled agent is the objective of the trial rebuilt on the host from the donors' rows
        prev = cur
    for i in sample:
        row = pop.rows(int(i), 1)[0]
        assert abs(B.objective(B.F64, B.RASTRIGIN, row) - cur[i]) <= 1e-12 * cur[i]
    pop.close()
    ctx.close()


def test_config3_shard_full_size_properties():
    """Accelerated PSO, Ackley, d = 256, 2^21 particles (one GPU's share of BASELINE configs[2])."""
    P, d, G = 1 << 21, 256, 4
    up = np.full(d, 32.768)
    ctx = nb.Context(0)
    sw = nb.PSOSwarm(ctx, nb.pso_cfg(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, n_particles=P, dim=d, eps=0.0,
                                     max_iter=1 << 40, best_val_no_change=1 << 40, seed=0x7c26ca28fb68bc1b), -up, up)
    st = sw.sync()
    pbest, best = sw.pbest_values(), st["f_value"]
    assert best == sw.last_values().min() and st["best_index"] == int(np.argmin(sw.last_values()))
    for g in range(1, G + 1):
        sw.step(1)
        st = sw.sync()
        last, cur = sw.last_values(), sw.pbest_values()
        assert st["iterations"] == g and st["function_calls"] == P * (g + 1)
        assert np.array_equal(cur, np.minimum(pbest, last))               # particle_best_values (nlsolver.h:2730-2732)
        assert st["f_value"] == min(best, last.min())                     # strict-< running minimum (:2723-2729)
        if last.min() < best:
            assert st["best_index"] == int(np.argmin(last))
        pbest, best = cur, st["f_value"]
    pos = sw.positions()
    for i in np.random.default_rng(2).choice(P, size=32, replace=False):
        assert abs(B.objective(B.F64, B.ACKLEY, pos[i]) - last[i]) <= 1e-12 * abs(last[i])
    assert np.array_equal(sw.best(), pos[st["best_index"]]) or st["f_value"] < last.min()
    sw.close()
    ctx.close()
