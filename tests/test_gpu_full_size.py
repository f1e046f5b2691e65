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
        prev = cur
    for i in sample:
        row = pop.rows(int(i), 1)[0]
        assert abs(B.objective(B.F64, B.RASTRIGIN, row) - cur[i]) <= 1e-12 * cur[i]
    pop.close()
    ctx.close()


# ------------------------------------------------------------------ trials rebuilt on the host from the tape -----
GOLDEN = 0x9E3779B97F4A7C15
M64 = (1 << 64) - 1


def mix64(z):
    """splitmix64 output function (nlsolver.h:1267-1270) on numpy uint64 arrays (wrapping arithmetic)."""
    z = np.asarray(z, np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def tape_keys(seed, gen, agents):
    """key(seed, gen, agent) = mix64(mix64(seed + GOLDEN * (gen + 1)) ^ agent)."""
    gk = mix64(np.array([(seed + GOLDEN * (gen + 1)) & M64], np.uint64))[0]
    return mix64(gk ^ np.asarray(agents, np.uint64))


def tape_draws(keys, k):
    """raw(key, k) = mix64(key + GOLDEN * (k + 1)); keys [n], k scalar or [n] or [n, m] (keys then broadcast)."""
    k = np.atleast_1d(np.asarray(k, np.uint64))
    keys = np.asarray(keys, np.uint64)
    if k.ndim == 2:
        keys = keys[:, None]
    return mix64(keys + np.full(1, GOLDEN, np.uint64) * (k + np.uint64(1)))


def unit64(raw):
    """T(raw) / T(2^64 - 1) for T = double (nlsolver.h:1358-1359): the divisor rounds to 2^64."""
    return np.asarray(raw, np.uint64).astype(np.float64) * 2.0 ** -64


def host_decisions(seed, gen, agents, P, d, fixed):
    """generate_indices + the forced coordinate (nlsolver.h:2331-2355, 2362) for `agents` from the draw tape."""
    keys = tape_keys(seed, gen, agents)
    n = len(agents)
    donors = np.zeros((n, 3), np.int64)
    rej = np.zeros(n, np.int64)
    k = np.zeros(n, np.uint64)
    for slot in range(3):
        todo = np.ones(n, bool)
        while todo.any():
            idx = np.minimum((unit64(tape_draws(keys, k)) * float(P)).astype(np.int64), P - 1)
            ok = idx != fixed
            for prev in range(slot):
                ok &= idx != donors[:, prev]
            take = todo & ok
            donors[take, slot] = idx[take]
            rej[todo & ~ok] += 1
            k[todo] += np.uint64(1)
            todo &= ~ok
    dim_idx = np.minimum((unit64(tape_draws(keys, k)) * float(d)).astype(np.int64), d - 1)
    return keys, donors, rej, dim_idx


def check_tape_against_oracle():
    """the numpy tape above == the oracle's own tape functions (which the reference harness is driven with)"""
    lib = B.oracle()
    for seed, gen, agent, k in [(0x7c26ca28fb68bc1b, 1, 0, 0), (5, 7, 123456, 1003), (M64, 3, 1 << 20, 4)]:
        key = lib.oracle_tape_key(seed, gen, agent)
        assert int(tape_keys(seed, gen, [agent])[0]) == key
        raw = lib.oracle_tape_draw(key, k)
        assert int(tape_draws(np.array([key], np.uint64), k)[0]) == raw
        assert unit64(np.array([raw], np.uint64))[0] == lib.oracle_unit_f64(raw)


@pytest.mark.parametrize("objective,F,warm,name", [
    (nb.RASTRIGIN, 0.8, 0, "config 2 proper (at d = 1000 and F = 0.8 no trial is ever accepted)"),
    (nb.SPHERE, 0.2, 4, "same shape, accepting regime: a ~ 4 % of the trials, the in-place repair is active")])
def test_config2_full_size_trials_rebuilt_from_the_tape(objective, F, warm, name):
    """At BASELINE's config-2 size (P = 2^20, d = 1000, fp64) the donors, forced coordinate, rejected proposals, accept
    flag (exact) and trial score (1e-12) of sampled agents equal a host rebuild from the draw tape and the donors' rows
    (generate_indices / propose_new_agent / greedy selection, nlsolver.h:2331-2375, 2449-2472) — including the in-place
    rule: a donor r < i that was accepted this generation contributes its NEW row."""
    check_tape_against_oracle()
    P, d, CR, seed = 1 << 20, 1000, 0.9, 0x7c26ca28fb68bc1b
    obj_b = {nb.RASTRIGIN: B.RASTRIGIN, nb.SPHERE: B.SPHERE}[objective]
    tol = 1e-12 if objective == nb.RASTRIGIN else 0.0
    ctx = nb.Context(0)
    pop = nb.DEPopulation(ctx, nb.de_cfg(objective=objective, pop_size=P, dim=d, crossover_prob=CR, differential_weight=F,
                                         eps=0.0, max_iter=1 << 40, best_val_no_change=1 << 40, seed=seed),
                          np.full(d, 10.24))
    pop.step(warm)
    rng = np.random.default_rng(3)
    checked = lower_accepted = accepted_seen = 0
    for g in range(warm + 1, warm + 3):
        st0 = pop.sync()
        assert st0["iterations"] == g - 1
        prev = pop.scores()
        # candidates: random agents, biased to high indices where lower donors are likely
        cand = np.unique(np.concatenate([rng.choice(P, 768, replace=False), P - 1 - rng.choice(P // 16, 256, replace=False)]))
        keys, donors, rej, dim_idx = host_decisions(seed, g, cand, P, d, fixed=cand)
        need = np.unique(np.concatenate([cand, donors.ravel()]))
        pre = {int(a): pop.rows(int(a), 1)[0] for a in need}
        pop.step(1)
        st = pop.sync()
        dec = pop.decisions()
        cur = pop.scores()
        assert np.array_equal(dec["donors"][cand].astype(np.int64), donors), "donor ids"
        assert np.array_equal(dec["rejects"][cand].astype(np.int64), rej), "rejected proposals"
        assert np.array_equal(dec["dim_idx"][cand].astype(np.int64), dim_idx), "forced coordinate"
        acc = dec["accepted"].astype(bool)
        # keep every candidate that has an accepted lower donor, fill up with the others to >= 256 agents
        hot = np.array([bool(np.any(acc[donors[n][donors[n] < a]])) for n, a in enumerate(cand)])
        order = np.concatenate([np.nonzero(hot)[0], np.nonzero(~hot)[0]])[:max(256, int(hot.sum()))]
        post = {}
        jj = np.arange(d, dtype=np.uint64)
        for n in order:
            a = int(cand[n])
            rows = []
            for r in (a, *[int(x) for x in donors[n]]):
                if r < a and acc[r]:                       # in place: r was processed before a and overwrote its row
                    if r not in post:
                        post[r] = pop.rows(r, 1)[0]
                    rows.append(post[r])
                    lower_accepted += 1
                else:
                    rows.append(pre[r])
            x0, x1, x2, x3 = rows
            u = unit64(tape_draws(keys[n:n + 1], (np.uint64(4 + rej[n]) + jj)[None, :])[0])
            mask = (u < CR) | (jj == np.uint64(dim_idx[n]))
            trial = np.where(mask, x1 + F * (x2 - x3), x0)
            score = B.objective(B.F64, obj_b, trial)
            got = dec["trial_scores"][a]
            assert abs(got - score) <= tol * abs(score), (name, g, a, got, score)
            margin = abs(score - prev[a])
            if margin > 1e-11 * abs(score):
                assert bool(acc[a]) == (score < prev[a]), (name, g, a, "accept flag")
            if acc[a]:
                accepted_seen += 1
                assert cur[a] == got
                new_row = pop.rows(a, 1)[0]
                assert np.array_equal(new_row, trial) if tol == 0.0 else np.all(np.abs(new_row - trial) <= 1e-12 * np.max(np.abs(trial)))
            else:
                assert cur[a] == prev[a]
            checked += 1
        assert st["iterations"] == g
    assert checked >= 512
    if objective == nb.SPHERE:
        assert lower_accepted > 0 and accepted_seen > 0, "the accepting case must exercise the in-place rule"
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
