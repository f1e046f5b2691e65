"""The N > 1 host path on CPU: world_size-2 `gloo` process groups driving nlsolver_b200.distributed with the
oracle-backed engines of tests/cpu_engines.py.

  * sharded PSO over 2 ranks == the single-process oracle swarm, bit for bit (positions, best, counters);
  * island DE: each island equals an independent oracle DE up to the first migration, and a harness-level
    restatement (oracle steppers + explicit ring exchange in one process) afterwards — the reference has no way to
    inject migrants, so that is what parity means there."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import nlsolver_b200 as nb
from nlsolver_b200 import distributed as D
from oracle import binding as B
from tests.cpu_engines import OracleDEEngine, OraclePSOEngine, OracleSANNEngine, oracle_de_cfg


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_ranks(fn, world, *args):
    port = free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return dict(out)


def _entry(fn, rank, world, port, q, *args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world, *args)))
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------ pure host logic -------------------------------
def test_slice_bounds_cover_the_swarm():
    for n, w in ((10, 3), (16, 8), (7, 8), (1 << 24, 8), (5, 1)):
        spans = [D.slice_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[r][1] == spans[r + 1][0] for r in range(w - 1))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_ring_and_schedule():
    assert [D.ring_neighbors(r, 4) for r in range(4)] == [(1, 3), (2, 0), (3, 1), (0, 2)]
    assert [g for g in range(0, 31) if D.migration_due(g, 10)] == [10, 20, 30]
    assert not D.migration_due(10, 0)


def test_record_roundtrip_and_select():
    rec = D.pack_record(3.5, 17, (4.0, 1.0, 2.0), np.arange(5.0))
    assert rec.size == D.record_bytes(8, 5)
    h = D.parse_record(rec)
    assert (h["value"], h["index"], h["n"], h["mean"], h["m2"], h["valid"]) == (3.5, 17, 4.0, 1.0, 2.0, 1)
    recs = [{"value": 2.0, "valid": 1}, {"value": 1.0, "valid": 1}, {"value": 1.0, "valid": 1}, {"value": 0.5, "valid": 0}]
    assert D.select_best(recs) == 1                    # strict <: the first of equal values wins, invalid skipped
    assert D.select_best(recs, running_best=1.0) == -1  # nothing strictly below the running best
    x = np.random.default_rng(0).normal(size=1000)
    parts = [(float(c.size), float(c.mean()), float(((c - c.mean()) ** 2).sum())) for c in np.array_split(x, 7)]
    assert abs(D.std_err_from_moments(*D.merge_moments(parts)) - x.std(ddof=1)) < 1e-12


# ------------------------------------------------------------------ sharded PSO -----------------------------------
PSO_KW = dict(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, n_particles=37, dim=6, eps=0.0, max_iter=1 << 40,
              best_val_no_change=1 << 40, seed=11)


def _sharded_pso(rank, world, kw, gens):
    up = np.full(kw["dim"], 32.768)
    sw = D.ShardedPSO(nb.pso_cfg(**kw), -up, up, engine_factory=OraclePSOEngine)
    sw.step(gens)
    st = sw.sync()
    return st, sw.best(), sw.engine.positions(), D.slice_bounds(kw["n_particles"], world, rank)


@pytest.mark.parametrize("kw", [PSO_KW, dict(PSO_KW, pso_type=nb.PSO_VANILLA, n_particles=6, dim=8, objective=nb.SPHERE),
                                dict(PSO_KW, constrained=True, objective=nb.RASTRIGIN)])
def test_sharded_pso_two_ranks_equals_single_swarm(kw):
    gens = 12
    up = np.full(kw["dim"], 32.768)
    okw = {k: v for k, v in kw.items()}
    ocfg = B.pso_cfg(**dict(okw, max_iter=gens))
    so, ao = B.pso_run(B.oracle(), ocfg, -up, up)
    out = run_ranks(_sharded_pso, 2, kw, gens)
    whole = np.zeros_like(ao["positions"])
    for rank, (st, best, pos, (b, e)) in out.items():
        for k in ("f_value", "iterations", "function_calls", "best_index", "val_no_change"):
            assert st[k] == so[k], (rank, k, st[k], so[k])
        assert np.array_equal(best, ao["x_best"])
        whole[b:e] = pos
    assert np.array_equal(whole, ao["positions"])


# ------------------------------------------------------------------ island DE -------------------------------------
DE_KW = dict(objective=nb.ROSENBROCK, strategy=nb.DE_BEST, pop_size=40, dim=5, eps=0.0, max_iter=1 << 40,
             best_val_no_change=1 << 40, seed=5)


def _islands(rank, world, kw, gens, every, k, exchange="nccl"):
    isl = D.IslandDE(nb.de_cfg(**kw), np.full(kw["dim"], 4.096), migrate_every=every, migrants=k,
                     engine_factory=OracleDEEngine, exchange=exchange)
    isl.step(gens)
    st = isl.sync()
    _, a = isl.engine.s.report()
    return st, a["rows"], a["scores"], isl.global_best_row()


def _restated_islands(kw, world, gens, every, k):
    """Harness-level restatement: `world` oracle steppers in one process, explicit ring exchange."""
    x0 = np.full(kw["dim"], 4.096)
    isl = [B.DEStepper(oracle_de_cfg(nb.de_cfg(**dict(kw, agent_offset=r * kw["pop_size"]))), x0) for r in range(world)]
    for g in range(1, gens + 1):
        for s in isl:
            s.advance(1)
        if D.migration_due(g, every):
            out = [s.export_top(k) for s in isl]
            for r, s in enumerate(isl):
                rows, scores = out[D.ring_neighbors(r, world)[1]]
                s.import_migrants(rows, scores)
    return [s.report() for s in isl]


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_islands_two_ranks_before_and_after_migration(exchange):
    """exchange="nccl": export + all-gather after every generation; "peer": the host logic of the fused path — one
    engine call per migration interval, the records read out of the window between two barriers (the window itself is
    emulated by tests/cpu_engines.py; the real one is checked on GPUs by tests/tools/multi_gpu_check.py)."""
    every, k = 4, 3
    for gens in (3, 9):     # before the first migration; after two of them
        out = run_ranks(_islands, 2, DE_KW, gens, every, k, exchange)
        want = _restated_islands(DE_KW, 2, gens, every, k)
        best = min(range(2), key=lambda r: (want[r][0]["f_value"], r))
        for rank, (st, rows, scores, grow) in out.items():
            wst, wa = want[rank]
            assert np.array_equal(rows, wa["rows"]) and np.array_equal(scores, wa["scores"]), (gens, rank)
            for key in ("f_value", "iterations", "function_calls", "best_index"):
                assert st[key] == wst[key], (gens, rank, key)
            assert st["global_best_value"] == want[best][0]["f_value"] and st["global_best_rank"] == best
            assert np.array_equal(grow, want[best][1]["x_best"])
        if gens < every:    # no migration yet: every island is exactly an independent reference-exact DE run
            for rank in range(2):
                cfg = oracle_de_cfg(nb.de_cfg(**dict(DE_KW, agent_offset=rank * DE_KW["pop_size"], max_iter=gens)))
                so, ao = B.de_run(B.oracle(), cfg, np.full(DE_KW["dim"], 4.096))
                assert np.array_equal(out[rank][1], ao["rows"])


# ------------------------------------------------------------------ sharded SANN chains ---------------------------
SANN_KW = dict(objective=nb.RASTRIGIN, n_chains=11, dim=5, max_iter=30, temperature_iter=10, seed=21)


def _sharded_sann(rank, world, kw, x0, first):
    job = D.ShardedSANN(nb.sann_cfg(**kw), x0, engine_factory=OracleSANNEngine)
    job.step(first)
    mid = job.sync()
    job.run()
    st, row = job.global_best()
    return mid, st, row, job.engine.chains(), (job.begin, job.end)


@pytest.mark.parametrize("shared", [True, False])
def test_sharded_sann_two_ranks_equals_one_batch(shared):
    kw = SANN_KW
    x0 = np.full(kw["dim"], 2.5) if shared else np.random.default_rng(3).uniform(-3, 3, size=(kw["n_chains"], kw["dim"]))
    so, ao = B.sann_run(B.oracle(), B.sann_cfg(**kw), x0)
    out = run_ranks(_sharded_sann, 2, kw, x0, 100)
    whole_x, whole_f = np.zeros_like(ao["x_best"]), np.zeros_like(ao["f_best"])
    for rank, (mid, st, row, chains, (b, e)) in out.items():
        assert mid["function_calls"] == (e - b) * 101 and mid["iterations"] == 100 // 9 and not mid["stopped"]
        assert (st["f_value"], st["best_index"]) == (so["f_value"], so["best_index"])      # same answer on every rank
        assert st["function_calls"] == so["function_calls"] and st["iterations"] == kw["max_iter"]
        assert np.array_equal(row, ao["x_best"][so["best_index"]])
        whole_x[b:e], whole_f[b:e] = chains["x_best"], chains["f_best"]
    assert np.array_equal(whole_x, ao["x_best"]) and np.array_equal(whole_f, ao["f_best"])


def test_sharded_sann_three_ranks_uneven_slices():
    kw = dict(SANN_KW, n_chains=10, objective=nb.ACKLEY, minimize=False)     # slices of 4, 3, 3 chains; maximize
    x0 = np.random.default_rng(8).uniform(-2, 2, size=(10, kw["dim"]))
    so, ao = B.sann_run(B.oracle(), B.sann_cfg(**kw), x0)
    out = run_ranks(_sharded_sann, 3, kw, x0, 50)
    assert sorted(e - b for _, _, _, _, (b, e) in out.values()) == [3, 3, 4]
    for rank, (mid, st, row, chains, (b, e)) in out.items():
        assert (st["f_value"], st["best_index"]) == (so["f_value"], so["best_index"])
        assert np.array_equal(row, ao["x_best"][so["best_index"]])
        assert np.array_equal(chains["f_best"], ao["f_best"][b:e]) and np.array_equal(chains["p_cur"], ao["p_cur"][b:e])
