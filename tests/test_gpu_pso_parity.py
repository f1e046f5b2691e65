"""PSO on the B200 (through the C ABI) against the oracle on the same draw tape, generation by generation."""
import os

import numpy as np
import pytest

import nlsolver_b200 as nb
from oracle import binding as B
from tests.golden_util import golden_files, load_pso
from tests.gpu_util import bits, rel_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context(0)
    yield c
    c.close()


def pso_tol(dtype, obj, ptype):
    # the accelerated move calls log / sqrt / cos per coordinate; vanilla with Sphere / Rosenbrock is + - * only
    exact = ptype == B.PSO_VANILLA and obj in (B.SPHERE, B.ROSENBROCK, B.ROSENBROCK_EX)
    if exact:
        return 0.0
    if dtype == B.F64:
        return 1e-12
    # fp32: rnorm is evaluated in double and rounded once, as the reference's rnorm<float> does (nlsolver.h:2479-2485
    # under libstdc++), so with a + - * objective results agree to 2 ulp of float; cosf / expf objectives keep 2e-5
    return 2.4e-7 if obj in (B.SPHERE, B.ROSENBROCK, B.ROSENBROCK_EX) else 2e-5


CASES = [
    # dtype, objective, type, minimize, constrained, P, d, G, bound
    (B.F64, B.SPHERE, B.PSO_VANILLA, True, False, 8, 8, 25, 10.24),
    (B.F64, B.ROSENBROCK, B.PSO_VANILLA, True, True, 6, 8, 25, 4.096),
    (B.F64, B.SPHERE, B.PSO_VANILLA, True, False, 100, 130, 10, 5.12),
    (B.F64, B.RASTRIGIN, B.PSO_VANILLA, True, False, 30, 33, 10, 5.12),
    (B.F64, B.SPHERE, B.PSO_ACCELERATED, True, False, 40, 8, 20, 10.24),
    (B.F64, B.ACKLEY, B.PSO_ACCELERATED, True, False, 1000, 32, 8, 32.768),
    (B.F64, B.ACKLEY, B.PSO_ACCELERATED, True, False, 300, 256, 5, 32.768),     # config-3 row shape
    (B.F64, B.RASTRIGIN, B.PSO_ACCELERATED, True, True, 40, 8, 30, 5.12),
    (B.F64, B.ROSENBROCK_EX, B.PSO_ACCELERATED, False, True, 33, 5, 10, 2.0),
    (B.F32, B.SPHERE, B.PSO_ACCELERATED, True, False, 100, 16, 8, 10.24),
    (B.F32, B.SPHERE, B.PSO_VANILLA, True, True, 12, 16, 8, 10.24),
]


@pytest.mark.parametrize("dtype,obj,ptype,minimize,constrained,P,d,G,bound", CASES)
def test_pso_matches_oracle_every_generation(ctx, oracle_lib, dtype, obj, ptype, minimize, constrained, P, d, G, bound):
    seed = 42 + d * 7 + P
    up = np.full(d, bound)
    tol = pso_tol(dtype, obj, ptype)
    cfg = nb.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, minimize=minimize, n_particles=P, dim=d, eps=0.0,
                     max_iter=1 << 40, best_val_no_change=1 << 40, constrained=constrained, seed=seed)
    swarm = nb.PSOSwarm(ctx, cfg, -up, up)
    for g in range(G + 1):
        if g:
            swarm.step(1)
        st = swarm.sync()
        ocfg = B.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, minimize=minimize, n_particles=P, dim=d, eps=0.0,
                         max_iter=g, best_val_no_change=1 << 40, constrained=constrained, seed=seed)
        so, ao = B.pso_run(oracle_lib, ocfg, -up, up)
        assert st["iterations"] == so["iterations"] == g and st["function_calls"] == so["function_calls"]
        assert st["best_valid"] == so["best_valid"] and st["val_no_change"] == so["val_no_change"], g
        got = {"positions": swarm.positions(), "pbest_values": swarm.pbest_values(), "last_values": swarm.last_values()}
        if ptype == B.PSO_VANILLA:
            got["velocities"] = swarm.velocities()
        for k, v in got.items():
            if tol == 0.0:
                assert np.array_equal(bits(v), bits(ao[k])), (g, k)
            else:
                assert rel_close(v, ao[k], tol), (g, k, np.max(np.abs(v - ao[k])))
        if so["best_valid"]:
            assert st["best_index"] == so["best_index"], g
            assert (np.array_equal(bits(swarm.best()), bits(ao["x_best"])) if tol == 0.0
                    else rel_close(swarm.best(), ao["x_best"], tol))
            assert st["f_value"] == so["f_value"] if tol == 0.0 else rel_close(st["f_value"], so["f_value"], tol)
    swarm.close()


@pytest.mark.parametrize("path", golden_files("pso_"), ids=os.path.basename)
def test_pso_matches_reference_fixture(ctx, path):
    cfg, up, z = load_pso(path)
    tol = pso_tol(cfg.dtype, cfg.objective, cfg.pso_type)
    ncfg = nb.pso_cfg(dtype=cfg.dtype, objective=cfg.objective, pso_type=cfg.pso_type, minimize=bool(cfg.minimize),
                      n_particles=cfg.n_particles, dim=cfg.dim, eps=0.0, max_iter=cfg.max_iter,
                      best_val_no_change=1 << 40, constrained=bool(cfg.constrained), seed=cfg.seed)
    swarm = nb.PSOSwarm(ctx, ncfg, -up, up)
    swarm.step(cfg.max_iter)
    st = swarm.sync()
    assert st["stopped"] and st["iterations"] == z["iterations"].item()
    assert st["function_calls"] == z["function_calls"].item()
    for k, v in (("positions", swarm.positions()), ("pbest_values", swarm.pbest_values()), ("x_best", swarm.best())):
        assert (np.array_equal(bits(v), bits(z[k])) if tol == 0.0 else rel_close(v, z[k], tol)), k
    assert st["f_value"] == z["f_value"].item() if tol == 0.0 else rel_close(st["f_value"], z["f_value"].item(), tol)
    swarm.close()


def test_pso_stop_rules_on_device(ctx, oracle_lib):
    for ptype, P, d in ((B.PSO_VANILLA, 8, 8),):
        up = np.full(d, 3.0)
        so, ao = B.pso_run(oracle_lib, B.pso_cfg(objective=B.SPHERE, pso_type=ptype, n_particles=P, dim=d, seed=5), -up, up)
        swarm = nb.PSOSwarm(ctx, nb.pso_cfg(objective=nb.SPHERE, pso_type=ptype, n_particles=P, dim=d, seed=5), -up, up)
        swarm.step(so["iterations"] + 10)
        st = swarm.sync()
        assert st["stopped"] and st["stop_reason"] == so["stop_reason"] and st["iterations"] == so["iterations"]
        assert st["f_value"] == so["f_value"] and np.array_equal(bits(swarm.best()), bits(ao["x_best"]))
        swarm.close()


def test_pso_vanilla_more_particles_than_dims_needs_corrected_flag(ctx, oracle_lib):
    up = np.full(4, 2.0)
    with pytest.raises(nb.NlsError):    # the reference reads out of bounds here (nlsolver.h:2674)
        nb.PSOSwarm(ctx, nb.pso_cfg(n_particles=10, dim=4), -up, up)
    cfg = nb.pso_cfg(n_particles=10, dim=4, flags=nb.FLAG_SOCIAL_INDEX_J, eps=0.0, max_iter=1 << 40,
                     best_val_no_change=1 << 40, seed=3)
    swarm = nb.PSOSwarm(ctx, cfg, -up, up)
    swarm.step(15)
    st = swarm.sync()
    so, ao = B.pso_run(oracle_lib, B.pso_cfg(n_particles=10, dim=4, social_index_j=True, eps=0.0, max_iter=15,
                                             best_val_no_change=1 << 40, seed=3), -up, up)
    assert st["iterations"] == 15 and np.array_equal(bits(swarm.positions()), bits(ao["positions"]))
    swarm.close()


def test_sharded_swarm_equals_single_swarm(ctx):
    """Two shards on one GPU exchanging candidate records (what the multi-GPU path all-gathers) == one swarm."""
    import torch
    P, d, G = 600, 48, 6
    up = np.full(d, 32.768)
    kw = dict(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, dim=d, eps=0.0, max_iter=1 << 40,
              best_val_no_change=1 << 40, seed=77)
    whole = nb.PSOSwarm(ctx, nb.pso_cfg(n_particles=P, **kw), -up, up)
    cut = 250
    shards = [nb.PSOSwarm(ctx, nb.pso_cfg(n_particles=cut, particle_offset=0, n_particles_global=P, **kw), -up, up),
              nb.PSOSwarm(ctx, nb.pso_cfg(n_particles=P - cut, particle_offset=cut, n_particles_global=P, **kw), -up, up)]
    rb = nb.lib().nls_record_bytes(nb.F64, d)
    rec = torch.zeros(2 * rb, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()     # torch filled `rec` on its own stream; the context stream is independent of it
    for k, s in enumerate(shards):
        s.export_candidate(rec.data_ptr() + k * rb)
    for s in shards:
        s.apply_candidates(rec.data_ptr(), 2)
    for g in range(G):
        whole.step(1)
        for k, s in enumerate(shards):
            s.step_local(rec.data_ptr() + k * rb)
        for s in shards:
            s.sync()
        for s in shards:
            s.apply_candidates(rec.data_ptr(), 2)
    sw = whole.sync()
    for s in shards:
        ss = s.sync()
        for k in ("f_value", "iterations", "function_calls", "best_index", "val_no_change"):
            assert ss[k] == sw[k], (k, ss[k], sw[k])
        # the stop statistic merges per-shard moments in a different order than the single swarm's block tree
        assert abs(ss["std_err"] - sw["std_err"]) <= 1e-12 * abs(sw["std_err"])
        assert np.array_equal(bits(s.best()), bits(whole.best()))
    assert np.array_equal(bits(np.concatenate([s.positions() for s in shards])), bits(whole.positions()))


def test_fused_peer_exchange_single_rank_equals_plain_step(ctx):
    """nls_pso_step_fused with a one-rank exchange window (records published into the rank's own window) must equal
    nls_pso_step; the multi-rank case needs several GPUs and is checked by tests/tools/multi_gpu_check.py."""
    P, d, G = 500, 40, 7
    up = np.full(d, 5.12)
    kw = dict(objective=nb.RASTRIGIN, pso_type=nb.PSO_ACCELERATED, n_particles=P, dim=d, eps=0.0, max_iter=1 << 40,
              best_val_no_change=1 << 40, seed=99)
    plain = nb.PSOSwarm(ctx, nb.pso_cfg(**kw), -up, up)
    fused = nb.PSOSwarm(ctx, nb.pso_cfg(**kw), -up, up)
    win = nb.ExchangeWindow(ctx, nb.lib().nls_record_bytes(nb.F64, d), 1, 0)
    fused.attach_exchange(win)
    plain.step(G)
    fused.step_fused(G)
    a, b = plain.sync(), fused.sync()
    for k in ("f_value", "iterations", "function_calls", "best_index", "val_no_change", "std_err"):
        assert a[k] == b[k], k
    assert np.array_equal(bits(plain.positions()), bits(fused.positions()))
    assert np.array_equal(bits(plain.best()), bits(fused.best()))
    fused.close()
    win.close()
    plain.close()


@pytest.mark.parametrize("dtype,obj,ptype,P,d", [(B.F64, B.ACKLEY, B.PSO_ACCELERATED, 1024, 64),
                                                (B.F64, B.SPHERE, B.PSO_VANILLA, 10, 16),
                                                (B.F32, B.SPHERE, B.PSO_ACCELERATED, 300, 40),
                                                (B.F64, B.ROSENBROCK, B.PSO_VANILLA, 64, 200)])
def test_pso_one_launch_path_equals_separate_kernels(ctx, monkeypatch, dtype, obj, ptype, P, d):
    """Swarms of up to 2^16 elements run a whole step in ONE launch on one thread-block cluster (pso_persistent_kernel);
    NLS_DE_ONE_LAUNCH=0 keeps the separate kernels.  Same bits either way."""
    up = np.full(d, 5.12)
    out = []
    for env in ("1", "0"):
        monkeypatch.setenv("NLS_DE_ONE_LAUNCH", env)
        sw = nb.PSOSwarm(ctx, nb.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, n_particles=P, dim=d, eps=0.0,
                                         max_iter=1 << 40, best_val_no_change=1 << 40, seed=5), -up, up)
        for n in (1, 9, 30):
            sw.step(n)
        st = sw.sync()
        out.append((st, sw.positions(), sw.pbest_values(), sw.best()))
        sw.close()
    (sa, pa, ba, xa), (sb, pb, bb, xb) = out
    assert sa["iterations"] == sb["iterations"] == 40
    for k in ("f_value", "function_calls", "best_index", "val_no_change", "std_err"):
        assert sa[k] == sb[k], k
    assert np.array_equal(bits(pa), bits(pb)) and np.array_equal(bits(ba), bits(bb)) and np.array_equal(bits(xa), bits(xb))


@pytest.mark.parametrize("dtype,obj,ptype,P,d", [(B.F64, B.ACKLEY, B.PSO_ACCELERATED, 3000, 64),
                                                (B.F32, B.SPHERE, B.PSO_VANILLA, 40, 48),
                                                (B.F64, B.RASTRIGIN, B.PSO_ACCELERATED, 700, 300)])
def test_pso_fused_candidate_apply_equals_the_two_kernels(ctx, monkeypatch, dtype, obj, ptype, P, d):
    """An unsharded swarm reduces and applies in ONE launch (the block that finishes the reduction applies its own
    record); the shard-style path — step_local, then apply_candidates on the exported record — is the two kernels.
    Same bits either way, stop statistic included."""
    import torch
    monkeypatch.setenv("NLS_DE_ONE_LAUNCH", "0")
    up = np.full(d, 5.12)
    kw = dict(dtype=dtype, objective=obj, pso_type=ptype, n_particles=P, dim=d, eps=0.0, max_iter=1 << 40,
              best_val_no_change=1 << 40, seed=11)
    a = nb.PSOSwarm(ctx, nb.pso_cfg(**kw), -up, up)
    b = nb.PSOSwarm(ctx, nb.pso_cfg(**kw), -up, up)
    rec = torch.zeros(nb.lib().nls_record_bytes(dtype, d), dtype=torch.uint8, device="cuda")
    for n in (1, 8, 11):
        a.step(n)
    for _ in range(20):
        b.step_local(rec.data_ptr())
        b.apply_candidates(rec.data_ptr(), 1)
    sa, sb = a.sync(), b.sync()
    assert sa["iterations"] == sb["iterations"] == 20
    for k in ("f_value", "function_calls", "best_index", "val_no_change", "std_err"):
        assert sa[k] == sb[k], k
    assert np.array_equal(bits(a.positions()), bits(b.positions())) and np.array_equal(bits(a.best()), bits(b.best()))
    assert np.array_equal(bits(a.pbest_values()), bits(b.pbest_values()))
    a.close()
    b.close()
