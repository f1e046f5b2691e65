"""Randomised shapes against the oracle: population sizes around tile / block boundaries, dimensions around the
lane-group thresholds (W = 4 / 8 / 16 / 32 lanes per agent), both dtypes, all objectives, both strategies / PSO types."""
import numpy as np
import pytest

import nlsolver_b200 as nb
from oracle import binding as B
from tests.gpu_util import bits, gpu_de, oracle_de, rel_close, tolerance

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(20260118)
DIMS = [1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 129, 200]
POPS = [4, 5, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1000, 2049]


def de_case():
    return (int(RNG.integers(0, 2)), int(RNG.integers(0, 5)), int(RNG.integers(0, 2)), bool(RNG.integers(0, 2)),
            int(RNG.choice(POPS)), int(RNG.choice(DIMS)), int(RNG.integers(1, 5)), float(RNG.choice([0.3, 0.9, 1.0])),
            float(RNG.choice([0.5, 0.8])), int(RNG.integers(1, 1 << 62)))


@pytest.mark.parametrize("case", [de_case() for _ in range(40)], ids=lambda c: "-".join(str(v) for v in c[:7]))
def test_de_random_shapes(case):
    dtype, obj, strategy, minimize, P, d, G, cr, f, seed = case
    ctx = nb.Context(0)
    x0 = np.full(d, 3.0)
    tol = tolerance(dtype, obj)
    pop = gpu_de(ctx, dtype, obj, strategy, minimize, P, d, seed, x0, cr=cr, f=f)
    pop.step(G)
    st = pop.sync()
    so, ao = oracle_de(B.oracle(), dtype, obj, strategy, minimize, P, d, G, seed, x0, cr=cr, f=f)
    dec = pop.decisions(masks=True)
    for k in ("donors", "dim_idx", "rejects", "masks", "accepted"):
        assert np.array_equal(dec[k], ao[k]), k
    rows = pop.population()
    if tol == 0.0:
        assert np.array_equal(bits(rows), bits(ao["rows"])) and np.array_equal(bits(pop.scores()), bits(ao["scores"]))
    else:
        assert rel_close(rows, ao["rows"], tol) and rel_close(pop.scores(), ao["scores"], tol)
    assert st["best_index"] == so["best_index"] and st["iterations"] == G
    pop.close()
    ctx.close()


def pso_case():
    ptype = int(RNG.integers(0, 2))
    d = int(RNG.choice(DIMS))
    P = int(RNG.choice([1, 2, 7, 32, 33, 100, 257, 1025]))
    return (int(RNG.integers(0, 2)), int(RNG.integers(0, 5)), ptype, bool(RNG.integers(0, 2)), bool(RNG.integers(0, 2)),
            P, d, int(RNG.integers(1, 6)), int(RNG.integers(1, 1 << 62)))


@pytest.mark.parametrize("case", [pso_case() for _ in range(40)], ids=lambda c: "-".join(str(v) for v in c[:8]))
def test_pso_random_shapes(case):
    dtype, obj, ptype, minimize, constrained, P, d, G, seed = case
    ctx = nb.Context(0)
    up = np.full(d, 4.0)
    social_j = ptype == B.PSO_VANILLA and P > d      # the reference's [i] indexing is undefined there
    exact = ptype == B.PSO_VANILLA and obj in (B.SPHERE, B.ROSENBROCK, B.ROSENBROCK_EX)
    tol = 0.0 if exact else (1e-12 if dtype == B.F64 else 5e-5)
    kw = dict(dtype=dtype, objective=obj, pso_type=ptype, minimize=minimize, n_particles=P, dim=d, eps=0.0,
              best_val_no_change=1 << 40, constrained=constrained, seed=seed)
    sw = nb.PSOSwarm(ctx, nb.pso_cfg(max_iter=1 << 40, flags=nb.FLAG_SOCIAL_INDEX_J if social_j else 0, **kw), -up, up)
    sw.step(G)
    st = sw.sync()
    so, ao = B.pso_run(B.oracle(), B.pso_cfg(max_iter=G, social_index_j=social_j, **kw), -up, up)
    assert st["iterations"] == G and st["best_valid"] == so["best_valid"]
    pos = sw.positions()
    if tol == 0.0:
        assert np.array_equal(bits(pos), bits(ao["positions"]))
        assert np.array_equal(bits(sw.pbest_values()), bits(ao["pbest_values"]))
        assert st["f_value"] == so["f_value"] and st["best_index"] == so["best_index"]
    else:
        assert rel_close(pos, ao["positions"], tol) and rel_close(sw.pbest_values(), ao["pbest_values"], tol)
        assert rel_close(st["f_value"], so["f_value"], tol)
    sw.close()
    ctx.close()
