"""Objective plugins: a user objective compiled from functor source runs through the same DE / PSO
kernels; checked against the oracle with the same objective installed as Python callbacks (canonical lane order)."""
import numpy as np
import pytest

import nlsolver_b200 as nb
from nlsolver_b200 import plugins
from oracle import binding as B
from tests.gpu_util import bits, rel_close

import shutil

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(shutil.which("nvcc") is None, reason="objective plugins are built with nvcc")]

STYBLINSKI = """
template <class T> struct StyblinskiTang {          // f(x) = 0.5 * sum(x^4 - 16 x^2 + 5 x)  (test_functions.h StyblinskiTang, N-D)
  static constexpr bool pairwise = false;
  static __device__ T lane0_seed(unsigned) { return T(0); }
  static __device__ T term(T x, T, unsigned, unsigned) { const T x2 = x * x; return x2 * x2 - T(16) * x2 + T(5) * x; }
  static __device__ T finish(T sum, unsigned) { return T(0.5) * sum; }
};
"""
CHAIN = """
template <class T> struct Chain {                   // pairwise form: sum_{j>=1} (x_j - x_{j-1})^2 + 3, then sqrt
  static constexpr bool pairwise = true;
  static __device__ T lane0_seed(unsigned) { return T(3); }
  static __device__ T term(T x, T xp, unsigned, unsigned) { const T t = x - xp; return t * t; }
  static __device__ T finish(T sum, unsigned d) { return sqrt(sum) + T(d); }
};
"""


BEALE = """
template <class T> struct Beale {                   // test_functions.h:94-105, closed form over the whole 2-vector
  static constexpr unsigned full_dim = 2;
  static __device__ T full(const T (&x)[2]) {
    const T a = T(1.5) - x[0] + x[0] * x[1];
    const T b = T(2.25) - x[0] + x[0] * x[1] * x[1];
    const T c = T(2.625) - x[0] + x[0] * x[1] * x[1] * x[1];
    return a * a + b * b + c * c;
  }
};
"""
QUAD5 = """
template <class T> struct Quad5 {                   // a coupled 5-D form: spans three lanes of the 4-lane group in fp64
  static constexpr unsigned full_dim = 5;
  static __device__ T full(const T (&x)[5]) {
    T s = T(0);
    for (int k = 0; k < 5; k++) s = s + (x[k] - T(k)) * (x[k] - T(k));
    return s + x[0] * x[4] - x[1] * x[3];
  }
};
"""


def beale(x):
    a = 1.5 - x[0] + x[0] * x[1]
    b = 2.25 - x[0] + x[0] * x[1] * x[1]
    c = 2.625 - x[0] + x[0] * x[1] * x[1] * x[1]
    return a * a + b * b + c * c


def quad5(x):
    s = 0.0
    for k in range(5):
        s = s + (x[k] - float(k)) * (x[k] - float(k))
    return s + x[0] * x[4] - x[1] * x[3]


@pytest.fixture(scope="module")
def ctx():
    c = nb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def plugin_ids(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("objectives"))
    from concurrent.futures import ThreadPoolExecutor
    srcs = (("StyblinskiTang", STYBLINSKI), ("Chain", CHAIN), ("Beale", BEALE), ("Quad5", QUAD5))
    with ThreadPoolExecutor(4) as ex:     # nvcc instantiates every kernel for the functor: a few seconds each
        built = list(ex.map(lambda ns: plugins.compile_objective(ns[1], ns[0], out_dir=out, extra_flags=("-fmad=false",)),
                            srcs))
    ids = {name: plugins.load_objective(so) for (name, _), so in zip(srcs, built)}
    assert all(v >= 100 for v in ids.values()) and len(set(ids.values())) == 4
    return ids


def install(name):
    if name == "StyblinskiTang":
        def term(x, xp, j, d):
            x2 = x * x
            return x2 * x2 - 16.0 * x2 + 5.0 * x
        B.set_custom_objective(term, lambda s, d: 0.5 * s)
    else:
        def term(x, xp, j, d):
            t = x - xp
            return t * t
        B.set_custom_objective(term, lambda s, d: float(np.sqrt(s)) + float(d), pairwise=True, lane0_seed=3.0)


@pytest.mark.parametrize("name,strategy,P,d,G", [("StyblinskiTang", nb.DE_RANDOM, 96, 11, 6),
                                                 ("Chain", nb.DE_BEST, 64, 70, 5), ("StyblinskiTang", nb.DE_BEST, 40, 2, 8)])
def test_de_with_plugin_objective_matches_oracle(ctx, plugin_ids, name, strategy, P, d, G):
    install(name)
    x0 = np.full(d, 8.0)
    pop = nb.DEPopulation(ctx, nb.de_cfg(objective=plugin_ids[name], strategy=strategy, pop_size=P, dim=d, eps=0.0,
                                         max_iter=1 << 40, best_val_no_change=1 << 40, seed=31,
                                         flags=nb.FLAG_RECORD_MASKS), x0)
    pop.step(G)
    st = pop.sync()
    so, ao = B.de_run(B.oracle(), B.de_cfg(objective=B.CUSTOM, strategy=strategy, pop_size=P, dim=d, eps=0.0, max_iter=G,
                                           best_val_no_change=1 << 40, seed=31), x0, masks=True)
    dec = pop.decisions(masks=True)
    for k in ("donors", "dim_idx", "rejects", "masks", "accepted"):
        assert np.array_equal(dec[k], ao[k]), k
    if name == "StyblinskiTang":      # + - * only, compiled without FMA contraction: bit-exact
        assert np.array_equal(bits(pop.population()), bits(ao["rows"])) and st["f_value"] == so["f_value"]
    else:                             # sqrt in finish()
        assert rel_close(pop.population(), ao["rows"], 1e-12) and rel_close(st["f_value"], so["f_value"], 1e-12)
    assert st["best_index"] == so["best_index"]
    pop.close()


def test_pso_with_plugin_objective_matches_oracle(ctx, plugin_ids):
    install("StyblinskiTang")
    P, d, G = 50, 9, 8
    up = np.full(d, 5.0)
    kw = dict(pso_type=nb.PSO_ACCELERATED, n_particles=P, dim=d, eps=0.0, max_iter=1 << 40, best_val_no_change=1 << 40, seed=8)
    sw = nb.PSOSwarm(ctx, nb.pso_cfg(objective=plugin_ids["StyblinskiTang"], **kw), -up, up)
    sw.step(G)
    st = sw.sync()
    so, ao = B.pso_run(B.oracle(), B.pso_cfg(objective=B.CUSTOM, **dict(kw, max_iter=G)), -up, up)
    assert st["best_index"] == so["best_index"] and rel_close(st["f_value"], so["f_value"], 1e-12)
    assert rel_close(sw.positions(), ao["positions"], 1e-12)
    sw.close()


@pytest.mark.parametrize("name,d", [("StyblinskiTang", 9), ("Chain", 40)])
def test_sann_chains_with_plugin_objective_match_oracle(ctx, plugin_ids, name, d):
    install(name)
    n, it = 12, 40
    x0 = np.random.default_rng(d).uniform(-3, 3, size=(n, d))
    ch = nb.SANNChains(ctx, nb.sann_cfg(objective=plugin_ids[name], n_chains=n, dim=d, max_iter=it, seed=17), x0)
    ch.run()
    st = ch.sync()
    res = ch.chains()
    ch.close()
    so, ao = B.sann_run(B.oracle(), B.sann_cfg(objective=B.CUSTOM, n_chains=n, dim=d, max_iter=it, seed=17), x0)
    assert np.array_equal(res["n_accepted"], ao["n_accepted"]) and np.array_equal(res["n_improved"], ao["n_improved"])
    assert rel_close(res["f_best"], ao["f_best"], 1e-12) and rel_close(res["x_best"], ao["x_best"], 1e-12)
    assert st["best_index"] == so["best_index"]


def test_sann_chains_with_closed_form_plugin_match_oracle(ctx, plugin_ids):
    B.set_custom_full(quad5)
    n, it = 10, 60
    x0 = np.random.default_rng(5).uniform(-3, 3, size=(n, 5))
    ch = nb.SANNChains(ctx, nb.sann_cfg(objective=plugin_ids["Quad5"], n_chains=n, dim=5, max_iter=it, seed=23), x0)
    ch.run()
    ch.sync()
    res = ch.chains()
    ch.close()
    so, ao = B.sann_run(B.oracle(), B.sann_cfg(objective=B.CUSTOM, n_chains=n, dim=5, max_iter=it, seed=23), x0)
    assert np.array_equal(res["n_accepted"], ao["n_accepted"]) and rel_close(res["x_best"], ao["x_best"], 1e-12)
    assert rel_close(res["f_best"], ao["f_best"], 1e-12)


def test_plugin_objective_through_the_solver_mirror(plugin_ids):
    """DE(...).minimize with a plugin id as the Callable: converges to the known minimum x_j = -2.903534."""
    class Gen:
        def __init__(self):
            self.v = iter([0.3, 0.7])

        def __call__(self):
            return next(self.v)
    x = [4.0] * 6
    st = nb.DE(plugin_ids["StyblinskiTang"], Gen(), pop_size=400, max_iter=400, best_val_no_change=400, eps=0.0,
               differential_weight=0.5).minimize(x)
    assert st.iteration == 400 and np.allclose(x, -2.903534, atol=1e-3), x
    assert abs(st.f_value - 6 * -39.16616570377142) < 1e-3


def test_unknown_objective_ids_are_rejected(ctx):
    with pytest.raises(nb.NlsError):
        nb.DEPopulation(ctx, nb.de_cfg(objective=9999, pop_size=10, dim=2), np.ones(2))


@pytest.mark.parametrize("name,fn,d,strategy", [("Beale", beale, 2, nb.DE_RANDOM), ("Quad5", quad5, 5, nb.DE_BEST)])
def test_closed_form_plugin_matches_oracle(ctx, plugin_ids, name, fn, d, strategy):
    """full_dim mode: the objective sees the whole (short) vector — the shape of the reference's 2-D test problems."""
    B.set_custom_full(fn)
    P, G = 60, 12
    x0 = np.full(d, 9.0)
    pop = nb.DEPopulation(ctx, nb.de_cfg(objective=plugin_ids[name], strategy=strategy, pop_size=P, dim=d, eps=0.0,
                                         max_iter=1 << 40, best_val_no_change=1 << 40, seed=17,
                                         flags=nb.FLAG_RECORD_MASKS), x0)
    pop.step(G)
    st = pop.sync()
    so, ao = B.de_run(B.oracle(), B.de_cfg(objective=B.CUSTOM, strategy=strategy, pop_size=P, dim=d, eps=0.0, max_iter=G,
                                           best_val_no_change=1 << 40, seed=17), x0, masks=True)
    dec = pop.decisions(masks=True)
    for k in ("donors", "dim_idx", "rejects", "masks", "accepted"):
        assert np.array_equal(dec[k], ao[k]), k
    assert np.array_equal(bits(pop.population()), bits(ao["rows"])) and st["f_value"] == so["f_value"]
    pop.close()
    up = np.full(d, 4.5)
    kw = dict(pso_type=nb.PSO_ACCELERATED, n_particles=30, dim=d, eps=0.0, max_iter=1 << 40, best_val_no_change=1 << 40, seed=4)
    sw = nb.PSOSwarm(ctx, nb.pso_cfg(objective=plugin_ids[name], **kw), -up, up)
    sw.step(6)
    sp = sw.sync()
    so, ao = B.pso_run(B.oracle(), B.pso_cfg(objective=B.CUSTOM, **dict(kw, max_iter=6)), -up, up)
    assert sp["best_index"] == so["best_index"] and rel_close(sw.positions(), ao["positions"], 1e-12)
    sw.close()
    with pytest.raises(nb.NlsError):      # a closed form of dimension D only runs with dim == D
        nb.DEPopulation(ctx, nb.de_cfg(objective=plugin_ids[name], pop_size=10, dim=d + 1), np.ones(d + 1))


def test_reference_beale_problem_converges(plugin_ids):
    """The reference runs its solvers on Beale from x0 = (-0.5, -0.5) and expects (3, 0.5) +- 0.05
    (test_functions.h:94-105, 397-404, 431-432); same check through the mirror with the plugin objective."""
    class Gen:
        def __init__(self):
            self.v = iter([0.40764453281267443, 0.82621863718638611])

        def __call__(self):
            return next(self.v)
    x = [-0.5, -0.5]
    nb.DE(plugin_ids["Beale"], Gen()).minimize(x)
    assert abs(x[0] - 3.0) <= 0.05 and abs(x[1] - 0.5) <= 0.05, x
