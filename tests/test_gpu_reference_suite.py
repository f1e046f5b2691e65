"""The reference's test driver on the GPU (test_functions.h:431-524): its 15 enabled problems x its DE / PSO variants,
from x0 = (-0.5, ...) with default hyper-parameters and one generator shared across the solvers of a problem.

For every (problem, solver) the GPU run is compared with the oracle run on the same seed (iterations, calls; f and x
exact for the + - * objectives, 1e-9 otherwise), and the reference's own pass criterion (|x - minimum| <= 0.05,
test_functions.h:397-404) must give the same verdict for the GPU run and the oracle run.  Which problems pass depends
on the draws (the reference's DE-random/xorshift passes 13/15; McCormick is unbounded below outside its
usual box, so an unlucky seed walks away in both implementations alike) — a summary test asserts that DE passes the
large majority."""
import numpy as np
import pytest

import nlsolver_b200 as nb
from nlsolver_b200.solvers import seed_from_generator
from oracle import binding as B
from tests.gpu_util import rel_close
from tests.test_gpu_convergence import XorShift

pytestmark = pytest.mark.gpu

PROBLEMS = {  # id: (name, dim, minimum, exact arithmetic)
    nb.SPHERE: ("Sphere", 2, (0.0, 0.0), True), nb.ROSENBROCK: ("Rosenbrock", 2, (1.0, 1.0), True),
    nb.RASTRIGIN: ("Rastrigin", 2, (0.0, 0.0), False), nb.ACKLEY: ("Ackley", 2, (0.0, 0.0), False),
    nb.BEALE: ("Beale", 2, (3.0, 0.5), True), nb.GOLDSTEIN_PRICE: ("Goldstein_Price", 2, (0.0, -1.0), True),
    nb.THREE_HUMP_CAMEL: ("ThreeHumpCamel", 2, (0.0, 0.0), True), nb.MCCORMICK: ("McCormick", 2, (-0.54719, -1.54719), False),
    nb.SCHAFFER_N2: ("SchafferN2", 2, (0.0, 0.0), False), nb.STYBLINSKI_TANG: ("StyblinskiTang", 2, (-2.903534, -2.903534), True),
    nb.SHEKEL: ("Shekel", 4, (4.0, 4.0, 4.0, 4.0), True), nb.BOOTH: ("Booth", 2, (1.0, 3.0), True),
    nb.BUKIN_N6: ("BukinN6", 2, (-10.0, 1.0), False), nb.MATYAS: ("Matyas", 2, (0.0, 0.0), True),
    nb.LEVI_N13: ("LeviN13", 2, (1.0, 1.0), False),
}
VERDICTS = {}


@pytest.mark.parametrize("obj", sorted(PROBLEMS), ids=lambda o: PROBLEMS[o][0])
def test_problem_through_every_de_and_pso_variant(obj):
    name, d, minimum, exact = PROBLEMS[obj]
    gen, shadow = XorShift(), XorShift()
    x0 = [-0.5] * d
    variants = (("DE random", "de", B.DE_RANDOM), ("DE best", "de", B.DE_BEST), ("PSO vanilla", "pso", B.PSO_VANILLA),
                ("PSO accelerated", "pso", B.PSO_ACCELERATED))
    for label, kind, mode in variants:
        x = list(x0)
        if kind == "de":
            st = nb.DE(obj, gen, recombination=mode).minimize(x)
            so, ao = B.de_run(B.oracle(), B.de_cfg(objective=obj, strategy=mode, pop_size=50, dim=d,
                                                   seed=seed_from_generator(shadow)), x0)
        else:
            st = nb.PSO(obj, gen, pso_type=mode).minimize(x)
            up = np.abs(np.array(x0))
            so, ao = B.pso_run(B.oracle(), B.pso_cfg(objective=obj, pso_type=mode, n_particles=10, dim=d,
                                                     social_index_j=(mode == B.PSO_VANILLA),   # 10 particles > d
                                                     seed=seed_from_generator(shadow)), -up, up)
        assert gen.draws == shadow.draws
        exact_here = exact and not (kind == "pso" and mode == B.PSO_ACCELERATED)
        if exact_here:
            assert (st.iteration, st.function_calls_used) == (so["iterations"], so["function_calls"]), (name, label)
            assert st.f_value == so["f_value"] and x == ao["x_best"].tolist(), (name, label)
        else:
            # libm in the objective / the move: last-bit differences can shift the stop iteration by the noise in the
            # std_err statistic only if it sits exactly on eps; everything else must agree to 1e-12
            assert (st.iteration, st.function_calls_used) == (so["iterations"], so["function_calls"]), (name, label)
            # BukinN6 is 100*sqrt(|.|) around 0: last-bit position differences are amplified ~1e4 times in f
            tol = 1e-5 if obj == nb.BUKIN_N6 else 1e-9
            assert rel_close(st.f_value, so["f_value"], tol) and rel_close(np.array(x), ao["x_best"], tol), (name, label)
        ours = all(abs(x[k] - minimum[k]) <= 0.05 for k in range(d))
        theirs = all(abs(ao["x_best"][k] - minimum[k]) <= 0.05 for k in range(d))
        assert ours == theirs, (name, label, x, ao["x_best"])
        VERDICTS[(name, label)] = ours


def test_de_passes_the_reference_criterion_on_most_problems():
    de = {k: v for k, v in VERDICTS.items() if k[1].startswith("DE")}
    if len(de) < 2 * len(PROBLEMS):
        pytest.skip("runs after the per-problem tests")
    for label in ("DE random", "DE best"):
        passed = [k[0] for k, v in de.items() if k[1] == label and v]
        assert len(passed) >= 11, (label, sorted(set(p[0] for p in PROBLEMS.values()) - set(passed)))
