"""The drop-in C++ header on the GPU: examples/example_de_pso.cpp (the reference's example.cpp DE / PSO call sites) is
built with g++, run, and its solver_status::print() output compared with the same solves through the Python mirror
(same generator, hence the same two seed draws) — both go through nls_de_solve / nls_pso_solve."""
import os
import re
import subprocess

import numpy as np
import pytest

import nlsolver_b200 as nb
from tests.test_gpu_convergence import XorShift

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Xoshiro:
    """rng::xoshiro<double> (nlsolver.h:1289-1341) as a Python callable, quirks included (s[2] = s[3] = 0 after seeding,
    rotation by 45)."""

    def __init__(self):
        m = (1 << 64) - 1
        z = (12374563468 + 0x9E3779B97F4A7C15) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        s0 = z ^ (z >> 31)
        self.s = [s0, s0 >> 32, 0, 0]

    def __call__(self):
        m = (1 << 64) - 1
        s = self.s
        result = (s[0] + s[3]) & m
        t = (s[1] << 17) & m
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = ((s[3] << 45) & m) | (s[3] >> 19)
        return float(np.float64(result) / np.float64(2.0 ** 64))


def parse(out):
    """-> list of (title, calls, iterations, f_value string, x strings) per printed solver block."""
    blocks = []
    for m in re.finditer(r"([^\n]*): ?\nFunction calls used: (\d+)\nAlgorithm iterations used: (\d+)\n"
                         r"With final function value of ([^\n]+)\n([^\n]*)\n", out):
        blocks.append((m.group(1).strip(), int(m.group(2)), int(m.group(3)), m.group(4), m.group(5)))
    return blocks


def test_example_program_matches_python_mirror(tmp_path):
    exe = tmp_path / "example_de_pso"
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "example_de_pso.cpp"), "-L", os.path.join(ROOT, "nlsolver_b200"),
                    "-lnls_b200", "-Wl,-rpath," + os.path.join(ROOT, "nlsolver_b200"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    blocks = parse(out)
    assert len(blocks) >= 7, out
    gen = XorShift()                      # example: DE-best with a fresh xorshift, x0 = {2, 7}
    x = [2.0, 7.0]
    st = nb.DE(nb.RosenbrockExample, gen, recombination=nb.RecombinationStrategy.best).minimize(x)
    title, calls, iters, fval, xs = blocks[0]
    assert (calls, iters) == (st.function_calls_used, st.iteration)
    assert fval == "%g" % st.f_value and xs == "".join("%g," % v for v in x)
    gen = XorShift()                      # README snippet: DE-random after gen.reset(), x0 = {5, 7}
    x = [5.0, 7.0]
    st = nb.DE(nb.RosenbrockExample, gen).minimize(x)
    title, calls, iters, fval, xs = blocks[1]
    assert (calls, iters) == (st.function_calls_used, st.iteration) and fval == "%g" % st.f_value
    gen = XorShift()                      # PSO vanilla (10 particles > 2 dims -> corrected social index), x0 = {3, 3}
    x = [3.0, 3.0]
    st = nb.PSO(nb.RosenbrockExample, gen).minimize(x)
    title, calls, iters, fval, xs = blocks[2]
    assert (calls, iters) == (st.function_calls_used, st.iteration) and fval == "%g" % st.f_value
    gen = Xoshiro()                       # example.cpp:216-223: SANN with a fresh xoshiro, x0 = {5, 5}
    x = [5.0, 5.0]
    sann = nb.SANN(nb.RosenbrockExample, gen)
    st = sann.minimize(x)
    title, calls, iters, fval, xs = blocks[5]
    assert "Annealing" in title and (calls, iters) == (st.function_calls_used, st.iteration) == (45001, 5000)
    assert fval == "%g" % st.f_value and xs == "".join("%g," % v for v in x)
    x = [5.0, 5.0]                        # 4096 chains on the same solver object: f_evals keeps accumulating
    st = sann.minimize_multistart(x, 4096)
    title, calls, iters, fval, xs = blocks[6]
    assert (calls, iters) == (st.function_calls_used, st.iteration) == (4097 * 45001, 5000)
    assert fval == "%g" % st.f_value and xs == "".join("%g," % v for v in x)
    # the big Rastrigin run at the end re-evaluates its result on the host with the header's own functor
    m = re.search(r"With final function value of ([^\n]+)\nhost re-evaluation of the returned point: ([^\n]+)", out)
    assert m and abs(float(m.group(1)) - float(m.group(2))) <= 1e-5 * abs(float(m.group(2)))


def test_plugin_example_program(tmp_path):
    import shutil
    if shutil.which("nvcc") is None:
        pytest.skip("objective plugins are built with nvcc")
    so = tmp_path / "libstyblinski_tang.so"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler",
                    "-fPIC", "-I", os.path.join(ROOT, "nlsolver_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "objectives", "styblinski_tang.cu"), "-o", str(so)], check=True,
                   capture_output=True)
    exe = tmp_path / "example_plugin_objective"
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "example_plugin_objective.cpp"), "-L",
                    os.path.join(ROOT, "nlsolver_b200"), "-lnls_b200", "-Wl,-rpath," + os.path.join(ROOT, "nlsolver_b200"),
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), str(so)], check=True, capture_output=True, text=True).stdout
    xs = [float(v) for v in out.strip().splitlines()[-1].rstrip(",").split(",")]
    assert len(xs) == 8 and np.allclose(xs, -2.903534, atol=2e-2), out


def test_nmpso_example_program_matches_python_mirror(tmp_path):
    """nlsolver::NelderMeadPSO through the header (one solver, then a batch of 512) == the Python mirror fed the same
    generator draws: both go through nls_nmpso_solve, so every figure is identical."""
    exe = tmp_path / "example_nmpso"
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "example_nmpso.cpp"), "-L", os.path.join(ROOT, "nlsolver_b200"),
                    "-lnls_b200", "-Wl,-rpath," + os.path.join(ROOT, "nlsolver_b200"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    gen = XorShift()
    nm = nb.NelderMeadPSO(nb.Rosenbrock, gen)
    x = [2.0, 5.0]
    st = nm.minimize(x)
    title, calls, iters, fval, _ = parse(out)[0]
    assert "NelderMeadPSO" in title and (calls, iters) == (st.function_calls_used, st.iteration) and fval == "%g" % st.f_value
    lines = out.strip().splitlines()
    assert lines[4] == "".join("%.17g," % v for v in x)
    n, d = 512, 8
    starts = np.full((n, d), 1.5)
    for c in range(n):
        starts[c, c % d] += 0.001 * (c + 1)
    xs, res = nm.minimize_batch(starts)
    best = min(range(n), key=lambda c: (res[c].f_value, c))
    m = re.search(r"batch of 512 solvers: iterations (\d+) calls (\d+) best solver (\d+) f (\S+)", out)
    assert m and int(m.group(1)) == sum(r.iteration for r in res) and int(m.group(2)) == sum(r.function_calls_used for r in res)
    assert int(m.group(3)) == best and m.group(4) == "%.17g" % res[best].f_value
    assert lines[-1] == "".join("%.17g," % v for v in xs[best])
