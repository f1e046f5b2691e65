"""CPU checks of the drop-in C++ header: it compiles as plain C++17 against the C ABI, and its generators reproduce
the reference's streams (the latter only where /root/reference exists: the reference header is compiled in place)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/nlsolver.h"

DUMP = r"""
#include <cstdio>
#include "%s"
template <class G> void dump(const char* n, G g) { printf("%%s", n); for (int i = 0; i < 64; i++) printf(" %%.17g", (double)g()); printf("\n"); }
int main() {
  dump("xorshift<double>", nlsolver::rng::xorshift<double>());
  dump("xorshift<float>", nlsolver::rng::xorshift<float>());
  dump("xoshiro<double>", nlsolver::rng::xoshiro<double>());
  dump("xoshiro<float>", nlsolver::rng::xoshiro<float>());
  dump("recurrent<double>", nlsolver::rng::recurrent<double>());
  dump("recurrent<float>", nlsolver::rng::recurrent<float>());
  dump("splitmix<double>", nlsolver::rng::splitmix<double>());
  nlsolver::rng::xorshift<double> g; g(); g(); g.reset(); dump("xorshift reset", g);
}
"""


def build_and_run(tmp_path, name, header, extra):
    src = tmp_path / (name + ".cpp")
    src.write_text(DUMP % header)
    exe = tmp_path / name
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-w", str(src), "-o", str(exe)] + extra, check=True)
    return subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout


def test_example_compiles_against_the_header(tmp_path):
    exe = tmp_path / "example"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "example_de_pso.cpp"), "-L", os.path.join(ROOT, "nlsolver_b200"),
                    "-lnls_b200", "-Wl,-rpath," + os.path.join(ROOT, "nlsolver_b200"), "-o", str(exe)], check=True)
    import torch
    if not torch.cuda.is_available():   # without a GPU the program must fail loudly, not fall back
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree absent")
def test_header_generators_reproduce_reference_streams(tmp_path):
    lib = ["-L", os.path.join(ROOT, "nlsolver_b200"), "-lnls_b200", "-Wl,-rpath," + os.path.join(ROOT, "nlsolver_b200")]
    ours = build_and_run(tmp_path, "ours", "nlsolver_b200.hpp", ["-I", os.path.join(ROOT, "include")] + lib)
    theirs = build_and_run(tmp_path, "theirs", "nlsolver.h", ["-I", os.path.dirname(REF)])
    assert ours == theirs
