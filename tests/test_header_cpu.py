"""CPU checks of the drop-in C++ header: it compiles as plain C++17 against the C ABI, and its generators reproduce
the reference's streams (the latter only where /root/reference exists: the reference header is compiled in place)."""
import os
import subprocess

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
  dump("halton<double>", nlsolver::rng::halton<double>());
  dump("halton<float> base 3", nlsolver::rng::halton<float>(3));
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


def test_c_abi_header_is_plain_c(tmp_path):
    """include/nls_b200.h must be consumable from C (cgo / FFI generators read it as C)."""
    src = tmp_path / "abi.c"
    src.write_text('#include "nls_b200.h"\n'
                   'int main(void) { nls_de_cfg c; nls_pso_cfg p; nls_sann_cfg a; nls_status s; (void)c; (void)p; (void)a; (void)s;\n'
                   '  return nls_version() == NLS_B200_VERSION ? 0 : 1; }\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-L", os.path.join(ROOT, "nlsolver_b200"), "-lnls_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "nlsolver_b200"), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
