"""INTEGRATION.md §3 ("Option B") shows the stub a maintainer of the reference would paste into nlsolver.h.  This test
pastes exactly that text into a SCRATCH copy of the reference header (under tmp_path — nothing of the reference enters
the repo, /root/reference is only read), compiles a program that calls DE / PSO / SANN with a device-tagged objective,
and links it against libnls_b200.so.  Without a GPU the program must fail loudly in nls_ctx_create (no CPU fallback);
objectives without the tag must still take the reference's own CPU path."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

MARKERS = {
    "helper": "// nlsolver.h, next to the includes",
    "de": "// DE::solve, first lines",
    "pso": "// PSO::solve<minimize, constrained>, first lines",
    "sann": "// SANN::solve<minimize>, first lines",
}

PROGRAM = r"""
#define NLSOLVER_WITH_B200 1
#include "nlsolver.h"
#include <cstdio>
#include <cstring>
// a device-tagged objective: the tag selects the GPU branch, operator() keeps the untaken CPU branch well-formed
struct DeviceRosenbrock {
  static constexpr int nls_objective = NLS_ROSENBROCK_EX;
  double operator()(std::vector<double> &x) { const double a = 1 - x[0], b = x[1] - x[0] * x[0]; return a * a + 100 * b * b; }
};
struct HostSphere {   // no tag: the reference's own CPU path
  double operator()(std::vector<double> &x) { return x[0] * x[0] + x[1] * x[1]; }
};
int main(int argc, char **argv) {
  nlsolver::rng::xorshift<double> gen;
  if (argc > 1 && !std::strcmp(argv[1], "host")) {
    HostSphere f;
    nlsolver::DE<HostSphere, nlsolver::rng::xorshift<double>, double> de(f, gen);
    std::vector<double> x = {5, 7};
    auto st = de.minimize(x);
    std::printf("host %g\n", std::get<2>(st.get_summary()));
    return 0;
  }
  DeviceRosenbrock f;
  try {
    std::vector<double> x = {5, 7};
    nlsolver::DE<DeviceRosenbrock, nlsolver::rng::xorshift<double>, double> de(f, gen);
    de.minimize(x).print();
    nlsolver::PSO<DeviceRosenbrock, nlsolver::rng::xorshift<double>, double, nlsolver::PSOType::Accelerated> pso(f, gen);
    x = {3, 3};
    pso.minimize(x).print();
    nlsolver::SANN<DeviceRosenbrock, nlsolver::rng::xorshift<double>, double> sann(f, gen);
    x = {5, 5};
    sann.minimize(x).print();
  } catch (const std::exception &e) {
    std::printf("error: %s\n", e.what());
    return 3;
  }
  return 0;
}
"""


def stub_blocks():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = text[text.index("## 3. Option B"):text.index("## 4. Entry points")]
    code = "\n".join(re.findall(r"```cpp\n(.*?)```", section, flags=re.S))
    starts = sorted((code.index(m), k) for k, m in MARKERS.items())
    blocks = {}
    for n, (pos, key) in enumerate(starts):
        end = starts[n + 1][0] if n + 1 < len(starts) else len(code)
        blocks[key] = code[pos:end]
    return blocks


def insert_after(lines, predicate, block, what):
    for i, line in enumerate(lines):
        if predicate(i, line):
            return lines[:i + 1] + block.splitlines() + lines[i + 1:]
    raise AssertionError(f"anchor for {what} not found in the reference header")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "nlsolver.h")), reason="reference tree absent")
def test_documented_stub_compiles_into_the_reference_header(tmp_path):
    blocks = stub_blocks()
    assert set(blocks) == set(MARKERS)
    lines = open(os.path.join(REF, "nlsolver.h")).read().splitlines()

    def in_class(name):
        start = next(i for i, ln in enumerate(lines) if ln.startswith(f"class {name} {{"))
        return lambda i, ln: i > start and ln.strip().startswith("solver_status<scalar_t> solve(std::vector<scalar_t> &x) {")

    # innermost first so that earlier insertions do not move later anchors: SANN (:2778) > PSO (:2593) > DE (:2414)
    lines = insert_after(lines, in_class("SANN"), blocks["sann"], "SANN::solve")
    lines = insert_after(lines, in_class("PSO"), blocks["pso"], "PSO::solve")
    lines = insert_after(lines, in_class("DE"), blocks["de"], "DE::solve")
    lines = insert_after(lines, lambda i, ln: ln.startswith('#include "./tinyqr.h"'), blocks["helper"], "the include block")
    (tmp_path / "nlsolver.h").write_text("\n".join(lines) + "\n")
    for name in ("tinyqr.h", "utils.h"):
        (tmp_path / name).write_text(open(os.path.join(REF, name)).read())
    (tmp_path / "main.cpp").write_text(PROGRAM)
    exe = tmp_path / "stub_demo"
    subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-I", str(tmp_path), "-I", os.path.join(ROOT, "include"),
                    str(tmp_path / "main.cpp"), "-L", os.path.join(ROOT, "nlsolver_b200"), "-lnls_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "nlsolver_b200"), "-o", str(exe)], check=True)
    host = subprocess.run([str(exe), "host"], capture_output=True, text=True)
    assert host.returncode == 0 and host.stdout.startswith("host "), host.stdout + host.stderr
    import torch
    dev = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert dev.returncode == 0 and dev.stdout.count("Function calls used") == 3, dev.stdout + dev.stderr
    else:   # no GPU: the device branch must refuse, never fall back to the CPU loop
        assert dev.returncode == 3 and "no CPU path" in dev.stdout, dev.stdout + dev.stderr
