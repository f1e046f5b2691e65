"""The build's own report (nvcc -Xptxas -v, kept per translation unit under nlsolver_b200/csrc/build/ by the Makefile) is
checked, not just collected: the long-row and reduction kernels the BASELINE configurations run in fp64 must not spill
at all, the short-row kernels of the d = 64 sweep at most 16 bytes, and no shipped kernel may spill more than 192 bytes
(a launch bound that starts to spill a hot loop shows up here, on CPU, before it shows up as a slower number)."""
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOGS = sorted(glob.glob(os.path.join(ROOT, "nlsolver_b200", "csrc", "build", "*.ptxas.log")))

# (demangled-name pattern, what runs it)
NO_SPILL = [
    (r"de_generation_bulk_kernel<double, 2, 2, 2>", "config 2: DE-random Rastrigin d=1000"),
    (r"de_generation_bulk_kernel<double, 1, 2, 2>", "config 4: DE-best Rosenbrock d=4096"),
    (r"de_commit_kernel<double>", "configs 2 / 4 / 5: commit + reduce"),
    (r"pso_move_kernel<double, 3, 1, 32, 1, 1>", "config 3: accelerated PSO, Ackley d=256"),
    (r"pso_candidate_kernel<double>|pso_apply_kernel<double>|pso_candidate_publish_kernel<double>|pso_gather_apply_kernel<double>",
     "config 3: min-loc reduction and exchange"),
    (r"de_generation_kernel<double, 0, 16, 2, 2, false>", "config 5: DE Sphere d=64 fp64"),
]
# short-row kernels that keep two steps of every row plus the trip's draws in registers at 80 registers: a few bytes
FEW_BYTES = [
    (r"pso_move_kernel<(double|float), 0, 0, (8|16), 2, (2|4)>", "config 5: vanilla PSO Sphere d=64"),
    (r"de_generation_kernel<float, 0, 8, 2, 4, false>", "config 5: DE Sphere d=64 fp32"),
]
MAX_FEW_BYTES = 16
MAX_SPILL_BYTES = 192


def kernels():
    rows = []
    for path in LOGS:
        for block in re.split(r"ptxas info\s+: Compiling entry function '", open(path).read())[1:]:
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", block)
            r = re.search(r"Used (\d+) registers", block)
            if m and r:
                rows.append([block.split("'")[0], int(r.group(1)), int(m.group(2)), int(m.group(3))])
    names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True, check=True)
    for row, name in zip(rows, names.stdout.splitlines()):
        row[0] = name.replace("nls::", "").split("(")[0].replace("void ", "")
    return rows


@pytest.mark.skipif(not LOGS, reason="no build logs: run __graft_entry__.build() first")
def test_headline_kernels_do_not_spill_and_no_kernel_spills_much():
    rows = kernels()
    assert len(rows) > 500, "the ptxas logs of every translation unit are expected"
    for pattern, what in NO_SPILL:
        hit = [r for r in rows if re.search(pattern, r[0])]
        assert hit, f"no kernel matches {pattern} ({what})"
        for name, regs, st, ld in hit:
            assert st == 0 and ld == 0, f"{name} ({what}) spills {st} / {ld} bytes at {regs} registers"
    for pattern, what in FEW_BYTES:
        hit = [r for r in rows if re.search(pattern, r[0])]
        assert hit, f"no kernel matches {pattern} ({what})"
        for name, regs, st, ld in hit:
            assert st <= MAX_FEW_BYTES, f"{name} ({what}) spills {st} bytes at {regs} registers"
    worst = max(rows, key=lambda r: r[2])
    assert worst[2] <= MAX_SPILL_BYTES, f"{worst[0]} spills {worst[2]} bytes"
