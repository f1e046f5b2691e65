"""The device `log_unit` (nlsolver_b200/csrc/pso_impl.cuh) uses only IEEE operations and the table log_table.h, so its host
model tools/log_unit_check.c is bit-identical to it; this pins the model's accuracy against glibc (the reference's log),
checks that the model and the device source still carry the same coefficients and the same table, and that the committed
table is what tools/gen_log_table.py writes."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_log_unit_model_is_close_to_glibc(tmp_path):
    exe = tmp_path / "log_unit_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tools", "log_unit_check.c"), "-lm",
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), "3000000"], check=True, capture_output=True, text=True).stdout
    m = re.search(r"max error ([0-9.]+) ulp .* bit-equal to glibc in ([0-9.]+)%", out)
    assert m, out
    assert float(m.group(1)) < 1.5 and float(m.group(2)) > 75.0, out
    assert "x=1  ours=0  glibc=0" in out          # log(1) is exactly 0: a draw of 1.0 gives rnorm = 0


def test_device_source_and_host_model_share_coefficients_and_table():
    dev = open(os.path.join(ROOT, "nlsolver_b200", "csrc", "pso_impl.cuh")).read()
    host = open(os.path.join(ROOT, "tools", "log_unit_check.c")).read()
    block = dev[dev.index("kLogCoef[8]"):dev.index("// Square root for an operand of KNOWN range")]
    coef = re.findall(r"-?[0-9]\.[0-9]+e[-+][0-9]+", block)
    assert len(coef) == 8
    for c in coef:
        assert c.lstrip("-") in host, f"coefficient {c} of the device log_unit is missing from tools/log_unit_check.c"
    for magic in ("0x95f64", "0x3ff00000", "0x000fffff", "0x7f", "log_table.h", "NLS_LOG_TABLE_ROWS"):
        assert magic in dev and magic in host, magic


def test_committed_log_table_is_what_the_generator_writes(tmp_path):
    out = tmp_path / "log_table.h"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_log_table.py"), str(out)], check=True,
                   capture_output=True)
    assert out.read_text() == open(os.path.join(ROOT, "nlsolver_b200", "csrc", "log_table.h")).read()
