"""The device `log_unit` (nlsolver_b200/csrc/pso_impl.cuh) uses only IEEE operations, so its host model
tools/log_unit_check.c is bit-identical to it; this pins the model's accuracy against glibc (the reference's log) and
checks that the model and the device source still carry the same coefficients."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_log_unit_model_is_within_one_ulp_of_glibc(tmp_path):
    exe = tmp_path / "log_unit_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", os.path.join(ROOT, "tools", "log_unit_check.c"), "-lm",
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe), "3000000"], check=True, capture_output=True, text=True).stdout
    m = re.search(r"max error ([0-9.]+) ulp .* bit-equal to glibc in ([0-9.]+)%", out)
    assert m, out
    assert float(m.group(1)) < 1.0 and float(m.group(2)) > 90.0, out
    assert "x=1  ours=0  glibc=0" in out          # log(1) is exactly 0: a draw of 1.0 gives rnorm = 0


def test_device_source_and_host_model_share_the_coefficients():
    dev = open(os.path.join(ROOT, "nlsolver_b200", "csrc", "pso_impl.cuh")).read()
    host = open(os.path.join(ROOT, "tools", "log_unit_check.c")).read()
    block = dev[dev.index("kLogCoef[9]"):dev.index("__device__ __forceinline__ double log_unit")]
    coef = re.findall(r"[0-9]\.[0-9]+e[-+][0-9]+", block)
    assert len(coef) == 9
    for c in coef:
        assert c in host, f"coefficient {c} of the device log_unit is missing from tools/log_unit_check.c"
    for magic in ("0x95f64", "0x3ff00000", "0x000fffff"):
        assert magic in dev and magic in host
