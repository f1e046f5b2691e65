"""CPU-side checks of the objective-plugin toolchain: the example plugins cross-compile for sm_100a with nvcc, export the
entry point, and register through nls_load_objective (loading a plugin needs no GPU; running it does)."""
import os
import shutil
import subprocess

import pytest

import nlsolver_b200 as nb
from nlsolver_b200 import plugins

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="objective plugins are built with nvcc")


@pytest.mark.parametrize("example", ["styblinski_tang.cu", "beale.cu"])
def test_example_plugins_build_and_register(tmp_path, example):
    so = tmp_path / ("lib" + example.replace(".cu", ".so"))
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler",
                    "-fPIC", "-I", os.path.join(ROOT, "nlsolver_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "objectives", example), "-o", str(so)], check=True, capture_output=True)
    syms = subprocess.run(["nm", "-D", str(so)], check=True, capture_output=True, text=True).stdout
    assert " T nls_objective_plugin_v1" in syms
    oid = plugins.load_objective(str(so))
    assert oid >= 100


def test_loading_something_else_fails_cleanly(tmp_path):
    bogus = tmp_path / "libbogus.so"
    src = tmp_path / "bogus.c"
    src.write_text("int nothing(void) { return 0; }\n")
    subprocess.run(["gcc", "-shared", "-fPIC", str(src), "-o", str(bogus)], check=True)
    with pytest.raises(nb.NlsError) as e:
        plugins.load_objective(str(bogus))
    assert "nls_objective_plugin_v1" in str(e.value)
    with pytest.raises(nb.NlsError):
        plugins.load_objective(str(tmp_path / "missing.so"))


def test_compile_objective_reports_compiler_errors(tmp_path):
    with pytest.raises(RuntimeError) as e:
        plugins.compile_objective("template <class T> struct Broken { this is not C++ };", "Broken", out_dir=str(tmp_path))
    assert "nvcc failed" in str(e.value)
