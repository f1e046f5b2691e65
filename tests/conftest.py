import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (*.so is git-ignored): build the library and the CPU checkers once."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "nlsolver_b200", "libnls_b200.so")
    if not os.path.exists(lib) and shutil.which("nvcc"):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "nlsolver_b200", "csrc")], check=True)
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "liboracle.so")):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import binding
    return binding.oracle()


@pytest.fixture(scope="session")
def ref_lib():
    """The unmodified reference behind oracle/ref_harness.cpp; skipped where it was never built."""
    from oracle import binding
    lib = binding.reference()
    if lib is None:
        pytest.skip("oracle/_ref/libnls_ref.so not built (reference tree absent)")
    return lib
