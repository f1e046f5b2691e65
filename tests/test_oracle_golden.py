"""The oracle against the committed reference fixtures (tests/golden/*.npz) — runs where the reference cannot."""
import os

import numpy as np
import pytest

from oracle import binding as B
from tests.golden_util import golden_files, load_de, load_pso, load_sann


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({8: np.uint64, 4: np.uint32, 1: np.uint8}[a.dtype.itemsize]) if a.dtype.kind == "f" else a


@pytest.mark.parametrize("path", golden_files("de_"), ids=os.path.basename)
def test_de_oracle_reproduces_reference_fixture(oracle_lib, path):
    cfg, x0, z = load_de(path)
    st, a = B.de_run(oracle_lib, cfg, x0, masks=True)
    for k in ("f_value", "iterations", "function_calls", "draws_consumed", "best_index"):
        assert st[k] == z[k].item(), k
    for k in ("x_best", "rows", "scores", "trial_scores", "donors", "dim_idx", "rejects", "accepted", "masks"):
        assert np.array_equal(bits(a[k]), bits(z[k])), k


@pytest.mark.parametrize("path", golden_files("pso_"), ids=os.path.basename)
def test_pso_oracle_reproduces_reference_fixture(oracle_lib, path):
    cfg, up, z = load_pso(path)
    st, a = B.pso_run(oracle_lib, cfg, -up, up)
    for k in ("f_value", "iterations", "function_calls", "draws_consumed", "best_valid"):
        assert st[k] == z[k].item(), k
    for k in ("x_best", "positions", "pbest_values", "last_values"):
        assert np.array_equal(bits(a[k]), bits(z[k])), k


@pytest.mark.parametrize("path", golden_files("sann_"), ids=os.path.basename)
def test_sann_oracle_reproduces_reference_fixture(oracle_lib, path):
    cfg, x0, z = load_sann(path)
    st, a = B.sann_run(oracle_lib, cfg, x0)
    assert st["f_value"] == z["f_value"].item() and st["best_index"] == z["best_index"].item()
    assert st["function_calls"] == z["function_calls_total"].item()
    for k in ("x_best", "f_best", "draws", "iterations", "function_calls"):
        assert np.array_equal(bits(a[k]), bits(z[k])), k


@pytest.mark.parametrize("path", golden_files("nmpso_"), ids=os.path.basename)
def test_nmpso_oracle_reproduces_reference_fixture(oracle_lib, path):
    """NelderMeadPSO (nlsolver.h:3546-3920): the restatement against what the UNMODIFIED reference produced."""
    from tests.golden_util import load_nmpso
    cfg, x0, z = load_nmpso(path)
    st, a = B.nmpso_run(oracle_lib, cfg, x0)
    for k in ("x_best", "f_best", "iterations", "function_calls", "draws"):
        assert np.array_equal(a[k].view(np.uint8), z[k].view(np.uint8)), k
    assert st["best_index"] == z["best_index"].item() and st["f_value"] == z["f_value"].item()
    assert not a["ties"].any()
