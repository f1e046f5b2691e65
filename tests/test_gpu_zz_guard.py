"""Runs last (file name): with NLS_B200_GUARD=1 in the environment every device buffer of the library sits between two
NaN-filled guard zones that are verified when its handle is destroyed — `NLS_B200_GUARD=1 pytest -m gpu` turns the whole
parity suite into an out-of-bounds check (stores are counted here, stray loads poison results and fail parity).
Without the variable this only checks that the counter is wired."""
import os

import numpy as np
import pytest

import nlsolver_b200 as nb

pytestmark = pytest.mark.gpu


def test_guard_zones_intact_after_the_suite():
    ctx = nb.Context(0)
    pop = nb.DEPopulation(ctx, nb.de_cfg(objective=nb.SPHERE, pop_size=67, dim=5, eps=0.0, max_iter=1 << 40,
                                         best_val_no_change=1 << 40, seed=1), np.full(5, 3.0))
    pop.step(3)
    pop.sync()
    pop.close()
    ctx.close()
    assert nb.lib().nls_debug_guard_violations() == 0, "a kernel stored outside its device buffers"
    if os.environ.get("NLS_B200_GUARD") == "1":
        print("guard mode on: all guard zones intact")
