"""The numpy restatement of the draw tape used by the full-size GPU test (tests/test_gpu_full_size.py) against the
oracle's decisions: the checker is checked on the CPU before it is trusted on the GPU box."""
import numpy as np

from oracle import binding as B
from tests.test_gpu_full_size import check_tape_against_oracle, host_decisions, tape_draws, unit64


def test_numpy_tape_equals_oracle_tape():
    check_tape_against_oracle()


def test_host_decisions_equal_oracle_decisions():
    P, d, seed, gens = 300, 17, 99, 3
    so, ao = B.de_run(B.oracle(), B.de_cfg(objective=B.SPHERE, pop_size=P, dim=d, eps=0.0, max_iter=gens,
                                           best_val_no_change=1 << 40, seed=seed), np.full(d, 10.24), masks=True)
    keys, donors, rej, dim_idx = host_decisions(seed, gens, np.arange(P), P, d, np.arange(P))
    assert np.array_equal(donors, ao["donors"]) and np.array_equal(rej, ao["rejects"])
    assert np.array_equal(dim_idx, ao["dim_idx"])
    jj = np.arange(d, dtype=np.uint64)
    for n in range(P):
        u = unit64(tape_draws(keys[n:n + 1], (np.uint64(4 + rej[n]) + jj)[None, :])[0])
        assert np.array_equal(((u < 0.9) | (jj == np.uint64(dim_idx[n]))).astype(np.uint8), ao["masks"][n])
