"""Multi-GPU parity where `pytest -m gpu` runs it: spawns tests/tools/multi_gpu_check.py under torch.distributed.run on
min(2, visible GPUs) ranks — sharded PSO (NCCL and fused peer exchange) == the single-GPU swarm bit for bit, island DE
with ring migration == the restatement (nlsolver.h:2449-2472, 2716-2741 per island / shard), sharded SANN == the oracle
batch.  On a one-GPU box the multi-rank logic is covered by tests/test_gpu_islands.py (two islands on one GPU), the
single-process device group (tests/test_gpu_device_group.py) and the world_size-2 gloo tests."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_multi_gpu_check_under_torchrun():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the NCCL / IPC multi-rank path needs two devices")
    ranks = min(n, 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    for needle in ("identical to single-GPU swarm = True", "identical to the NCCL path = True",
                   "(nccl exchange): every island and the global best equal the restatement = True",
                   "(peer exchange): every island and the global best equal the restatement = True",
                   "equal the oracle = True"):
        assert needle in out.stdout, out.stdout[-4000:]
