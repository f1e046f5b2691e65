"""Where does the multi-GPU step overhead of the island bench come from?  torchrun, N ranks; prints ms/step variants."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nlsolver_b200 as nb  # noqa: E402
from nlsolver_b200 import distributed as D  # noqa: E402

rank, world = D.init_from_env("nccl")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
P, d, K = 1 << 20, 1000, 40
cfg = nb.de_cfg(objective=nb.RASTRIGIN, pop_size=P, dim=d, eps=0.0, max_iter=1 << 60, best_val_no_change=1 << 60, seed=1)
isl = D.IslandDE(cfg, np.full(d, 10.24), device=local, migrate_every=10, migrants=64)
stream = isl.stream


def timed(fn):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    with torch.cuda.stream(stream):
        for g in range(K):
            fn(g)
    e1.record(stream)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


isl.step(5)
res = {}
res["generation only"] = timed(lambda g: isl.engine.step(1))
res["+ export_best"] = timed(lambda g: (isl.engine.step(1), isl.engine.export_best(isl.mine[0])))
res["+ blocking all_gather"] = timed(lambda g: (isl.engine.step(1), isl.engine.export_best(isl.mine[0]), isl.comm.all_gather(isl.all[0], isl.mine[0])))


def with_mig(g):
    isl.engine.step(1)
    if (g + 1) % 10 == 0:
        isl.engine.export_top(isl.k, isl.out_rows, isl.out_scores)
        isl.comm.ring_exchange(isl.out_rows, isl.in_rows)
        isl.comm.ring_exchange(isl.out_scores, isl.in_scores)
        isl.engine.import_migrants(isl.k, isl.in_rows, isl.in_scores)


res["generation + migration/10"] = timed(with_mig)
res["IslandDE.step (all of it)"] = timed(lambda g: isl.step(1))
if rank == 0:
    for k, v in res.items():
        print(f"{k:32s} {v:8.3f} ms/step")
isl.close()
dist.destroy_process_group()
