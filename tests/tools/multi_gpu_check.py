"""Multi-GPU parity check over NCCL (run under torchrun on N GPUs of one box):
  * ShardedPSO over N ranks == the single-GPU swarm (bit for bit) == the oracle (1e-12);
  * IslandDE over N ranks with ring migration == the harness-level restatement on the oracle steppers;
  * ShardedSANN over N ranks == the oracle batch (slices by global chain id, batch best by one all-gather).
Prints one line per check on rank 0; exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nlsolver_b200 as nb  # noqa: E402
from nlsolver_b200 import distributed as D  # noqa: E402
from oracle import binding as B  # noqa: E402
from tests.cpu_engines import oracle_de_cfg  # noqa: E402


def main():
    rank, world = D.init_from_env("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ok = True

    # ---- sharded accelerated PSO on Ackley (config-3 shape, small) ----
    P, d, gens = 4096 + 3, 256, 8
    up = np.full(d, 32.768)
    kw = dict(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, n_particles=P, dim=d, eps=0.0, max_iter=1 << 40,
              best_val_no_change=1 << 40, seed=123)
    sw = D.ShardedPSO(nb.pso_cfg(**kw), -up, up, device=local)
    sw.step(gens)
    st = sw.sync()
    best = sw.best()
    b, e = D.slice_bounds(P, world, rank)
    pos = torch.zeros(P, d, dtype=torch.float64, device=f"cuda:{local}")
    pos[b:e] = torch.from_numpy(sw.engine.swarm.positions()).to(pos.device)
    dist.all_reduce(pos)
    sw.close()
    if rank == 0:
        ctx = nb.Context(local)
        whole = nb.PSOSwarm(ctx, nb.pso_cfg(**kw), -up, up)
        whole.step(gens)
        ws = whole.sync()
        same = all(st[k] == ws[k] for k in ("f_value", "iterations", "function_calls", "best_index", "val_no_change"))
        same &= np.array_equal(best, whole.best()) and np.array_equal(pos.cpu().numpy(), whole.positions())
        so, ao = B.pso_run(B.oracle(), B.pso_cfg(**dict(kw, max_iter=gens)), -up, up)
        rel = np.max(np.abs(whole.positions() - ao["positions"]) / np.max(np.abs(ao["positions"]), axis=1, keepdims=True))
        print(f"sharded PSO x{world}: identical to single-GPU swarm = {same}; vs oracle max rel {rel:.2e}, "
              f"best_index {ws['best_index']} == {so['best_index']}", flush=True)
        ok &= same and rel < 1e-12 and ws["best_index"] == so["best_index"]
        whole.close()
        ctx.close()

    # ---- the same swarm with the fused peer-memory exchange (no NCCL in the loop) ----
    sw = D.ShardedPSO(nb.pso_cfg(**kw), -up, up, device=local, exchange="peer")
    sw.step(gens)
    fst = sw.sync()
    fbest = sw.best()
    same = all(fst[k] == st[k] for k in ("f_value", "iterations", "function_calls", "best_index", "val_no_change"))
    same &= np.array_equal(fbest, best)
    flag = torch.tensor([1 if same else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    sw.close()
    if rank == 0:
        print(f"sharded PSO x{world}, fused peer exchange: identical to the NCCL path = {bool(flag.item())}", flush=True)
    ok &= bool(flag.item())

    # ---- island DE with ring migration ----
    Pi, di, every, k, gens = 512, 24, 3, 8, 10
    dkw = dict(objective=nb.ROSENBROCK, strategy=nb.DE_BEST, pop_size=Pi, dim=di, eps=0.0, max_iter=1 << 40,
               best_val_no_change=1 << 40, seed=77)
    x0 = np.full(di, 4.096)
    steppers = [B.DEStepper(oracle_de_cfg(nb.de_cfg(**dict(dkw, agent_offset=r * Pi))), x0) for r in range(world)]
    for g in range(1, gens + 1):
        for s in steppers:
            s.advance(1)
        if D.migration_due(g, every) and world > 1:
            out = [s.export_top(k) for s in steppers]
            for r, s in enumerate(steppers):
                s.import_migrants(*out[D.ring_neighbors(r, world)[1]])
    want = [s.report() for s in steppers]
    gbest = min(range(world), key=lambda r: (want[r][0]["f_value"], r))
    # "nccl": export kernel + all-gather after every generation; "peer": the commit kernel stores the island's record
    # into every peer's window over NVLink, generations between two migrations are one call
    for exchange in ("nccl", "peer"):
        isl = D.IslandDE(nb.de_cfg(**dkw), x0, device=local, migrate_every=every, migrants=k, exchange=exchange)
        isl.step(gens)
        ist = isl.sync()
        rows = isl.engine.pop.population()
        grow = isl.global_best_row()
        mine = (np.array_equal(rows, want[rank][1]["rows"]) and ist["f_value"] == want[rank][0]["f_value"]
                and ist["global_best_rank"] == gbest and ist["global_best_value"] == want[gbest][0]["f_value"]
                and np.array_equal(grow, want[gbest][1]["x_best"]))
        flag = torch.tensor([1 if mine else 0], device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"island DE x{world} ({exchange} exchange): every island and the global best equal the restatement = "
                  f"{bool(flag.item())}", flush=True)
        ok &= bool(flag.item())
        isl.close()

    # ---- SANN chains sharded by global chain id: no exchange while the chains run, one all-gather for the batch best ----
    n, ds, it = 1000 + 3, 24, 30
    skw = dict(objective=nb.RASTRIGIN, n_chains=n, dim=ds, max_iter=it, seed=99)
    xs = np.random.default_rng(4).uniform(-3, 3, size=(n, ds))
    job = D.ShardedSANN(nb.sann_cfg(**skw), xs, device=local)
    job.run()
    gst, grow = job.global_best()
    part = job.engine.chains()
    so, ao = B.sann_run(B.oracle(), B.sann_cfg(**skw), xs)
    rel = np.max(np.abs(part["x_best"] - ao["x_best"][job.begin:job.end])
                 / np.max(np.abs(ao["x_best"][job.begin:job.end]), axis=1, keepdims=True))
    mine = (rel < 1e-12 and np.array_equal(part["n_accepted"], ao["n_accepted"][job.begin:job.end])
            and gst["best_index"] == so["best_index"] and abs(gst["f_value"] - so["f_value"]) <= 1e-12 * abs(so["f_value"])
            and gst["function_calls"] == so["function_calls"]
            and np.max(np.abs(grow - ao["x_best"][so["best_index"]])) <= 1e-12 * np.max(np.abs(grow)))
    flag = torch.tensor([1 if mine else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sharded SANN x{world}: every slice and the batch best equal the oracle = {bool(flag.item())}", flush=True)
    ok &= bool(flag.item())
    job.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
