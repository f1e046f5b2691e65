"""Measurements of BASELINE.json configs[2] and configs[3] (and single-GPU shards of them) — bench.py covers configs[1].

  torchrun ... tools/bench_configs.py --config 3 [--steps K]   accelerated PSO, Ackley d=256, 2^24 particles sharded over
                                                              the ranks (2^21 per GPU at 8; per-GPU share kept at 2^21
                                                              for fewer ranks: weak scaling), per-generation exchange
  torchrun ... tools/bench_configs.py --config 4 [--steps K]   island DE-best, Rosenbrock d=4096, 2^21 agents per island,
                                                              ring migration of 64 rows every 10 generations
Prints one JSON line on rank 0: agent-evals/s (whole job), ms per generation (max over ranks, CUDA events)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nlsolver_b200 as nb  # noqa: E402
from nlsolver_b200 import distributed as D  # noqa: E402

NEVER = 1 << 62


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 4, 5])
    ap.add_argument("--solver", default="de", choices=["de", "pso-accelerated", "pso-vanilla"], help="config 5")
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"], help="config 5")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--per-gpu", type=int, default=1 << 21)
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"], help="config 3: record exchange path")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        D.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == 3:
        d, P = 256, args.per_gpu * world
        up = np.full(d, 32.768)
        cfg = nb.pso_cfg(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, n_particles=P, dim=d, inertia=0.8,
                         cognitive_coef=1.8, social_coef=1.8, eps=0.0, max_iter=NEVER, best_val_no_change=NEVER,
                         seed=0x7c26ca28fb68bc1b)
        job = D.ShardedPSO(cfg, -up, up, device=local, exchange=args.exchange)
        units, name = P, f"PSO-accelerated Ackley d={d}, {P} particles over {world} GPU(s), fp64, {args.exchange} exchange every generation"
        alg_bytes = 2 * d * 8 + 2 * 8
    elif args.config == 5:
        # population sweep point: d = 64, Sphere; PSO is one global swarm sharded over the ranks, DE one island per rank
        d = 64
        dtype = nb.F64 if args.dtype == "f64" else nb.F32
        es = 8 if args.dtype == "f64" else 4
        if args.solver == "de":
            cfg = nb.de_cfg(dtype=dtype, objective=nb.SPHERE, pop_size=args.per_gpu, dim=d, eps=0.0, max_iter=NEVER,
                            best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b)
            job = D.IslandDE(cfg, np.full(d, 10.24), device=local, migrate_every=10, migrants=64)
            alg_bytes = 4 * d * es
        else:
            ptype = nb.PSO_ACCELERATED if args.solver == "pso-accelerated" else nb.PSO_VANILLA
            up = np.full(d, 10.24)
            cfg = nb.pso_cfg(dtype=dtype, objective=nb.SPHERE, pso_type=ptype, n_particles=args.per_gpu * world, dim=d,
                             eps=0.0, max_iter=NEVER, best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b,
                             flags=nb.FLAG_SOCIAL_INDEX_J)
            job = D.ShardedPSO(cfg, -up, up, device=local, exchange=args.exchange)
            alg_bytes = (2 if ptype == nb.PSO_ACCELERATED else 4) * d * es
        units = args.per_gpu * world
        name = f"sweep point: {args.solver} Sphere d={d} {args.dtype}, {args.per_gpu} per GPU x {world} GPU(s)"
    else:
        d, P = 4096, args.per_gpu
        cfg = nb.de_cfg(objective=nb.ROSENBROCK, strategy=nb.DE_BEST, pop_size=P, dim=d, eps=0.0, max_iter=NEVER,
                        best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b)
        job = D.IslandDE(cfg, np.full(d, 4.096), device=local, migrate_every=10, migrants=64)
        units, name = P * world, f"island DE-best Rosenbrock d={d}, {world} island(s) x {P} agents, fp64, ring migration 64 rows / 10 generations"
        alg_bytes = 3 * d * 8 + 8
    stream = job.stream
    job.step(args.warmup)
    st0 = job.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    job.step(args.steps)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st1 = job.sync()
    assert st1["iterations"] - st0["iterations"] == args.steps
    job.close()
    if rank == 0:
        per_gen = float(ms.item()) / args.steps
        print(json.dumps({"config": args.config, "workload": name, "n_gpus": world, "steps": args.steps,
                          "ms_per_generation": per_gen, "agent_evals_per_sec": units / (per_gen * 1e-3),
                          "algorithmic_GBps_per_gpu": alg_bytes * (units / world) / (per_gen * 1e-3) / 1e9,
                          "f_value": st1.get("global_best_value", st1["f_value"])}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
