"""Throughput of a SANN chain batch: Rastrigin, d = 64, fp64, reference defaults
(temperature_iter 10, temperature_max 10), chains sharded over the ranks by global chain id (weak scaling: --per-gpu
chains on every rank, no collective while the chains run).

  python tools/bench_sann.py [--per-gpu N] [--dim D] [--steps K]            one GPU
  torchrun --nproc-per-node N ... tools/bench_sann.py                       N GPUs

Prints one JSON line on rank 0: chain-evaluations/s of the whole job (objective calls per second, the reference's
f_evals), CUDA-event time (max over ranks), and — at N = 1 — the UNMODIFIED reference (oracle/_ref) timed on one host
core on a bounded sample of the same workload."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nlsolver_b200 as nb  # noqa: E402
from nlsolver_b200 import distributed as D  # noqa: E402


def cpu_reference(objective, dim, dtype):
    """The reference's own SANN (one chain after the other, its xorshift generator) on one core: ~10 s of work."""
    from oracle import binding as B   # the checker, timed as the baseline — never on the product path
    ref = B.reference()
    if ref is None:
        return None
    n, it = 64, max(int(1.1e8 / (dim * 64 * 9)), 10)   # ~1e7 coordinate updates per second on one core
    cfg = B.sann_cfg(dtype=dtype, objective=objective, n_chains=n, dim=dim, max_iter=it, rng_mode=B.RNG_XORSHIFT)
    x0 = np.full(dim, 2.5, dtype=B.np_dtype(dtype))
    sec, st = C.c_double(), B.Status()
    assert ref.ref_sann_time(C.byref(cfg), x0.ctypes.data, C.byref(sec), C.byref(st)) == 0
    return {"value": st.function_calls / sec.value, "unit": "chain-evals/s", "cores": 1, "kind": "reference",
            "sample": f"{n} chains x {it} iterations x 9 candidates, d={dim}, nlsolver::SANN + xorshift, {sec.value:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu", type=int, default=1 << 20)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--steps", type=int, default=90, help="candidates per chain in the timed region")
    ap.add_argument("--warmup", type=int, default=9)
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--objective", type=int, default=nb.RASTRIGIN)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        D.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dtype = nb.F64 if args.dtype == "f64" else nb.F32
    n, d = args.per_gpu * world, args.dim
    cfg = nb.sann_cfg(dtype=dtype, objective=args.objective, n_chains=n, dim=d, max_iter=1 << 30, temperature_iter=10,
                      temperature_max=10.0, seed=0x7c26ca28fb68bc1b)
    job = D.ShardedSANN(cfg, np.full(d, 2.5), device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    job.step(args.warmup)
    job.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(job.stream)
    job.step(args.steps)
    e1.record(job.stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st, row = job.global_best()
    job.close()
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        line = {"workload": f"SANN chains, objective {args.objective}, d={d}, {args.per_gpu} chains per GPU x {world} GPU(s), "
                            f"{args.dtype}, temperature_iter 10, temperature_max 10",
                "metric": "chain-evals/s", "value": n * args.steps / sec, "n_gpus": world, "steps": args.steps,
                "ms_per_candidate": sec * 1e3 / args.steps, "coordinate_updates_per_sec": n * args.steps * d / sec,
                "scaling": "weak", "f_value": st["f_value"], "best_chain": st["best_index"]}
        if world == 1 and not args.skip_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(args.objective, d, dtype)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
