"""Device groups (ONE process, N GPUs, peer memory; nls_group_* in include/nls_b200.h): per-generation time of a
sharded accelerated-PSO swarm on Ackley d = 256 (BASELINE configs[2] shape) at a small and a large shard size, and of DE
islands with ring migration.  Host wall clock around step + sync (the generations of all devices are enqueued
asynchronously; sync waits for every device).

    python tools/bench_group.py [--gpus N]          -> one JSON line
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nlsolver_b200 as nb  # noqa: E402

NEVER = 1 << 40


def timed(job, gens, reps=3):
    job.step(gens)
    job.sync()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        job.step(gens)
        job.sync()
        sec = time.perf_counter() - t0
        best = sec if best is None else min(best, sec)
    return 1e6 * best / gens


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    args = ap.parse_args()
    n = args.gpus
    group = nb.DeviceGroup(n)
    out = {"n_gpus": n, "sharded_pso_accelerated_ackley_d256": [], "de_islands": []}
    d = 256
    up = np.full(d, 32.768)
    for per_gpu, gens in ((1 << 14, 256), (1 << 21, 16)):
        cfg = nb.pso_cfg(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, n_particles=per_gpu * n, dim=d, eps=0.0,
                         max_iter=NEVER, best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b)
        sw = nb.ShardedSwarm(group, cfg, -up, up)
        us = timed(sw, gens)
        st = sw.sync()
        sw.close()
        out["sharded_pso_accelerated_ackley_d256"].append(
            {"particles_per_gpu": per_gpu, "us_per_generation": us, "agent_evals_per_sec": per_gpu * n / (us * 1e-6),
             "f_value": st["f_value"]})
    for P, dd, gens in ((1 << 12, 64, 200), (1 << 20, 1000, 20)):
        cfg = nb.de_cfg(objective=nb.RASTRIGIN, pop_size=P, dim=dd, eps=0.0, max_iter=NEVER, best_val_no_change=NEVER,
                        seed=0x7c26ca28fb68bc1b)
        isl = nb.DEIslands(group, cfg, np.full(dd, 10.24), migrate_every=10, migrants=64)
        us = timed(isl, gens)
        st = isl.sync()
        isl.close()
        out["de_islands"].append({"agents_per_island": P, "dim": dd, "us_per_generation": us,
                                  "agent_evals_per_sec": P * n / (us * 1e-6), "f_value": st["f_value"]})
    group.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
