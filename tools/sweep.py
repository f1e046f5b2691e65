"""Population sweep (BASELINE.json configs[4]): P = 2^10 .. 2^24 at d = 64, Sphere, {DE-random, PSO-accelerated,
PSO-vanilla (corrected social index)} x {fp32, fp64} on one GPU; prints a markdown table (agent-evals/s, ms per
generation, algorithmic GB/s).  Small populations are launch-latency-bound and are reported as such."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nlsolver_b200 as nb  # noqa: E402

d = 64
stream = torch.cuda.Stream()
ctx = nb.Context(0, stream.cuda_stream)
NEVER = 1 << 40


def timed(step, sync, gens):
    step(3)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    step(gens)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / gens


rows = []
pops = [1 << e for e in range(10, 25, 2)] + ([1 << 25, 1 << 26] if "--big" in sys.argv else [])
for P in pops:
    gens = 200 if P <= 1 << 14 else 50 if P <= 1 << 20 else 10
    for dtype, name, s in ((nb.F64, "fp64", 8), (nb.F32, "fp32", 4)):
        pop = nb.DEPopulation(ctx, nb.de_cfg(dtype=dtype, objective=nb.SPHERE, pop_size=P, dim=d, eps=0.0,
                                             max_iter=NEVER, best_val_no_change=NEVER, seed=1), np.full(d, 10.24))
        ms = timed(pop.step, pop.sync, gens)
        pop.close()
        rows.append(("DE-random", name, P, ms, P / ms * 1e3, 4 * d * s * P / ms / 1e6))
        for ptype, pname, mult in ((nb.PSO_ACCELERATED, "PSO-accelerated", 2), (nb.PSO_VANILLA, "PSO-vanilla[j]", 4)):
            up = np.full(d, 10.24)
            sw = nb.PSOSwarm(ctx, nb.pso_cfg(dtype=dtype, objective=nb.SPHERE, pso_type=ptype, n_particles=P, dim=d,
                                             eps=0.0, max_iter=NEVER, best_val_no_change=NEVER, seed=1,
                                             flags=nb.FLAG_SOCIAL_INDEX_J), -up, up)
            ms = timed(sw.step, sw.sync, gens)
            sw.close()
            rows.append((pname, name, P, ms, P / ms * 1e3, mult * d * s * P / ms / 1e6))
    ctx.trim()
print("| solver | dtype | P | ms / generation | agent-evals/s | algorithmic GB/s | of measured HBM |")
print("|---|---|---|---|---|---|---|")
for solver, name, P, ms, ev, gbs in rows:
    print(f"| {solver} | {name} | 2^{P.bit_length() - 1} | {ms:.4f} | {ev:.3e} | {gbs:.0f} | {gbs / 6550.1:.3f} |")
