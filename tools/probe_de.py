"""Per-kernel device time of the DE generation (K2 generation pass / K2r repair / K3 commit + reduce) against the
acceptance rate, for any shape:

    python tools/probe_de.py --pop 1048576 --dim 1000 --objective sphere [--dtype f32] [--strategy best] [--blocks 6]

Prints one line per block of 10 generations and a final JSON line (the last block) with the algorithmic bandwidth
((4 + a) d s + (1 + a) s bytes per agent-generation for random, (3 + a) d s + (1 + a) s for best)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nlsolver_b200 as nb  # noqa: E402

OBJ = {"sphere": nb.SPHERE, "rosenbrock": nb.ROSENBROCK, "rastrigin": nb.RASTRIGIN, "ackley": nb.ACKLEY}


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pop", type=int, required=True)
    ap.add_argument("--dim", type=int, required=True)
    ap.add_argument("--objective", default="sphere", choices=sorted(OBJ))
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--strategy", default="random", choices=["random", "best"])
    ap.add_argument("--F", type=float, default=0.8)
    ap.add_argument("--CR", type=float, default=0.9)
    ap.add_argument("--x0", type=float, default=10.24)
    ap.add_argument("--blocks", type=int, default=6)
    ap.add_argument("--gens", type=int, default=10, help="generations per block")
    ap.add_argument("--tag", default=os.path.basename(os.environ.get("NLS_B200_LIB", "shipped")))
    ap.add_argument("--peak", type=float, default=6550.1)
    args = ap.parse_args()
    P, d = args.pop, args.dim
    es = 8 if args.dtype == "f64" else 4
    stream = torch.cuda.Stream()
    ctx = nb.Context(0, stream.cuda_stream)
    pop = nb.DEPopulation(ctx, nb.de_cfg(dtype=nb.F64 if es == 8 else nb.F32, objective=OBJ[args.objective],
                                         strategy=nb.DE_RANDOM if args.strategy == "random" else nb.DE_BEST,
                                         pop_size=P, dim=d, differential_weight=args.F, crossover_prob=args.CR,
                                         eps=0.0, max_iter=1 << 40, best_val_no_change=1 << 40, seed=1),
                          np.full(d, args.x0))
    pop.enable_kernel_timing(True)
    prev = pop.sync()
    last = None
    for block in range(args.blocks):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pop.step(args.gens)
        e1.record(stream)
        mhz, watts = sm_clock()      # sampled while the block is (most likely) still running
        st = pop.sync()
        ms, n = pop.kernel_times()
        g = args.gens
        acc = (st["accepted_total"] - prev["accepted_total"]) / (g * P)
        rer = (st["repair_reruns"] - prev["repair_reruns"]) / (g * P)
        rounds = (st["repair_rounds"] - prev["repair_rounds"]) / g
        rows = (4 if args.strategy == "random" else 3) + acc
        alg = P * (rows * d * es + (1 + acc) * es)
        total = e0.elapsed_time(e1) / g
        last = {"pop": P, "dim": d, "dtype": args.dtype, "objective": args.objective, "strategy": args.strategy,
                "generations": [block * g + 1, block * g + g], "accepted_fraction": acc, "repair_rerun_fraction": rer,
                "repair_rounds": rounds, "k2_ms": ms[0] / n, "k2r_ms": ms[1] / n, "k3_ms": ms[2] / n,
                "ms_per_generation": total, "algorithmic_bytes": alg,
                "achieved_GBps_generation": alg / (total * 1e-3) / 1e9,
                "achieved_GBps_k2": alg / (ms[0] / n * 1e-3) / 1e9,
                "frac_of_measured_hbm_generation": alg / (total * 1e-3) / 1e9 / args.peak,
                "f_value": st["f_value"]}
        print(f"[{args.tag}] gens {block*g+1:3d}-{block*g+g:3d}: accepted {acc:6.3f} rerun {rer:6.3f} rounds {rounds:5.1f}  "
              f"K2 {ms[0]/n:7.3f} K2r {ms[1]/n:7.3f} K3 {ms[2]/n:6.3f} total {total:7.3f} ms  "
              f"{last['achieved_GBps_generation']:6.0f} GB/s ({last['frac_of_measured_hbm_generation']:.2f})  "
              f"{mhz} MHz {watts} W  f={st['f_value']:.5g}",
              flush=True)
        prev = st
    print(json.dumps(last), flush=True)
    pop.close()
    ctx.close()


if __name__ == "__main__":
    main()
