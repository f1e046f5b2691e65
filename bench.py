#!/usr/bin/env python
"""bench.py — agent-evaluations per second of the DE generation loop on B200 (BASELINE.json metric).

Workload at N = 1: BASELINE.json configs[1] — DE, random recombination, Rastrigin N-D, d = 1000, population 2^20,
fp64, CR = 0.9, F = 0.8, x0[j] = 10.24 (agents start uniform in [-5.12, 5.12]); stop rules disabled (eps = 0,
best_val_no_change = inf) so only the step count ends the run.  A "step" is one generation: one pass
of the hot path over the whole population = P agent evaluations.  At N > 1 DE does not shard a single population
(SURVEY.md §8e), so every rank runs one such island (weak scaling) with the per-generation best all-gather and the
ring migration every 10 generations; `value` is the whole-job agent-evaluations per second.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's own CPU path, all host threads

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks.  The population (2 x 8.4 GB) is far larger than the 126 MB L2, so no flush is needed between steps.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POP, DIM = 1 << 20, 1000
X0 = 10.24
CR, F = 0.9, 0.8
MIGRATE_EVERY, MIGRANTS = 10, 64
NEVER = 1 << 62
METRIC, UNIT = "agent_evals_per_sec", "agent-evals/s"
WORKLOAD = "DE-random Rastrigin d=1000 P=1048576 fp64 (BASELINE.json configs[1])"


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture, or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(path))[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


# ------------------------------------------------------------------ clocks during the timed region ---------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        self.samples, self.mask, self.max_mhz, self.power = [], 0, None, []
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------ the reference's CPU path ----------------------
def cpu_runner():
    """(timing function, kind): oracle/_ref (the unmodified reference, compiled from its own sources) when present,
    else the oracle port."""
    from oracle import binding as B
    ref = B.reference()
    if ref is not None:
        def run(pop, gens, dim=DIM):
            cfg = B.de_cfg(objective=B.RASTRIGIN, strategy=B.DE_RANDOM, pop_size=pop, dim=dim, crossover_prob=CR,
                           differential_weight=F, eps=0.0, max_iter=gens, best_val_no_change=NEVER)
            x0 = np.full(dim, X0)
            sec, st = C.c_double(), B.Status()
            rc = ref.ref_de_time(C.byref(cfg), x0.ctypes.data, C.byref(sec), C.byref(st))
            assert rc == 0 and st.function_calls == pop * (gens + 1)
            return st.function_calls, sec.value
        return run, "reference"
    lib = B.oracle()

    def run(pop, gens, dim=DIM):
        cfg = B.de_cfg(objective=B.RASTRIGIN, strategy=B.DE_RANDOM, pop_size=pop, dim=dim, crossover_prob=CR,
                       differential_weight=F, eps=0.0, max_iter=gens, best_val_no_change=NEVER,
                       rng_mode=B.RNG_XORSHIFT)
        x0 = np.full(dim, X0)
        st = B.Status()
        t0 = time.perf_counter()
        rc = lib.oracle_de_run(C.byref(cfg), x0.ctypes.data, None, C.byref(st))
        sec = time.perf_counter() - t0
        assert rc == 0
        return st.function_calls, sec
    return run, "port"


def cpu_baseline_single_thread():
    run, kind = cpu_runner()
    pop, gens = 8192, 24          # ~10 s of single-thread CPU work at ~50 us per agent evaluation
    run(256, 1)
    calls, sec = run(pop, gens)
    return {"value": calls / sec, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"same workload (DE-random Rastrigin d={DIM} fp64, xorshift<double>) at population {pop}, "
                      f"{gens} generations + init = {calls} evaluations in {sec:.2f} s on one host thread; per-agent "
                      "CPU cost is population-independent beyond cache size"}


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation on all host threads (independent replicas: the reference
    is single-threaded by design, README.md:143-145).  Rank 0 alone runs; other ranks exit."""
    if rank != 0:
        return
    run, kind = cpu_runner()
    threads = os.cpu_count() or 1
    pop, gens = 1024, 3           # one step per thread = 4096 evaluations, ~0.2 s

    def step():
        out = [None] * threads

        def work(k):
            out[k] = run(pop, gens)
        ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return sum(o[0] for o in out), time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    calls, sec = 0, 0.0
    for _ in range(args.steps):
        c, s = step()
        calls += c
        sec += s
    value = calls / sec
    sample = (f"each step = {threads} independent replicas (one per host thread) of DE-random Rastrigin d={DIM} fp64 at "
              f"population {pop}, {gens} generations + init; {args.steps} steps, {calls} evaluations in {sec:.2f} s")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "population_per_gpu": POP, "dim": DIM,
                       "note": "CPU arm runs a bounded sample of the workload (see cpu_baseline.sample)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ this repo's CUDA path -------------------------
def run_b200_arm(args, rank, world):
    import torch
    import torch.distributed as dist

    import nlsolver_b200 as nb
    from nlsolver_b200 import _lib as L
    from nlsolver_b200.distributed import IslandDE

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(dev)
    pop, dim, K, W = args.pop, args.dim, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = nb.de_cfg(dtype=nb.F64, objective=nb.RASTRIGIN, strategy=nb.DE_RANDOM, pop_size=pop, dim=dim,
                    crossover_prob=CR, differential_weight=F, eps=0.0, max_iter=NEVER, best_val_no_change=NEVER,
                    seed=0x7c26ca28fb68bc1b)      # first raw output of the reference's default-seeded generator
    x0 = np.full(dim, X0)
    island = IslandDE(cfg, x0, device=local, migrate_every=MIGRATE_EVERY, migrants=MIGRANTS, stream=stream)
    island.step(W)
    st0 = island.sync()
    island.engine.pop.enable_kernel_timing(True)
    launches0 = island.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0.record(stream)
    island.step(K)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    kernel_ms, timed_gens = island.engine.pop.kernel_times()
    island.engine.pop.enable_kernel_timing(False)
    st1 = island.sync()
    assert st1["iterations"] - st0["iterations"] == K and timed_gens == K
    launches = island.launches - launches0
    accepted = st1["accepted_total"] - st0["accepted_total"]
    reruns = st1["repair_reruns"] - st0["repair_reruns"]
    island.close()

    value = world * pop * K / (total_ms * 1e-3)

    # ---- end to end through the public API: DE(...).minimize(x) with host buffers (nls_de_solve) -----------------
    # One call = allocate, H2D of x0, init + K generations, D2H of the best row and the status.
    class TwoDraws:   # stands in for the user's RNG: the header / mirror takes two draws for the tape seed
        def __init__(self):
            self.v = [0.40764453281267443, 0.82621863718638611]

        def __call__(self):
            return self.v.pop(0)
    ctx = nb.Context(local, stream.cuda_stream)
    # warm-up call of the same shape (3 generations): first-use costs (module load, cudaMalloc of 2 x 8.4 GB, which
    # the context then keeps for the next solve) are not part of the steady-state call a user repeats
    nb.DE(nb.Rastrigin, TwoDraws(), CR, F, 0.0, pop, 3, NEVER, ctx=ctx).minimize(np.full(dim, X0))
    solver = nb.DE(nb.Rastrigin, TwoDraws(), CR, F, 0.0, pop, K, NEVER, ctx=ctx)
    x = np.full(dim, X0)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record(stream)
    status = solver.minimize(x)
    s1.record(stream)
    barrier()
    e2e_ms = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    assert status.iteration == K and status.function_calls_used == pop * (K + 1)
    e2e_value = world * status.function_calls_used / (float(e2e_ms.item()) * 1e-3)
    ctx.close()

    if rank != 0:
        return
    peak, peak_src = measured_hbm_peak()
    a = accepted / float(pop * K)
    alg_bytes = pop * ((4 + a) * dim * 8 + (1 + a) * 8)       # (4+a)*d*s + (1+a)*s per agent
    k2_ms = kernel_ms[0] / K
    achieved = alg_bytes / (k2_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD if (pop, dim) == (POP, DIM) else f"DE-random Rastrigin d={dim} P={pop} fp64",
                   "population_per_gpu": pop, "dim": dim, "crossover_prob": CR, "differential_weight": F,
                   "islands": world, "migrate_every": MIGRATE_EVERY, "migrants": MIGRANTS,
                   "accepted_fraction": a, "repair_rerun_fraction": reruns / float(pop * K),
                   "l2": "inputs larger than L2: two row buffers of %.1f GB vs 126 MB, no flush" % (pop * dim * 8 / 1e9)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": dim * 8 / K,
                "d2h_bytes_per_step": (dim * 8 + C.sizeof(L.Status)) / K,
                "call": "nlsolver_b200.DE(...).minimize(x) -> nls_de_solve: H2D x0 + init + K generations + D2H best "
                        "row/status, after one warm-up call of the same shape (device buffers are cached by the "
                        "context); the population is generated on the device, as in the reference"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "de_generation_kernel<double, Rastrigin>", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "traffic": committed_traffic("de_generation_kernel"),
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k2_ms,
                     "step_share": {"generation": kernel_ms[0] / total_ms, "repair": kernel_ms[1] / total_ms,
                                    "commit_reduce": kernel_ms[2] / total_ms}},
    }
    if world == 1 and not args.skip_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single_thread()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pop", type=int, default=POP, help="population per GPU (default: the BASELINE configuration)")
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs only (ncu): skip the CPU leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world > 1:
        from nlsolver_b200.distributed import init_from_env
        init_from_env("nccl")
    try:
        run_b200_arm(args, rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
