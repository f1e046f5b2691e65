#!/usr/bin/env python
"""bench.py — agent-evaluations per second of the DE generation loop on B200 (BASELINE.json metric).

Workload at N = 1: BASELINE.json configs[1] — DE, random recombination, Rastrigin N-D, d = 1000, population 2^20,
fp64, CR = 0.9, F = 0.8, x0[j] = 10.24 (agents start uniform in [-5.12, 5.12]); stop rules disabled (eps = 0,
best_val_no_change = inf) so only the step count ends the run.  A "step" is one generation: one pass
of the hot path over the whole population = P agent evaluations.  At N > 1 DE does not shard a single population
(SURVEY.md §8e), so every rank runs one such island (weak scaling); the island bests are exchanged every generation —
the commit kernel stores the island's record into every peer's exchange window over NVLink — and the 64 best rows go
around the ring every 10 generations; `value` is the whole-job agent-evaluations per second.

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


def committed_traffic(kernel, key="dram_bytes_per_launch"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (or another committed counter) from the committed
    ncu --set full capture of `kernel`, or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(path))[kernel][key]
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
    pop, gens = 8192, 60          # ~12 s of single-thread CPU work at ~23 us per agent evaluation
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
class TwoDraws:
    """Stands in for the user's RNG: the header / Python mirror takes two draws for the tape seed (the first two
    draws of the reference's default-seeded xorshift<double>)."""

    def __init__(self):
        self.v = [0.40764453281267443, 0.82621863718638611]

    def __call__(self):
        return self.v.pop(0)


class Bench:
    """Shared plumbing of the measured sections: device, stream, barrier, CUDA-event timing (max over ranks), clocks."""

    def __init__(self, args, rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.world = args, rank, world
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.stream = torch.cuda.Stream(self.dev)
        self.peak, self.peak_src = measured_hbm_peak()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, flag):
        t = self.torch.tensor([1 if flag else 0], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def timed(self, run):
        """run() enqueues on self.stream; returns (ms, clocks) with barrier + synchronize on both sides."""
        e0, e1 = (self.torch.cuda.Event(enable_timing=True) for _ in range(2))
        sampler = ClockSampler(self.local)
        self.barrier()
        sampler.start()
        e0.record(self.stream)
        run()
        e1.record(self.stream)
        self.barrier()
        clocks = sampler.stop()
        return self.max_over_ranks(e0.elapsed_time(e1)), clocks

    def roof(self, alg_bytes_per_gpu, ms):
        gbs = alg_bytes_per_gpu / (ms * 1e-3) / 1e9
        return {"achieved_GBps_per_gpu": gbs, "frac_of_measured_hbm": gbs / self.peak}


def short_clocks(c):
    return {"sm_mhz": c["sm_mhz"], "reasons": c["reasons"], "power_w_max": c["power_w_max"]}


def bench_cold_call(b, pop, dim, K):
    """First DE(...).minimize(x) of the process on a fresh context: cubin load, cudaMalloc of every buffer, H2D, init,
    K generations, D2H — host wall clock around the call (allocation is host-blocking), max over ranks."""
    import nlsolver_b200 as nb
    ctx = nb.Context(b.local, b.stream.cuda_stream)
    solver = nb.DE(nb.Rastrigin, TwoDraws(), CR, F, 0.0, pop, K, NEVER, ctx=ctx)
    x = np.full(dim, X0)
    b.barrier()
    t0 = time.perf_counter()
    status = solver.minimize(x)
    b.torch.cuda.synchronize()
    sec = b.max_over_ranks(time.perf_counter() - t0)
    ctx.close()
    assert status.iteration == K
    return {"cold_value": b.world * status.function_calls_used / sec, "cold_ms": sec * 1e3,
            "cold_call": "the first minimize() of the process on a fresh context (module load + allocation of every "
                         "device buffer included), host wall clock, same K generations"}


def bench_config2(b):
    """The headline: BASELINE configs[1] as one island per rank (IslandDE), CUDA events, K2 timed by the library."""
    import nlsolver_b200 as nb
    from nlsolver_b200 import _lib as L
    from nlsolver_b200.distributed import IslandDE
    args = b.args
    pop, dim, K, W = args.pop, args.dim, args.steps, args.warmup
    cfg = nb.de_cfg(dtype=nb.F64, objective=nb.RASTRIGIN, strategy=nb.DE_RANDOM, pop_size=pop, dim=dim,
                    crossover_prob=CR, differential_weight=F, eps=0.0, max_iter=NEVER, best_val_no_change=NEVER,
                    seed=0x7c26ca28fb68bc1b)      # first raw output of the reference's default-seeded generator
    island = IslandDE(cfg, np.full(dim, X0), device=b.local, migrate_every=MIGRATE_EVERY, migrants=MIGRANTS,
                      stream=b.stream)
    island.step(W)
    st0 = island.sync()
    island.engine.pop.enable_kernel_timing(True)
    launches0 = island.launches
    total_ms, clocks = b.timed(lambda: island.step(K))
    kernel_ms, timed_gens = island.engine.pop.kernel_times()
    island.engine.pop.enable_kernel_timing(False)
    st1 = island.sync()
    assert st1["iterations"] - st0["iterations"] == K and timed_gens == K
    launches = island.launches - launches0
    accepted = st1["accepted_total"] - st0["accepted_total"]
    reruns = st1["repair_reruns"] - st0["repair_reruns"]
    island.close()
    value = b.world * pop * K / (total_ms * 1e-3)

    # ---- end to end through the public API: DE(...).minimize(x) with host buffers (nls_de_solve) -----------------
    # One call = allocate (or reuse the context's cached buffers), H2D of x0, init + K generations, D2H of the best
    # row and the status.  The warm call is the one a user repeats; the cold one is reported beside it (bench_cold_call).
    ctx = nb.Context(b.local, b.stream.cuda_stream)
    nb.DE(nb.Rastrigin, TwoDraws(), CR, F, 0.0, pop, 3, NEVER, ctx=ctx).minimize(np.full(dim, X0))
    solver = nb.DE(nb.Rastrigin, TwoDraws(), CR, F, 0.0, pop, K, NEVER, ctx=ctx)
    x = np.full(dim, X0)
    box = {}
    e2e_ms, _ = b.timed(lambda: box.update(status=solver.minimize(x)))
    status = box["status"]
    assert status.iteration == K and status.function_calls_used == pop * (K + 1)
    e2e_value = b.world * status.function_calls_used / (e2e_ms * 1e-3)
    ctx.close()

    a = accepted / float(pop * K)
    alg_bytes = pop * ((4 + a) * dim * 8 + (1 + a) * 8)       # (4+a)*d*s + (1+a)*s per agent
    k2_ms = kernel_ms[0] / K
    achieved = alg_bytes / (k2_ms * 1e-3) / 1e9
    return {
        "value": value, "ms_per_step": total_ms / K, "clocks": clocks, "gpu_launches": launches,
        "config": {"workload": WORKLOAD if (pop, dim) == (POP, DIM) else f"DE-random Rastrigin d={dim} P={pop} fp64",
                   "population_per_gpu": pop, "dim": dim, "crossover_prob": CR, "differential_weight": F,
                   "islands": b.world, "migrate_every": MIGRATE_EVERY, "migrants": MIGRANTS,
                   "accepted_fraction": a, "repair_rerun_fraction": reruns / float(pop * K),
                   "l2": "inputs larger than L2: two row buffers of %.1f GB vs 126 MB, no flush" % (pop * dim * 8 / 1e9)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": dim * 8 / K,
                "d2h_bytes_per_step": (dim * 8 + C.sizeof(L.Status)) / K,
                "call": "nlsolver_b200.DE(...).minimize(x) -> nls_de_solve: H2D x0 + init + K generations + D2H best "
                        "row/status, after one warm-up call of the same shape (device buffers are cached by the "
                        "context); the population is generated on the device, as in the reference"},
        "roofline": {"bound": "hbm", "kernel": "de_generation_bulk_kernel<double, Rastrigin, 2, 2> (K2, rows staged by TMA bulk copies)",
                     "achieved": achieved,
                     "peak": b.peak, "unit": "GB/s", "frac": achieved / b.peak, "peak_source": b.peak_src,
                     "traffic": committed_traffic("de_generation_bulk_kernel"),
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k2_ms,
                     "step_share": {"generation": kernel_ms[0] / total_ms, "repair": kernel_ms[1] / total_ms,
                                    "commit_reduce": kernel_ms[2] / total_ms}},
    }


def de_windows(b, cfg, x0, windows, rows_read):
    """Run one DE population through consecutive windows of generations; per window: device ms per generation (CUDA
    events, max over ranks), the library's per-kernel times, measured acceptance / repair statistics and the
    algorithmic bandwidth (rows_read + a rows of d*s bytes + (1 + a) scores per agent-generation)."""
    import nlsolver_b200 as nb
    ctx = nb.Context(b.local, b.stream.cuda_stream)
    pop = nb.DEPopulation(ctx, cfg, x0)
    P, d, es = cfg.pop_size, cfg.dim, 8 if cfg.dtype == nb.F64 else 4
    out = []
    prev = pop.sync()
    for name, gens, timed in windows:
        if not timed:
            pop.step(gens)
            prev = pop.sync()
            continue
        pop.enable_kernel_timing(True)
        ms, clocks = b.timed(lambda: pop.step(gens))
        kms, n = pop.kernel_times()
        pop.enable_kernel_timing(False)
        st = pop.sync()
        assert n == gens and st["iterations"] - prev["iterations"] == gens
        a = (st["accepted_total"] - prev["accepted_total"]) / float(gens * P)
        alg = P * ((rows_read + a) * d * es + (1 + a) * es)
        rec = {"generations": [prev["iterations"] + 1, st["iterations"]], "accepted_fraction": a,
               "repair_rerun_fraction": (st["repair_reruns"] - prev["repair_reruns"]) / float(gens * P),
               "repair_iterations_per_generation": (st["repair_rounds"] - prev["repair_rounds"]) / float(gens),
               "ms_per_generation": ms / gens, "k2_ms": kms[0] / gens, "k2r_ms": kms[1] / gens, "k3_ms": kms[2] / gens,
               "agent_evals_per_sec": b.world * P * gens / (ms * 1e-3),
               "algorithmic_bytes_per_generation": alg, "clocks": short_clocks(clocks)}
        rec.update(b.roof(alg, ms / gens))
        if name:
            rec["window"] = name
        out.append(rec)
        prev = st
    pop.close()
    ctx.close()
    return out


def bench_accepting(b):
    """The regime config 2 never reaches (at d = 1000 and F = 0.8 no trial is ever accepted, so the in-place repair
    idles): the same shape on Sphere with F = 0.2, where a few % of the trials are accepted at first and ~30 % later."""
    import nlsolver_b200 as nb
    pop, dim = b.args.pop, b.args.dim
    cfg = nb.de_cfg(dtype=nb.F64, objective=nb.SPHERE, strategy=nb.DE_RANDOM, pop_size=pop, dim=dim, crossover_prob=CR,
                    differential_weight=0.2, eps=0.0, max_iter=NEVER, best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b)
    wins = de_windows(b, cfg, np.full(dim, X0), [(None, 2, False), ("generations 3-12", 10, True), (None, 13, False),
                                                 ("generations 26-35", 10, True)], rows_read=4)
    # a shape where every fifth trial is accepted (the repair re-evaluates ~30 % of the agents over ~8 iterations)
    cfg32 = nb.de_cfg(dtype=nb.F32, objective=nb.SPHERE, strategy=nb.DE_RANDOM, pop_size=1 << 22, dim=64, crossover_prob=CR,
                      differential_weight=0.3, eps=0.0, max_iter=NEVER, best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b)
    high = de_windows(b, cfg32, np.full(64, X0), [(None, 3, False), ("d=64 fp32 P=2^22 F=0.3", 10, True)], rows_read=4)
    return {"workload": f"DE-random Sphere d={dim} P={pop} fp64, F=0.2 CR={CR} (config-2 shape, trials are accepted)",
            "bytes_formula": "(4 + a) d s + (1 + a) s per agent-generation, a = measured accepted fraction",
            "windows": wins, "high_acceptance": high}


def bench_config4(b):
    """BASELINE configs[3]: island DE-best, Rosenbrock d = 4096, 2^21 agents per island (2 x 64 GiB of rows per GPU),
    ring migration of 64 rows every 10 generations.  One island per rank."""
    import nlsolver_b200 as nb
    from nlsolver_b200.distributed import IslandDE
    d, P = 4096, b.args.config4_pop
    cfg = nb.de_cfg(objective=nb.ROSENBROCK, strategy=nb.DE_BEST, pop_size=P, dim=d, eps=0.0, max_iter=NEVER,
                    best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b)
    job = IslandDE(cfg, np.full(d, 4.096), device=b.local, migrate_every=MIGRATE_EVERY, migrants=MIGRANTS,
                   stream=b.stream)
    job.step(3)
    st0 = job.sync()
    gens = 10
    job.engine.pop.enable_kernel_timing(True)
    ms, clocks = b.timed(lambda: job.step(gens))
    kms, n = job.engine.pop.kernel_times()
    st1 = job.sync()
    job.close()
    a = (st1["accepted_total"] - st0["accepted_total"]) / float(gens * P)
    alg = P * ((3 + a) * d * 8 + (1 + a) * 8)        # the base row is the single best row (L2 / L1 resident)
    rec = {"workload": f"island DE-best Rosenbrock d={d}, {b.world} island(s) x {P} agents, fp64, ring migration "
                       f"{MIGRANTS} rows / {MIGRATE_EVERY} generations",
           "ms_per_generation": ms / gens, "k2_ms": kms[0] / gens, "k2r_ms": kms[1] / gens, "k3_ms": kms[2] / gens,
           "agent_evals_per_sec": b.world * P * gens / (ms * 1e-3), "accepted_fraction": a,
           "algorithmic_bytes_per_generation": alg, "bytes_formula": "(3 + a) d s + (1 + a) s", "clocks": short_clocks(clocks)}
    rec.update(b.roof(alg, ms / gens))
    rec["k2_frac_of_measured_hbm"] = alg / (kms[0] / gens * 1e-3) / 1e9 / b.peak
    return rec


def bench_config3(b):
    """BASELINE configs[2]: accelerated PSO, Ackley d = 256, 2^21 particles per GPU (2^24 over 8), one min-loc
    exchange per generation — through NCCL (all-gather of the candidate records) and through the fused peer-memory
    kernels.  FP64-instruction-bound (log, sqrt, two cos and two 64-bit draws per coordinate)."""
    import nlsolver_b200 as nb
    from nlsolver_b200.distributed import ShardedPSO
    d, per_gpu = 256, b.args.config3_per_gpu
    P = per_gpu * b.world
    up = np.full(d, 32.768)
    out = {"workload": f"PSO-accelerated Ackley d={d}, {P} particles over {b.world} GPU(s) ({per_gpu} per GPU), fp64, "
                       "one min-loc exchange per generation",
           "bytes_formula": "2 d s + 2 s per particle-generation (read x, write x, pbest value r/w)"}
    for exchange in (["nccl", "peer"] if b.world > 1 else ["none"]):
        cfg = nb.pso_cfg(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, n_particles=P, dim=d, inertia=0.8,
                         cognitive_coef=1.8, social_coef=1.8, eps=0.0, max_iter=NEVER, best_val_no_change=NEVER,
                         seed=0x7c26ca28fb68bc1b)
        job = ShardedPSO(cfg, -up, up, device=b.local, stream=b.stream, exchange="peer" if exchange == "peer" else "nccl")
        job.step(3)
        st0 = job.sync()
        gens = 10
        ms, clocks = b.timed(lambda: job.step(gens))
        st1 = job.sync()
        job.close()
        assert st1["iterations"] - st0["iterations"] == gens
        alg = per_gpu * (2 * d * 8 + 2 * 8)
        rec = {"ms_per_generation": ms / gens, "agent_evals_per_sec": P * gens / (ms * 1e-3),
               "coordinates_per_sec_per_gpu": per_gpu * d * gens / (ms * 1e-3), "clocks": short_clocks(clocks)}
        rec.update(b.roof(alg, ms / gens))
        out[exchange] = rec
    pipe = committed_traffic("pso_move_kernel_accelerated", key="fp64_pipe_pct")
    out["fp64_pipe_pct_ncu"] = pipe
    return out


def bench_sweep(b):
    """BASELINE configs[4] at one point of the sweep, P = 2^22 per GPU, d = 64, Sphere: DE-random and PSO-vanilla
    (corrected social index) in fp32 and fp64."""
    import nlsolver_b200 as nb
    from nlsolver_b200.distributed import ShardedPSO
    d, P = 64, b.args.sweep_pop
    out = []
    for solver in ("DE-random", "PSO-vanilla"):
        for dtype, name, es in ((nb.F64, "fp64", 8), (nb.F32, "fp32", 4)):
            if solver == "DE-random":
                cfg = nb.de_cfg(dtype=dtype, objective=nb.SPHERE, pop_size=P, dim=d, eps=0.0, max_iter=NEVER,
                                best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b)
                w = de_windows(b, cfg, np.full(d, X0), [(None, 3, False), (None, 20, True)], rows_read=4)[0]
                rec = {k: w[k] for k in ("ms_per_generation", "k2_ms", "k2r_ms", "k3_ms", "agent_evals_per_sec",
                                         "accepted_fraction", "achieved_GBps_per_gpu", "frac_of_measured_hbm", "clocks")}
                rec["k2_frac_of_measured_hbm"] = w["algorithmic_bytes_per_generation"] / (w["k2_ms"] * 1e-3) / 1e9 / b.peak
                rec["bytes_formula"] = "(4 + a) d s + (1 + a) s"
            else:
                up = np.full(d, X0)
                cfg = nb.pso_cfg(dtype=dtype, objective=nb.SPHERE, pso_type=nb.PSO_VANILLA, n_particles=P * b.world,
                                 dim=d, eps=0.0, max_iter=NEVER, best_val_no_change=NEVER, seed=0x7c26ca28fb68bc1b,
                                 flags=nb.FLAG_SOCIAL_INDEX_J)
                job = ShardedPSO(cfg, -up, up, device=b.local, stream=b.stream, exchange="nccl")
                job.step(3)
                job.sync()
                gens = 20
                ms, clocks = b.timed(lambda: job.step(gens))
                job.close()
                alg = P * (4 * d * es + 2 * es)
                rec = {"ms_per_generation": ms / gens, "agent_evals_per_sec": b.world * P * gens / (ms * 1e-3),
                       "clocks": short_clocks(clocks), "bytes_formula": "4 d s + 2 s (x and v read + written)"}
                rec.update(b.roof(alg, ms / gens))
            rec.update({"solver": solver, "dtype": name, "population_per_gpu": P, "dim": d, "objective": "Sphere"})
            out.append(rec)
    return out


def multi_gpu_parity(b):
    """N > 1 only, outside every timed region: (1) this rank's island after two migrations over NCCL equals a replay
    of ALL islands on this rank's GPU alone (ring emulated by local copies) bit for bit; (2) the swarm sharded over the
    ranks (NCCL and fused peer exchange) equals the same swarm stepped on this GPU alone bit for bit."""
    import nlsolver_b200 as nb
    from nlsolver_b200 import distributed as D
    torch, world, rank = b.torch, b.world, b.rank
    P, d, every, k, gens = 2048, 48, 2, 8, 5
    kw = dict(objective=nb.ROSENBROCK, strategy=nb.DE_BEST, pop_size=P, dim=d, eps=0.0, max_iter=NEVER,
              best_val_no_change=NEVER, seed=77)
    x0 = np.full(d, 4.096)
    isl = D.IslandDE(nb.de_cfg(**kw), x0, device=b.local, migrate_every=every, migrants=k, stream=b.stream)
    isl.step(gens)
    st = isl.sync()
    mine_rows, mine_scores = isl.engine.pop.population(), isl.engine.pop.scores()
    isl.close()
    ctx = nb.Context(b.local, b.stream.cuda_stream)
    replay = [nb.DEPopulation(ctx, nb.de_cfg(**dict(kw, agent_offset=r * P)), x0) for r in range(world)]
    with torch.cuda.stream(b.stream):
        rows = [torch.zeros(k * d, dtype=torch.float64, device=b.dev) for _ in range(world)]
        scores = [torch.zeros(k, dtype=torch.float64, device=b.dev) for _ in range(world)]
        for g in range(1, gens + 1):
            for s in replay:
                s.step(1)
            if D.migration_due(g, every):
                for r in range(world):
                    replay[r].export_top(k, rows[r].data_ptr(), scores[r].data_ptr())
                for r in range(world):
                    src = D.ring_neighbors(r, world)[1]
                    replay[r].import_migrants(k, rows[src].data_ptr(), scores[src].data_ptr())
    rst = replay[rank].sync()
    islands_ok = (np.array_equal(replay[rank].population(), mine_rows) and np.array_equal(replay[rank].scores(), mine_scores)
                  and rst["f_value"] == st["f_value"] and rst["best_index"] == st["best_index"])
    for s in replay:
        s.close()
    # sharded swarm
    Pg, dp, gens = 4096 + 3, 64, 6
    up = np.full(dp, 32.768)
    pkw = dict(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, n_particles=Pg, dim=dp, eps=0.0, max_iter=NEVER,
               best_val_no_change=NEVER, seed=123)
    whole = nb.PSOSwarm(ctx, nb.pso_cfg(**pkw), -up, up)
    whole.step(gens)
    ws, wpos, wbest = whole.sync(), whole.positions(), whole.best()
    whole.close()
    ctx.close()
    lo, hi = D.slice_bounds(Pg, world, rank)
    swarm_ok = {}
    for exchange in ("nccl", "peer"):
        sw = D.ShardedPSO(nb.pso_cfg(**pkw), -up, up, device=b.local, stream=b.stream, exchange=exchange)
        sw.step(gens)
        ss = sw.sync()
        ok = all(ss[key] == ws[key] for key in ("f_value", "iterations", "function_calls", "best_index", "val_no_change"))
        ok = ok and np.array_equal(sw.best(), wbest) and np.array_equal(sw.engine.swarm.positions(), wpos[lo:hi])
        sw.close()
        swarm_ok[exchange] = b.all_true(ok)
    res = {"islands_equal_single_gpu_replay": b.all_true(islands_ok),
           "sharded_swarm_equals_single_gpu_swarm": swarm_ok,
           "what": f"{world} islands of {P} x {d} (DE-best Rosenbrock), {gens} generations with ring migration every "
                   f"{every}: every rank compares its island (rows, scores, best) with a replay of all islands on its own "
                   f"GPU; swarm of {Pg} x {dp} (PSO-accelerated Ackley) sharded over the ranks vs stepped whole on one GPU; "
                   "bit for bit"}
    assert res["islands_equal_single_gpu_replay"] and all(swarm_ok.values()), res
    return res


def run_b200_arm(args, rank, world):
    b = Bench(args, rank, world)
    headline_shape = (args.pop, args.dim) == (POP, DIM)
    def section(fn, *a):
        """An extra section must never cost the headline line: a failure (e.g. 2 x 64 GiB of rows on a box with less
        free memory) is reported in its place.  Every rank takes the same branch: the flag is reduced over the ranks."""
        try:
            out, err = fn(*a), None
        except Exception as exc:
            out, err = None, f"{type(exc).__name__}: {exc}"[:300]
        if not b.all_true(err is None):
            return {"unavailable": err or "failed on another rank"}
        return out

    cold = None if args.no_extras else section(bench_cold_call, b, args.pop, args.dim, args.steps)
    if cold is not None and "unavailable" in cold:
        cold = {"cold_call": "unavailable: " + cold["unavailable"]}
    main_part = bench_config2(b)
    extra = {}
    if not args.no_extras:
        if world > 1:
            extra["multi_gpu_parity"] = section(multi_gpu_parity, b)
        extra["accepting"] = section(bench_accepting, b)
        extra["configs"] = {"config3_pso_accelerated_ackley_d256": section(bench_config3, b),
                            "config4_island_de_best_rosenbrock_d4096": section(bench_config4, b),
                            "config5_sweep_d64": section(bench_sweep, b)}
    if rank != 0:
        return
    line = {"metric": METRIC, "value": main_part["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_part["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": main_part["config"],
            "clocks": main_part["clocks"], "e2e": main_part["e2e"], "gpu_launches": main_part["gpu_launches"],
            "roofline": main_part["roofline"]}
    if cold:
        line["e2e"].update(cold)
    if extra:
        line["extra"] = extra
    if world == 1 and not args.skip_cpu_baseline and headline_shape:
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
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip the cold call, the accepting regime, "
                                                               "configs 3 / 4 / 5 and the multi-GPU parity check")
    ap.add_argument("--config3-per-gpu", type=int, default=1 << 21)
    ap.add_argument("--config4-pop", type=int, default=1 << 21)
    ap.add_argument("--sweep-pop", type=int, default=1 << 22)
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
