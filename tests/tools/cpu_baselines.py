"""Single-thread CPU timings of the UNMODIFIED reference (oracle/_ref) on the shapes of BASELINE.json configs 1-5 at
reduced populations (per-agent CPU cost is population-independent beyond cache size).  Prints a
markdown table.  bench.py times config 2 itself in the same run as the GPU measurement."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import binding as B  # noqa: E402

ref = B.reference()
assert ref is not None, "oracle/_ref/libnls_ref.so missing"
NEVER = 1 << 40
rows = []


def de(name, dtype, obj, strategy, P, d, G, scale):
    cfg = B.de_cfg(dtype=dtype, objective=obj, strategy=strategy, pop_size=P, dim=d, eps=0.0, max_iter=G,
                   best_val_no_change=NEVER)
    x0 = np.full(d, scale, dtype=B.np_dtype(dtype))
    sec, st = C.c_double(), B.Status()
    assert ref.ref_de_time(C.byref(cfg), x0.ctypes.data, C.byref(sec), C.byref(st)) == 0
    rows.append((name, P, d, G, st.function_calls / sec.value))


def pso(name, dtype, obj, ptype, P, d, G, bound):
    cfg = B.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, n_particles=P, dim=d, eps=0.0, max_iter=G,
                    best_val_no_change=NEVER)
    up = np.full(d, bound, dtype=B.np_dtype(dtype))
    sec, st = C.c_double(), B.Status()
    assert ref.ref_pso_time(C.byref(cfg), up.ctypes.data, C.byref(sec), C.byref(st)) == 0
    rows.append((name, P, d, G, st.function_calls / sec.value))


de("config 1: DE-random Rosenbrock(example) d=2 fp64, defaults (pop 50)", B.F64, B.ROSENBROCK_EX, B.DE_RANDOM, 50, 2, 1000, 5.0)
de("config 2: DE-random Rastrigin d=1000 fp64", B.F64, B.RASTRIGIN, B.DE_RANDOM, 4096, 1000, 10, 10.24)
pso("config 3: PSO-accelerated Ackley d=256 fp64", B.F64, B.ACKLEY, B.PSO_ACCELERATED, 8192, 256, 20, 32.768)
de("config 4: DE-best Rosenbrock d=4096 fp64", B.F64, B.ROSENBROCK, B.DE_BEST, 1024, 4096, 10, 4.096)
de("config 5: DE-random Sphere d=64 fp64", B.F64, B.SPHERE, B.DE_RANDOM, 65536, 64, 10, 10.24)
de("config 5: DE-random Sphere d=64 fp32", B.F32, B.SPHERE, B.DE_RANDOM, 65536, 64, 10, 10.24)
pso("config 5: PSO-accelerated Sphere d=64 fp64", B.F64, B.SPHERE, B.PSO_ACCELERATED, 65536, 64, 10, 10.24)
pso("config 5: PSO-accelerated Sphere d=64 fp32", B.F32, B.SPHERE, B.PSO_ACCELERATED, 65536, 64, 10, 10.24)
pso("config 5: PSO-vanilla Sphere d=64 fp64 (P <= d: the reference reads out of bounds otherwise)", B.F64, B.SPHERE,
    B.PSO_VANILLA, 64, 64, 2000, 10.24)
print(f"host: {os.cpu_count()} hardware threads; reference = nlsolver::DE / nlsolver::PSO from /root/reference, "
      "g++ -O2 -ffp-contract=off, xorshift<T>, ONE thread\n")
print("| workload | P (CPU run) | d | generations | agent-evals/s, 1 thread |")
print("|---|---|---|---|---|")
for name, P, d, G, v in rows:
    print(f"| {name} | {P} | {d} | {G} | {v:.3e} |")
