"""Step timing of one PSO configuration (tuning aid). usage: quick_time_pso.py [P] [d] [G] [objective] [type] [dtype]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nlsolver_b200 as nb  # noqa: E402
from oracle import binding as B  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
G = int(sys.argv[3]) if len(sys.argv) > 3 else 20
obj = int(sys.argv[4]) if len(sys.argv) > 4 else nb.ACKLEY
ptype = int(sys.argv[5]) if len(sys.argv) > 5 else nb.PSO_ACCELERATED
dtype = int(sys.argv[6]) if len(sys.argv) > 6 else nb.F64
tag = os.path.basename(os.environ.get("NLS_B200_LIB", "default"))
stream = torch.cuda.Stream()
ctx = nb.Context(0, stream.cuda_stream)
flags = nb.FLAG_SOCIAL_INDEX_J if ptype == nb.PSO_VANILLA else 0
bound = 32.768
up = np.full(d, bound)

# parity probe
p2, g2 = 300, 3
kw = dict(dtype=dtype, objective=obj, pso_type=ptype, n_particles=p2, dim=d, eps=0.0, max_iter=1 << 40,
          best_val_no_change=1 << 40, seed=3)
sw = nb.PSOSwarm(ctx, nb.pso_cfg(flags=flags, **kw), -up, up)
sw.step(g2)
sw.sync()
so, ao = B.pso_run(B.oracle(), B.pso_cfg(social_index_j=bool(flags), **dict(kw, max_iter=g2)), -up, up)
pos = sw.positions()
rel = np.max(np.abs(pos - ao["positions"]) / np.max(np.abs(ao["positions"]), axis=1, keepdims=True))
sw.close()

sw = nb.PSOSwarm(ctx, nb.pso_cfg(dtype=dtype, objective=obj, pso_type=ptype, n_particles=P, dim=d, eps=0.0,
                                 max_iter=1 << 40, best_val_no_change=1 << 40, seed=1, flags=flags), -up, up)
sw.step(3)
sw.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
sw.step(G)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / G
st = sw.sync()
s = 8 if dtype == nb.F64 else 4
alg = (4 if ptype == nb.PSO_VANILLA else 2) * d * s * P
print(f"{tag:24s} PSO type={ptype} obj={obj} dtype={dtype} P={P} d={d}: {ms:7.3f} ms/gen  {P/ms*1e3:.4g} ev/s  "
      f"alg {alg/ms/1e6:7.1f} GB/s ({alg/ms/1e6/6550.1:.3f})  {P*d/ms/1e6:.2f} Gcoord/s  parity rel {rel:.1e} f={st['f_value']:.6g}")
