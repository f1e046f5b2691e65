"""Kernel-level timing of one DE configuration (tuning aid; bench.py is the measurement of record).
usage: python tests/tools/quick_time.py [P] [d] [G] [objective] ; env NLS_B200_LIB selects a library variant."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nlsolver_b200 as nb  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
G = int(sys.argv[3]) if len(sys.argv) > 3 else 20
obj = int(sys.argv[4]) if len(sys.argv) > 4 else nb.RASTRIGIN
strategy = int(sys.argv[5]) if len(sys.argv) > 5 else nb.DE_RANDOM
tag = os.path.basename(os.environ.get("NLS_B200_LIB", "default"))
stream = torch.cuda.Stream()
ctx = nb.Context(0, stream.cuda_stream)

# quick parity probe against the oracle (decisions bit-exact, rows 1e-12)
from oracle import binding as B  # noqa: E402
p2, d2, g2 = 384, d, 2
x0 = np.full(d2, 10.24)
pop = nb.DEPopulation(ctx, nb.de_cfg(objective=obj, strategy=strategy, pop_size=p2, dim=d2, eps=0.0, max_iter=1 << 40,
                                     best_val_no_change=1 << 40, seed=3, flags=nb.FLAG_RECORD_MASKS), x0)
pop.step(g2)
pop.sync()
so, ao = B.de_run(B.oracle(), B.de_cfg(objective=obj, strategy=strategy, pop_size=p2, dim=d2, eps=0.0, max_iter=g2,
                                       best_val_no_change=1 << 40, seed=3), x0, masks=True)
dec = pop.decisions(masks=True)
ok = all(np.array_equal(dec[k], ao[k]) for k in ("donors", "dim_idx", "rejects", "masks", "accepted"))
rows = pop.population()
relerr = np.max(np.abs(rows - ao["rows"]) / np.max(np.abs(ao["rows"]), axis=1, keepdims=True))
serr = np.max(np.abs(dec["trial_scores"] - ao["trial_scores"]) / np.abs(ao["trial_scores"]))
pop.close()

pop = nb.DEPopulation(ctx, nb.de_cfg(objective=obj, strategy=strategy, pop_size=P, dim=d, eps=0.0, max_iter=1 << 40,
                                     best_val_no_change=1 << 40, seed=1), np.full(d, 10.24))
pop.step(5)
pop.sync()
pop.enable_kernel_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
pop.step(G)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / G
kms, n = pop.kernel_times()
st = pop.sync()
alg = 4 * d * 8 * P
print(f"{tag:28s} P={P} d={d} obj={obj}: {ms:7.3f} ms/gen  K2 {kms[0]/n:7.3f} K2r {kms[1]/n:6.3f} K3 {kms[2]/n:6.3f}  "
      f"{P/ms*1e3:.4g} ev/s  K2 alg {alg/(kms[0]/n)/1e6:7.1f} GB/s ({alg/(kms[0]/n)/1e6/6550.1:.3f})  "
      f"parity={'ok' if ok else 'FAIL'} rows {relerr:.1e} scores {serr:.1e} acc {st['accepted_total']} reruns {st['repair_reruns']}")
