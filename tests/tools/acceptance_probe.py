"""How does the DE generation cost move with the acceptance rate?  (speculate + repair pays for accepted trials twice)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nlsolver_b200 as nb  # noqa: E402

P = int(sys.argv[1]); d = int(sys.argv[2]); F = float(sys.argv[3]); obj = int(sys.argv[4]); strategy = int(sys.argv[5]) if len(sys.argv) > 5 else 1
stream = torch.cuda.Stream()
ctx = nb.Context(0, stream.cuda_stream)
pop = nb.DEPopulation(ctx, nb.de_cfg(objective=obj, strategy=strategy, pop_size=P, dim=d, differential_weight=F, eps=0.0,
                                     max_iter=1 << 40, best_val_no_change=1 << 40, seed=1), np.full(d, 10.24))
pop.enable_kernel_timing(True)
prev = pop.sync()
for block in range(6):
    pop.step(10)
    st = pop.sync()
    ms, n = pop.kernel_times()
    acc = (st["accepted_total"] - prev["accepted_total"]) / (10 * P)
    rer = (st["repair_reruns"] - prev["repair_reruns"]) / (10 * P)
    rounds = (st["repair_rounds"] - prev["repair_rounds"]) / 10
    print(f"gens {block*10+1:3d}-{block*10+10:3d}: accepted {acc:6.3f} rerun {rer:6.3f} rounds {rounds:5.1f}  "
          f"K2 {ms[0]/n:7.3f} K2r {ms[1]/n:7.3f} K3 {ms[2]/n:6.3f} ms   f={st['f_value']:.5g}")
    prev = st
