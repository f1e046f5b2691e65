"""Device timing of a SANN chain batch (tuning aid; tools/bench_configs.py is the measurement of record).
usage: python tests/tools/quick_time_sann.py [n_chains] [d] [candidates] [objective] [dtype: 1 f64 | 0 f32]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import nlsolver_b200 as nb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 90
obj = int(sys.argv[4]) if len(sys.argv) > 4 else nb.RASTRIGIN
dtype = int(sys.argv[5]) if len(sys.argv) > 5 else nb.F64
stream = torch.cuda.Stream()
ctx = nb.Context(0, stream.cuda_stream)
cfg = nb.sann_cfg(dtype=dtype, objective=obj, n_chains=n, dim=d, max_iter=1 << 30, temperature_iter=10, seed=1)
ch = nb.SANNChains(ctx, cfg, np.full(d, 2.5))
ch.step(9)
ch.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
ch.step(steps)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
st = ch.sync()
res_f = ch.chains()["f_best"]
print(f"SANN chains={n} d={d} obj={obj} dtype={'f64' if dtype else 'f32'}: {ms/steps:8.4f} ms/candidate-sweep  "
      f"{n*steps/ms*1e3:.4g} chain-evals/s  {n*steps*d/ms*1e3:.4g} coord/s  best {st['f_value']:.6g} median {np.median(res_f):.6g}")
