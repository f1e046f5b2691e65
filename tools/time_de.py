"""Device milliseconds per DE generation (CUDA events around step(20), best of three) for one shape — the quick A/B aid
behind the NLS_DE_* environment switches.   usage: python tools/time_de.py <pop> <dim> <f32|f64> <F>"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import nlsolver_b200 as nb
P, d, dt, F = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], float(sys.argv[4])
stream = torch.cuda.Stream(); ctx = nb.Context(0, stream.cuda_stream)
pop = nb.DEPopulation(ctx, nb.de_cfg(dtype=nb.F64 if dt == "f64" else nb.F32, objective=nb.SPHERE, pop_size=P, dim=d,
      differential_weight=F, eps=0.0, max_iter=1 << 40, best_val_no_change=1 << 40, seed=1), np.full(d, 10.24))
pop.step(5); pop.sync()
res = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); pop.step(20); e1.record(stream); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 20)
st = pop.sync()
print(f"FUSE={os.environ.get('NLS_DE_FUSE_COMMIT','1')} P={P} d={d} {dt} F={F}: ms/gen {min(res):.4f} {res}  acc_total={st['accepted_total']}")
