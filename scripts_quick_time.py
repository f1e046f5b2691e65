import sys, time, numpy as np, torch
import nlsolver_b200 as nb
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
G = int(sys.argv[3]) if len(sys.argv) > 3 else 10
obj = int(sys.argv[4]) if len(sys.argv) > 4 else nb.RASTRIGIN
torch.cuda.init()
stream = torch.cuda.Stream()
ctx = nb.Context(0, stream.cuda_stream)
t0 = time.time()
pop = nb.DEPopulation(ctx, nb.de_cfg(objective=obj, pop_size=P, dim=d, eps=0.0, max_iter=1 << 40, best_val_no_change=1 << 40, seed=1), np.full(d, 10.24))
print("create", time.time() - t0, pop.sync())
pop.step(3); print(pop.sync())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); pop.step(G); e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / G
st = pop.sync()
print(st)
bytes_alg = 4 * d * 8 * P
print(f"P={P} d={d}: {ms:.3f} ms/gen, {P/ms*1e3:.4g} evals/s, alg {bytes_alg/ms/1e6:.1f} GB/s = {bytes_alg/ms/1e6/6550.1:.3f} of measured HBM")
