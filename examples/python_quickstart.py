"""Python quickstart for nlsolver_b200 (needs a CUDA device; there is no CPU path).

    python examples/python_quickstart.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/python_quickstart.py   # + multi-GPU part
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nlsolver_b200 as nb  # noqa: E402


class Uniform:
    """Any callable returning floats in [0, 1] works as the generator; two draws seed the device draw tape."""

    def __init__(self, seed=42):
        self.rng = np.random.default_rng(seed)

    def __call__(self):
        return float(self.rng.random())


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    gen = Uniform()

    # 1. the reference interface: DE(...).minimize(x) writes the best point back into x and returns the status
    x = [5.0, 7.0]
    status = nb.DE(nb.RosenbrockExample, gen).minimize(x)
    if rank == 0:
        status.print()
        print(x)

    # 2. a larger problem, accelerated PSO with bounds
    d = 64
    x = [0.0] * d
    status = nb.PSO(nb.Rastrigin, gen, n_particles=1 << 16, max_iter=200, pso_type=nb.PSOType.Accelerated).minimize(
        x, [-5.12] * d, [5.12] * d)
    if rank == 0:
        print("accelerated PSO, Rastrigin d=64, 65536 particles:", status.f_value, "after", status.iteration, "iterations")

    # 3. the loop cut at generation boundaries: population resident in HBM, inspect whatever you need
    ctx = nb.default_context()
    pop = nb.DEPopulation(ctx, nb.de_cfg(objective=nb.SPHERE, pop_size=1 << 18, dim=32, eps=0.0, max_iter=1 << 40,
                                         best_val_no_change=1 << 40, seed=7), np.full(32, 10.24))
    for _ in range(5):
        pop.step(20)
        st = pop.sync()
        if rank == 0:
            print(f"generation {st['iterations']:4d}: best {st['f_value']:.6g}, accepted so far {st['accepted_total']}")
    pop.close()

    # 4. simulated annealing: one chain as in the reference, then 65536 independent chains from the same start
    x = [5.0, 5.0]
    status = nb.SANN(nb.RosenbrockExample, gen).minimize(x)
    sann = nb.SANN(nb.Rastrigin, gen, max_iter=500)
    x8 = [3.3] * 8
    many = sann.minimize_multistart(x8, 1 << 16)
    if rank == 0:
        print("SANN, one chain, Rosenbrock:", status.f_value, x)
        print("SANN, 65536 chains, Rastrigin d=8: best", many.f_value, "after", many.function_calls_used, "objective calls")

    # 5. multi-GPU (under torchrun): one global swarm sharded across the ranks, exchange over peer memory
    if world > 1:
        from nlsolver_b200 import distributed as D
        D.init_from_env("nccl")
        local = int(os.environ.get("LOCAL_RANK", "0"))
        up = np.full(256, 32.768)
        swarm = D.ShardedPSO(nb.pso_cfg(objective=nb.ACKLEY, pso_type=nb.PSO_ACCELERATED, n_particles=(1 << 20) * world,
                                        dim=256, eps=0.0, max_iter=1 << 40, best_val_no_change=1 << 40, seed=1),
                             -up, up, device=local, exchange="peer")
        swarm.step(50)
        st = swarm.sync()
        if rank == 0:
            print(f"{world} GPUs, {(1 << 20) * world} particles: swarm best {st['f_value']:.6g} after {st['iterations']} generations")
        swarm.close()


if __name__ == "__main__":
    main()
