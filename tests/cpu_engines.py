"""Oracle-backed engines for nlsolver_b200.distributed (CPU tensors, gloo): they let the world_size-2 tests drive the
multi-rank host logic — slices, record exchange, min-loc select, ring migration schedule — without a GPU.
TEST INFRASTRUCTURE: the product's engines are the CUDA ones in nlsolver_b200/distributed.py."""
import numpy as np
import torch

from nlsolver_b200 import distributed as D
from oracle import binding as B


def oracle_de_cfg(c):
    return B.de_cfg(dtype=c.dtype, objective=c.objective, strategy=c.strategy, minimize=bool(c.minimize),
                    pop_size=c.pop_size, dim=c.dim, crossover_prob=c.crossover_prob,
                    differential_weight=c.differential_weight, eps=c.eps, max_iter=c.max_iter,
                    best_val_no_change=c.best_val_no_change, seed=c.seed, agent_offset=c.agent_offset)


def oracle_pso_cfg(c):
    return B.pso_cfg(dtype=c.dtype, objective=c.objective, pso_type=c.pso_type, minimize=bool(c.minimize),
                     n_particles=c.n_particles, dim=c.dim, inertia=c.inertia, cognitive_coef=c.cognitive_coef,
                     social_coef=c.social_coef, eps=c.eps, max_iter=c.max_iter,
                     best_val_no_change=c.best_val_no_change, constrained=bool(c.constrained),
                     social_index_j=bool(c.flags & 2), seed=c.seed, particle_offset=c.particle_offset,
                     n_particles_global=c.n_particles_global)


def moments_of(x):
    x = np.asarray(x, np.float64)
    mean = x.mean() if x.size else 0.0
    return float(x.size), float(mean), float(((x - mean) ** 2).sum())


class OracleDEEngine:
    def __init__(self, cfg, x0):
        self.cfg = cfg
        self.s = B.DEStepper(oracle_de_cfg(cfg), x0)

    def tensor(self, n, dtype):
        return torch.zeros(n, dtype=dtype)

    def step(self, n):
        self.s.advance(n)

    def export_best(self, record):
        st, a = self.s.report()
        rec = D.pack_record(st["f_value"], self.cfg.agent_offset + st["best_index"], moments_of(a["scores"]),
                            a["x_best"])
        record.copy_(torch.from_numpy(rec))

    def export_top(self, k, rows, scores):
        r, s = self.s.export_top(k)
        rows.copy_(torch.from_numpy(r.reshape(-1)))
        scores.copy_(torch.from_numpy(s))

    def import_migrants(self, k, rows, scores):
        self.s.import_migrants(rows.numpy().reshape(k, -1), scores.numpy())

    # the fused exchange on the CPU: the CUDA engine's commit kernel has stored every island's record into every
    # window by the time a reader looks; here the records are gathered when they are read (gloo), which is all the host
    # logic of IslandDE (one engine call per migration interval, barriers around the read) can see of it
    def open_peer_exchange(self, comm, record_bytes):
        self.comm, self.rb = comm, record_bytes

    def read_exchange(self, world):
        mine = torch.zeros(self.rb, dtype=torch.uint8)
        self.export_best(mine)
        gathered = torch.zeros(self.rb * world, dtype=torch.uint8)
        self.comm.all_gather(gathered, mine)
        return gathered.numpy().reshape(world, self.rb)

    def sync(self):
        return self.s.report()[0]

    def close(self):
        self.s.close()


class OraclePSOEngine:
    """Mirrors K7a / K7b on the host with the pure functions of nlsolver_b200.distributed."""

    def __init__(self, cfg, lower, upper):
        self.cfg = cfg
        self.s = B.PSOStepper(oracle_pso_cfg(cfg), lower, upper)
        self.running_best = 100000.0
        self.initial = True
        self._candidate = self._evaluate()

    def _evaluate(self):
        have, v, i, row, pbest = self.s.evaluate()
        return D.pack_record(v if have else np.inf, self.cfg.particle_offset + i, moments_of(pbest), row, have)

    def tensor(self, n, dtype):
        return torch.zeros(n, dtype=dtype)

    def export_candidate(self, record):
        record.copy_(torch.from_numpy(self._candidate))

    def step_local(self, record):
        self.s.move()
        self._candidate = self._evaluate()
        self.export_candidate(record)

    def apply_candidates(self, records, n):
        raw = records.numpy()
        rb = raw.size // n
        heads = [D.parse_record(raw[r * rb:r * rb + D.HEADER_BYTES]) for r in range(n)]
        win = D.select_best(heads, self.running_best)
        mom = D.merge_moments([(h["n"], h["mean"], h["m2"]) for h in heads])
        se = D.std_err_from_moments(*mom)
        n_global = self.cfg.n_particles_global or self.cfg.n_particles
        if win >= 0:
            row = raw[win * rb + D.HEADER_BYTES:(win + 1) * rb].view(np.float64)[:self.cfg.dim]
            self.running_best = heads[win]["value"]
            self.s.adopt(True, heads[win]["value"], heads[win]["index"], row, n_global, se, self.initial)
        else:
            self.s.adopt(False, 0.0, 0, np.zeros(self.cfg.dim), n_global, se, self.initial)
        self.initial = False

    def step(self, n):   # whole swarm on one rank
        rec = self.tensor(self._candidate.size, torch.uint8)
        for _ in range(n):
            self.step_local(rec)
            self.apply_candidates(rec, 1)

    def sync(self):
        return self.s.report()[0]

    def best(self):
        return self.s.report()[1]["x_best"]

    def positions(self):
        return self.s.report()[1]["positions"]

    def close(self):
        self.s.close()


class OracleSANNEngine:
    """A slice of a chain batch on the restatement: the chains are re-run from the start up to the candidates done so
    far (cheap at test sizes), which is exactly what stepping means for chains that never interact."""

    def __init__(self, cfg, x0):
        self.cfg, self.x0 = cfg, np.asarray(x0, np.float64)
        self.total = cfg.max_iter * max(cfg.temperature_iter - 1, 0)
        self.done = 0

    def tensor(self, n, dtype):
        return torch.zeros(n, dtype=dtype)

    def step(self, n):
        self.done = min(self.done + n, self.total)

    def _run(self):
        c = self.cfg
        kw = dict(dtype=c.dtype, objective=c.objective, minimize=bool(c.minimize), n_chains=c.n_chains, dim=c.dim,
                  max_iter=c.max_iter, temperature_iter=c.temperature_iter, temperature_max=c.temperature_max,
                  seed=c.seed, chain_offset=c.chain_offset)
        if self.done == 0:
            kw["max_iter"] = 0
        elif self.done < self.total:
            kw["max_steps"] = self.done
        return B.sann_run(B.oracle(), B.sann_cfg(**kw), self.x0)

    def sync(self):
        so, ao = self._run()
        inner = max(self.cfg.temperature_iter - 1, 1)
        return {"f_value": so["f_value"], "best_index": self.cfg.chain_offset + so["best_index"], "best_valid": 1,
                "iterations": self.cfg.max_iter if self.done >= self.total else self.done // inner,
                "function_calls": self.cfg.n_chains * (1 + self.done), "stopped": int(self.done >= self.total)}

    def best(self):
        so, ao = self._run()
        return ao["x_best"][so["best_index"]]

    def chains(self):
        return self._run()[1]

    def close(self):
        pass
