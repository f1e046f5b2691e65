"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing (NCCL on GPUs, gloo in the CPU tests).

What shards:
  * PSO (both types) shards exactly: particles only interact through the previous generation's swarm best, so each
    rank owns a contiguous slice of GLOBAL particle ids (draw streams are keyed by the global id) and a generation is
    `step_local` -> all-gather of one candidate record per rank -> `apply_candidates`.  Results are identical to the
    single-GPU swarm.  The all-gather + lowest-index-wins select is the min-loc all-reduce NCCL does not have natively.
  * DE does not shard one population (every agent gathers three uniformly random rows and the in-place order spans the
    whole population), so large DE runs as islands: each rank runs a reference-exact DE on its own population; every
    generation the island bests are all-gathered (global status), and every `migrate_every` generations the `migrants`
    best rows travel around the ring rank -> rank + 1 and overwrite the receiver's worst rows.

  * SANN chains never interact: a batch splits into contiguous slices of GLOBAL chain ids with no collective in the
    loop at all; the batch result (best value, global chain id, its point) is one all-gather of records at the end.

The pure functions at the top (slices, ring, record select) are the host logic the world_size-2 gloo tests cover; the
compute behind `engine` is any object with the small interface used below (the CUDA handles on GPUs).
"""
import struct

import numpy as np

HEADER_BYTES = 48   # RecordHeader: value f64, index u64, moments 3 x f64, valid i32, pad i32 (csrc/reduce.cuh)


# ------------------------------------------------------------------ pure host logic -------------------------------
def slice_bounds(n_global, world_size, rank):
    """Contiguous slice [begin, end) of rank: the first n_global % world_size ranks hold one extra element."""
    base, extra = divmod(n_global, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def ring_neighbors(rank, world_size):
    """(destination, source) of the migration ring."""
    return (rank + 1) % world_size, (rank - 1) % world_size


def record_bytes(elem_size, dim):
    return HEADER_BYTES + (dim * elem_size + 7) // 8 * 8


def parse_record(buf):
    """Header of one exchange record (bytes-like) -> dict."""
    value, index, n, mean, m2, valid, _ = struct.unpack_from("<dQdddii", bytes(buf[:HEADER_BYTES]))
    return {"value": value, "index": index, "n": n, "mean": mean, "m2": m2, "valid": valid}


def select_best(records, running_best=float("inf")):
    """Index of the record that wins a strict-< scan in rank order against `running_best`, or -1.
    Ranks hold ascending index ranges, so this is the reference's sequential scan (nlsolver.h:2723-2729)."""
    win, best = -1, running_best
    for r, rec in enumerate(records):
        if rec["valid"] and rec["value"] < best:
            best, win = rec["value"], r
    return win


def migration_due(generation, migrate_every):
    """Migration happens after generations migrate_every, 2*migrate_every, ... (generation counts from 1)."""
    return migrate_every > 0 and generation > 0 and generation % migrate_every == 0


# ------------------------------------------------------------------ process group ---------------------------------
def init_from_env(backend=None):
    """Join the process group torchrun describes (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world)."""
    import os

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world


class _Comm:
    """all_gather / ring exchange of byte tensors on the tensors' own device (NCCL for cuda, gloo for cpu)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.active = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.active else 0
        self.world = dist.get_world_size(group) if self.active else 1

    def all_gather(self, out, inp):
        if self.world == 1:
            out.copy_(inp)
        else:
            self.dist.all_gather_into_tensor(out, inp, group=self.group)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def ring_exchange(self, send, recv):
        """send -> rank + 1, recv <- rank - 1."""
        if self.world == 1:
            recv.copy_(send)
            return
        dst, src = ring_neighbors(self.rank, self.world)
        ops = [self.dist.P2POp(self.dist.isend, send, dst, self.group),
               self.dist.P2POp(self.dist.irecv, recv, src, self.group)]
        for w in self.dist.batch_isend_irecv(ops):
            w.wait()


def merge_moments(parts):
    """Chan / Golub / LeVeque combination of (n, mean, m2) triples in the given order -> (n, mean, m2)."""
    n, mean, m2 = 0.0, 0.0, 0.0
    for pn, pmean, pm2 in parts:
        if pn == 0.0:
            continue
        if n == 0.0:
            n, mean, m2 = pn, pmean, pm2
            continue
        tot = n + pn
        delta = pmean - mean
        mean, m2, n = mean + delta * (pn / tot), m2 + pm2 + delta * delta * (n * pn / tot), tot
    return n, mean, m2


def std_err_from_moments(n, mean, m2):
    """std_err (nlsolver.h:2037-2052) from merged moments: sqrt(sum((x - mean)^2) / (n - 1))."""
    return float(np.sqrt(m2 / (n - 1.0))) if n > 1.0 else float("nan")


def pack_record(value, index, moments, row, valid=True):
    """Build one exchange record (numpy uint8) — the layout the CUDA kernels write (csrc/reduce.cuh RecordHeader)."""
    row = np.ascontiguousarray(row)
    body = row.tobytes()
    body += b"\0" * (-len(body) % 8)
    head = struct.pack("<dQdddii", float(value), int(index) & 0xFFFFFFFFFFFFFFFF, *[float(m) for m in moments],
                       int(bool(valid)), 0)
    return np.frombuffer(head + body, dtype=np.uint8).copy()


# ------------------------------------------------------------------ engines ---------------------------------------
class CudaDEEngine:
    """The CUDA island: nls_de_* behind the interface IslandDE drives (tensors in, raw device pointers out)."""

    def __init__(self, cfg, x0, device, stream):
        import torch

        from .solvers import Context, DEPopulation
        self.torch = torch
        self.ctx = Context(device, stream.cuda_stream)
        self.pop = DEPopulation(self.ctx, cfg, x0)
        self.device = f"cuda:{device}"
        self.kernel_launches = 0

    def tensor(self, n, dtype):
        return self.torch.zeros(n, dtype=dtype, device=self.device)

    def step(self, n):
        self.pop.step(n)
        self.kernel_launches += 3 * n

    def export_best(self, record):
        self.pop.export_best(record.data_ptr())
        self.kernel_launches += 1

    def export_top(self, k, rows, scores):
        self.pop.export_top(k, rows.data_ptr(), scores.data_ptr())
        self.kernel_launches += 3

    def import_migrants(self, k, rows, scores):
        self.pop.import_migrants(k, rows.data_ptr(), scores.data_ptr())
        self.kernel_launches += 4

    def open_peer_exchange(self, comm, record_bytes):
        """Map every rank's exchange window into this process (CUDA IPC) and attach it to the island: from here on the
        commit kernel of every generation stores the island's record into every peer's window over NVLink."""
        from .solvers import ExchangeWindow
        self.window = ExchangeWindow(self.ctx, record_bytes, comm.world, comm.rank)
        if comm.world > 1:
            handles = [None] * comm.world
            comm.dist.all_gather_object(handles, self.window.ipc_handle(), group=comm.group)
            self.window.open_peers(handles)
        self.pop.attach_exchange(self.window)
        self.kernel_launches += 1

    def read_exchange(self, world):
        return self.pop.read_exchange(world)

    def sync(self):
        return self.pop.sync()

    def close(self):
        self.pop.close()
        if getattr(self, "window", None) is not None:
            self.window.close()
        self.ctx.close()


class CudaPSOEngine:
    def __init__(self, cfg, lower, upper, device, stream):
        import torch

        from .solvers import Context, PSOSwarm
        self.torch = torch
        self.ctx = Context(device, stream.cuda_stream)
        self.swarm = PSOSwarm(self.ctx, cfg, lower, upper)
        self.device = f"cuda:{device}"

    def tensor(self, n, dtype):
        return self.torch.zeros(n, dtype=dtype, device=self.device)

    def export_candidate(self, record):
        self.swarm.export_candidate(record.data_ptr())

    def step_local(self, record):
        self.swarm.step_local(record.data_ptr())

    def step(self, n):
        self.swarm.step(n)

    def apply_candidates(self, records, n):
        self.swarm.apply_candidates(records.data_ptr(), n)

    def open_peer_exchange(self, comm, record_bytes):
        """Map every rank's exchange window into this process (CUDA IPC) and attach it to the swarm: from here on a
        generation needs no host-side collective (nls_pso_step_fused)."""
        from .solvers import ExchangeWindow
        self.window = ExchangeWindow(self.ctx, record_bytes, comm.world, comm.rank)
        if comm.world > 1:
            handles = [None] * comm.world
            comm.dist.all_gather_object(handles, self.window.ipc_handle(), group=comm.group)
            self.window.open_peers(handles)
        self.swarm.attach_exchange(self.window)

    def step_fused(self, n):
        self.swarm.step_fused(n)

    def sync(self):
        return self.swarm.sync()

    def best(self):
        return self.swarm.best()

    def close(self):
        self.swarm.close()
        if getattr(self, "window", None) is not None:
            self.window.close()
        self.ctx.close()


class CudaSANNEngine:
    """The CUDA slice of a chain batch: nls_sann_* behind the interface ShardedSANN drives."""

    def __init__(self, cfg, x0, device, stream):
        import torch

        from .solvers import Context, SANNChains
        self.torch = torch
        self.ctx = Context(device, stream.cuda_stream)
        self.chains_h = SANNChains(self.ctx, cfg, x0)
        self.device = f"cuda:{device}"
        self.kernel_launches = 1

    def tensor(self, n, dtype):
        return self.torch.zeros(n, dtype=dtype, device=self.device)

    def step(self, n):
        self.chains_h.step(n)
        self.kernel_launches += 1

    def sync(self):
        return self.chains_h.sync()

    def best(self):
        return self.chains_h.best()

    def chains(self):
        return self.chains_h.chains()

    def close(self):
        self.chains_h.close()
        self.ctx.close()


class _NullStream:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def synchronize(self):
        pass


# ------------------------------------------------------------------ sharded PSO -----------------------------------
class ShardedPSO:
    """A global swarm of `cfg.n_particles` particles split across the ranks of `group`."""

    def __init__(self, cfg, lower, upper, device=0, group=None, stream=None, engine_factory=None, exchange="nccl"):
        """exchange: "nccl" — all-gather of the candidate records through torch.distributed every generation;
        "peer" — the fused kernels store the records straight into the peers' HBM over NVLink (CUDA IPC windows),
        no host-side collective in the loop.  Both give identical results."""
        import torch

        from . import _lib as L
        from .solvers import pso_cfg
        self.torch = torch
        self.comm = _Comm(group)
        n_global = cfg.n_particles
        begin, end = slice_bounds(n_global, self.comm.world, self.comm.rank)
        local = pso_cfg(cfg.dtype, cfg.objective, cfg.pso_type, bool(cfg.minimize), end - begin, cfg.dim, cfg.inertia,
                        cfg.cognitive_coef, cfg.social_coef, cfg.eps, cfg.max_iter, cfg.best_val_no_change,
                        bool(cfg.constrained), cfg.flags, cfg.seed, begin, n_global)
        self.n_local, self.n_global = end - begin, n_global
        if engine_factory is None:
            self.stream = stream or torch.cuda.Stream(device)
            self._scope = lambda: torch.cuda.stream(self.stream)
            self.engine = CudaPSOEngine(local, lower, upper, device, self.stream)
        else:
            self.stream = _NullStream()
            self._scope = lambda: self.stream
            self.engine = engine_factory(local, lower, upper)
        self.rb = record_bytes(8 if cfg.dtype == L.F64 else 4, cfg.dim)
        self.fused = exchange == "peer"
        with self._scope():
            self.mine = self.engine.tensor(self.rb, torch.uint8)
            self.all = self.engine.tensor(self.rb * self.comm.world, torch.uint8)
            if self.fused:
                self.engine.open_peer_exchange(self.comm, self.rb)
            elif self.comm.world > 1:   # finish the first update_best_positions across shards (nlsolver.h:2595)
                self.engine.export_candidate(self.mine)
                self.comm.all_gather(self.all, self.mine)
                self.engine.apply_candidates(self.all, self.comm.world)

    def step(self, n=1):
        with self._scope():
            if self.fused:
                self.engine.step_fused(n)
                return
            if self.comm.world == 1:
                self.engine.step(n)
                return
            for _ in range(n):
                self.engine.step_local(self.mine)
                self.comm.all_gather(self.all, self.mine)
                self.engine.apply_candidates(self.all, self.comm.world)

    def sync(self):
        return self.engine.sync()

    def best(self):
        return self.engine.best()

    def close(self):
        self.engine.close()


# ------------------------------------------------------------------ island DE -------------------------------------
class IslandDE:
    """One reference-exact DE island per rank; ring migration every `migrate_every` generations.  The island bests are
    exchanged every generation: exchange="peer" — the commit kernel itself stores the island's record into every
    peer's window over NVLink (CUDA IPC), nothing is launched or waited for, and generations between two migrations
    are one C call; exchange="nccl" — export kernel + all-gather through torch.distributed after every generation
    (also the path of the gloo CPU tests).  Default: "peer" on GPUs."""

    def __init__(self, cfg, x0, device=0, group=None, migrate_every=10, migrants=64, stream=None,
                 engine_factory=None, exchange=None):
        import torch

        from . import _lib as L
        from .solvers import de_cfg
        self.torch = torch
        self.comm = _Comm(group)
        self.migrate_every, self.k = migrate_every, min(migrants, cfg.pop_size)
        # islands draw from disjoint streams: global agent ids rank * P + i
        local = de_cfg(cfg.dtype, cfg.objective, cfg.strategy, bool(cfg.minimize), cfg.pop_size, cfg.dim,
                       cfg.crossover_prob, cfg.differential_weight, cfg.eps, cfg.max_iter, cfg.best_val_no_change,
                       cfg.seed, cfg.agent_offset + self.comm.rank * cfg.pop_size, cfg.flags)
        self.cfg = local
        if engine_factory is None:
            self.stream = stream or torch.cuda.Stream(device)
            self._scope = lambda: torch.cuda.stream(self.stream)
            self.engine = CudaDEEngine(local, x0, device, self.stream)
        else:
            self.stream = _NullStream()
            self._scope = lambda: self.stream
            self.engine = engine_factory(local, x0)
        if exchange is None:
            exchange = "peer" if engine_factory is None else "nccl"
        self.fused = exchange == "peer"
        self.island = getattr(self.engine, "pop", None)
        self.generation = 0
        self.rb = record_bytes(8 if cfg.dtype == L.F64 else 4, cfg.dim)
        self._records = None
        tdt = torch.float64 if cfg.dtype == L.F64 else torch.float32
        with self._scope():
            self.out_rows = self.engine.tensor(self.k * cfg.dim, tdt)
            self.out_scores = self.engine.tensor(self.k, tdt)
            self.in_rows = self.engine.tensor(self.k * cfg.dim, tdt)
            self.in_scores = self.engine.tensor(self.k, tdt)
            if self.fused:
                self.engine.open_peer_exchange(self.comm, self.rb)
            else:
                self.mine = self.engine.tensor(self.rb, torch.uint8)
                self.all = self.engine.tensor(self.rb * self.comm.world, torch.uint8)
            if self.comm.world > 1:
                # NCCL builds its collective and point-to-point channels lazily (seconds): do it here, not in the
                # first generation / first migration
                if not self.fused:
                    self.comm.all_gather(self.all, self.mine)
                self.comm.ring_exchange(self.out_rows, self.in_rows)
                self.comm.ring_exchange(self.out_scores, self.in_scores)

    @property
    def launches(self):
        return getattr(self.engine, "kernel_launches", 0)

    def _migrate(self):
        self.engine.export_top(self.k, self.out_rows, self.out_scores)
        self.comm.ring_exchange(self.out_rows, self.in_rows)
        self.comm.ring_exchange(self.out_scores, self.in_scores)
        self.engine.import_migrants(self.k, self.in_rows, self.in_scores)

    def step(self, n=1):
        migrating = self.migrate_every > 0 and self.comm.world > 1 and self.k > 0
        with self._scope():
            if self.fused:
                left = n
                while left > 0:        # up to the next migration point in one call
                    chunk = min(left, self.migrate_every - self.generation % self.migrate_every) if migrating else left
                    self.engine.step(chunk)
                    self.generation += chunk
                    left -= chunk
                    if migrating and migration_due(self.generation, self.migrate_every):
                        self._migrate()
                return
            for _ in range(n):
                self.engine.step(1)
                self.generation += 1
                self.engine.export_best(self.mine)
                self.comm.all_gather(self.all, self.mine)
                if migrating and migration_due(self.generation, self.migrate_every):
                    self._migrate()

    def _gathered(self):
        """uint8 [world, record_bytes]: the newest record of every island."""
        if self.fused:
            # every island has finished its enqueued generations (-> its stores into the peers' windows have landed)
            # before anybody reads, and nobody steps on before everybody has read
            self.stream.synchronize()
            with self._scope():
                self.comm.barrier()
                self.stream.synchronize()
                recs = self.engine.read_exchange(self.comm.world)
                self.comm.barrier()
            self.stream.synchronize()
            return recs
        self.stream.synchronize()
        return self.all.cpu().numpy().reshape(self.comm.world, self.rb)

    def sync(self):
        """Local island status plus the global best over the islands' newest records."""
        st = self.engine.sync()
        self._records = recs = self._gathered()
        heads = [parse_record(recs[r, :HEADER_BYTES]) for r in range(self.comm.world)]
        win = select_best(heads)
        st["global_best_value"] = heads[win]["value"] if win >= 0 else st["f_value"]
        st["global_best_rank"] = win
        return st

    def global_best_row(self):
        st = self.sync()
        r = max(st["global_best_rank"], 0)
        dt = np.float64 if self.cfg.dtype == 1 else np.float32
        return self._records[r, HEADER_BYTES:].view(dt)[:self.cfg.dim].copy()

    def close(self):
        self.engine.close()


# ------------------------------------------------------------------ sharded SANN chains ---------------------------
class ShardedSANN:
    """A batch of `cfg.n_chains` annealing chains split across the ranks of `group`: rank r owns the contiguous slice
    of global chain ids slice_bounds gives it (draw streams are keyed by the global id, so the batch is the same for
    any world size).  Nothing is exchanged while the chains run."""

    def __init__(self, cfg, x0, device=0, group=None, stream=None, engine_factory=None):
        import torch

        from . import _lib as L
        from .solvers import sann_cfg
        self.torch = torch
        self.comm = _Comm(group)
        self.n_global = cfg.n_chains
        begin, end = slice_bounds(cfg.n_chains, self.comm.world, self.comm.rank)
        self.begin, self.end = begin, end
        local = sann_cfg(cfg.dtype, cfg.objective, bool(cfg.minimize), end - begin, cfg.dim, cfg.max_iter,
                         cfg.temperature_iter, cfg.temperature_max, cfg.seed, cfg.chain_offset + begin, cfg.flags)
        self.cfg = local
        x0 = np.asarray(x0)
        x0_local = x0 if x0.ndim == 1 else x0[begin:end]
        if engine_factory is None:
            self.stream = stream or torch.cuda.Stream(device)
            self._scope = lambda: torch.cuda.stream(self.stream)
            self.engine = CudaSANNEngine(local, x0_local, device, self.stream)
        else:
            self.stream = _NullStream()
            self._scope = lambda: self.stream
            self.engine = engine_factory(local, x0_local)
        self.elem = 8 if cfg.dtype == L.F64 else 4
        self.rb = record_bytes(self.elem, cfg.dim)

    @property
    def launches(self):
        return getattr(self.engine, "kernel_launches", 0)

    def step(self, n=1):
        """n candidates per chain on every rank (clamped to the end of the schedule)."""
        with self._scope():
            self.engine.step(n)

    def run(self):
        self.step(0xFFFFFFFFFFFFFFFF)

    def sync(self):
        """This rank's slice status (f_value / best_index: its best chain, global id)."""
        return self.engine.sync()

    def global_best(self):
        """(status, point) of the best chain of the whole batch: lowest best_val, lowest global chain id on ties —
        one all-gather of one record per rank."""
        st = self.engine.sync()
        row = self.engine.best()
        rec = pack_record(st["f_value"], st["best_index"], (0.0, 0.0, 0.0), row, valid=bool(st["best_valid"]))
        with self._scope():
            mine = self.engine.tensor(self.rb, self.torch.uint8)
            allr = self.engine.tensor(self.rb * self.comm.world, self.torch.uint8)
            mine.copy_(self.torch.from_numpy(rec))
            self.comm.all_gather(allr, mine)
        self.stream.synchronize()
        raw = allr.cpu().numpy()
        heads = [parse_record(raw[r * self.rb:r * self.rb + HEADER_BYTES]) for r in range(self.comm.world)]
        win = select_best(heads)   # ranks hold ascending chain ids: strict < in rank order = lowest id on ties
        out = dict(st)
        out["function_calls"] = self.n_global * (st["function_calls"] // max(self.end - self.begin, 1))
        if win >= 0:
            out["f_value"], out["best_index"] = heads[win]["value"], heads[win]["index"]
            dt = np.float64 if self.elem == 8 else np.float32
            row = raw[win * self.rb + HEADER_BYTES:(win + 1) * self.rb].view(dt)[:self.cfg.dim].copy()
        out["best_rank"] = win
        return out, row

    def close(self):
        self.engine.close()
