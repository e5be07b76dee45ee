"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL (NVLink 5 / NVSwitch).

The only exchange on the path is the gather of the per-rank wake blocks — the reference's two
``comm.Allgatherv`` calls (CSR.py:447-448).  Here it is ONE all-gather of a padded ``[2, ceil(N/P)]``
buffer per rank (<= 512 KiB in total: latency-bound, so one collective instead of two), un-padded
with the reference's count/displ rule (CSR.py:121-125, test/test_mpi.py:14-17).
The same code runs on the gloo backend with CPU tensors for the world_size-2 tests.

On one NVLink/NVSwitch box the exchange is fused into the wake kernel instead (`PeerWakeGrid`): every rank maps the wake
grids of all ranks (torch symmetric memory = CUDA IPC mappings over NVLink) and K4 stores both results of each of its
observation points straight into all of them (`dfcsr_wake_grid_peers`); one cross-rank barrier per step publishes the
grid.  NCCL stays the fallback when peer mappings are not available (`DFCSR_FUSED_GATHER=0` forces it).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def split_counts(n: int, parts: int):
    """count[p] = n // parts (+1 for the first n % parts ranks); displ = exclusive prefix sum."""
    ave, res = divmod(n, parts)
    count = [ave + 1 if p < res else ave for p in range(parts)]
    displ = [sum(count[:p]) for p in range(parts)]
    return count, displ


def init_process_group(backend: str | None = None):
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  Returns (rank, world)."""
    if not dist.is_initialized():
        if "RANK" not in os.environ:
            raise RuntimeError("parallel=True needs a torchrun launch (RANK/WORLD_SIZE/LOCAL_RANK in the environment)")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
        else:
            dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def all_gather_blocks(send: torch.Tensor, count, n: int) -> torch.Tensor:
    """send: (F, pad) with this rank's block in [:, :count[rank]].  Returns (F, n) on every rank."""
    world = dist.get_world_size()
    fields, pad = send.shape
    flat = torch.empty(world * fields * pad, dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(flat, send.contiguous().view(-1))
    recv = flat.view(world, fields, pad)
    if all(c == pad for c in count):
        return recv.permute(1, 0, 2).reshape(fields, n)
    return torch.cat([recv[p, :, :count[p]] for p in range(world)], dim=1)


def xgroup_owner(nx: int, nz: int, world: int, device, group_points: int = 32) -> torch.Tensor:
    """Rank that computes mesh point (ix, iz) when x-groups are dealt out round-robin (group = iz * ceil(nx/gw) + ix // gw
    with gw = dfcsr_xgroup_plan.group_points, flat index = ix * nz + iz as in get_CSR_mesh, CSR.py:382-389)."""
    gw = int(group_points)
    ix = torch.arange(nx, device=device).unsqueeze(1)
    iz = torch.arange(nz, device=device).unsqueeze(0)
    return ((iz * ((nx + gw - 1) // gw) + ix // gw) % world).reshape(-1)


def all_gather_select(mine: torch.Tensor, owner: torch.Tensor) -> torch.Tensor:
    """mine: (F, n) with this rank's points filled in.  Returns (F, n) where point k is taken from rank owner[k]
    (an exact exchange: bits are copied, nothing is added)."""
    world = dist.get_world_size()
    fields, n = mine.shape
    flat = torch.empty(world * fields * n, dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(flat, mine.contiguous().view(-1))
    recv = flat.view(world, fields, n)
    idx = owner.view(1, 1, n).expand(1, fields, n)
    return torch.gather(recv, 0, idx)[0]


class PeerWakeGrid:
    """Wake grids of all ranks mapped into every rank (NVLink peer memory) for the fused K4 + exchange.

    Two (2, n) fp64 grids alternate from step to step: a rank may start writing step k+1 into its peers while they still
    read step k (kick kernel), and the single barrier of step k orders its writes of step k+1 after the peers' readers of
    step k-1, which used the same grid."""

    def __init__(self, n: int, device: torch.device):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 8:
            raise RuntimeError("peer grids are for one NVLink box (<= 8 ranks)")
        self.n = n
        self.grids, self.handles, self.ptrs = [], [], []
        for _ in range(2):
            g = symm.empty((2, n), dtype=torch.float64, device=device)
            g.zero_()
            h = symm.rendezvous(g, group)
            self.grids.append(g)
            self.handles.append(h)
            self.ptrs.append((C.c_uint64 * self.world)(*[int(p) for p in h.buffer_ptrs]))
        self.parity = 0
        self._self_test()

    def _self_test(self):
        """Every rank stores its rank id into slot `rank` of every peer's grid through the mapped pointers, barrier, check:
        a mapping that does not really reach the peers must not go unnoticed."""
        h, g = self.handles[0], self.grids[0]
        if self.n < self.world:
            return
        for p in range(self.world):
            h.get_buffer(p, (2, self.n), torch.float64)[0, self.rank] = float(self.rank + 1)
        h.barrier(channel=0)
        got = g[0, :self.world].cpu().tolist()
        h.barrier(channel=0)
        g.zero_()
        torch.cuda.synchronize()
        h.barrier(channel=0)
        if got != [float(r + 1) for r in range(self.world)]:
            raise RuntimeError(f"peer-memory self test failed on rank {self.rank}: {got}")

    def next(self):
        """(grid tensor, ctypes array of its addresses on all ranks, handle) for this step."""
        k = self.parity
        self.parity ^= 1
        return self.grids[k], self.ptrs[k], self.handles[k]


def make_peer_wake_grid(n: int, device: torch.device):
    """PeerWakeGrid, or None when this job cannot map peer memory (gloo/CPU tests, DFCSR_FUSED_GATHER=0, missing
    P2P access): the caller then uses the NCCL all-gather.  The decision is collective: all ranks agree."""
    if os.environ.get("DFCSR_FUSED_GATHER", "1") == "0" or dist.get_backend() != "nccl" or device.type != "cuda":
        return None
    # peer mappings exist inside one NVLink box only: a job that spans hosts (torchrun sets LOCAL_WORLD_SIZE < WORLD_SIZE)
    # must not even try the symmetric-memory rendezvous, which may hang there instead of raising
    local = int(os.environ.get("LOCAL_WORLD_SIZE", dist.get_world_size()))
    if local != dist.get_world_size() or dist.get_world_size() > 8:
        return None
    ok, grid = 1, None
    try:
        grid = PeerWakeGrid(n, device)
    except Exception as e:            # noqa: BLE001 - any failure here means "no peer memory", NCCL takes over
        ok = 0
        if dist.get_rank() == 0:
            import sys
            print(f"[pydfcsr_b200] fused K4 exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gather",
                  file=sys.stderr)
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return grid if int(flag[0]) == 1 else None


def shard_blocks(rank: int, world: int, blocks: int = 1024):
    """Chunks [first, first + count) of the particle index space owned by `rank`: the 1024 chunks of the
    distribution-independent reductions (include/dfcsr_b200.h: DFCSR_STAT_BLOCKS) are dealt out contiguously."""
    first = (rank * blocks) // world
    return first, ((rank + 1) * blocks) // world - first


class ParticleShards:
    """Particles sharded over the ranks of one job ("optional: shard particles for K1").

    The reference replicates the particles on every MPI rank (CSR.py:202-307).  Here a rank may hold only the chunks
    `shard_blocks` gives it; the three places where all particles meet are made exact so that nothing depends on N:
      * beam statistics / covariance: per-chunk totals in a (1024, NV) table that every rank sums in the same order;
      * CIC deposit: 64-bit fixed-point grids, added over ranks (integer addition is exact);
      * wake mesh: already sharded (block split of the reference), results gathered.
    The tables and grids live in symmetric memory: a rank's kernels store their rows straight into the tables of all ranks
    (NVLink peer memory) and read the fixed-point grids of all ranks; one device-side barrier separates writers from
    readers.  Without peer memory (or with world size 1) NCCL all-reduces do the same."""

    STATS_WORDS = 1024 * 8
    COV_WORDS = 1024 * 27

    def __init__(self, n_total: int, device: torch.device, max_cells: int, rank: int | None = None, world: int | None = None,
                 use_peers: bool | None = None):
        from . import _lib
        self.n_total = int(n_total)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.device = torch.device(device)
        self.chunk = int(_lib.lib.dfcsr_stat_chunk(self.n_total))
        self.first_block, self.n_blocks = shard_blocks(self.rank, self.world, _lib.STAT_BLOCKS)
        self.lo = min(self.first_block * self.chunk, self.n_total)
        self.hi = min((self.first_block + self.n_blocks) * self.chunk, self.n_total)
        self.n_local = self.hi - self.lo
        self.max_cells = int(max_cells)
        words = 2 * self.STATS_WORDS + self.COV_WORDS + 2 * 2 * self.max_cells      # two alternating fixed-point buffers
        self._offsets = {"stats0": 0, "stats1": self.STATS_WORDS, "cov": 2 * self.STATS_WORDS,
                         "q0": 2 * self.STATS_WORDS + self.COV_WORDS,
                         "q1": 2 * self.STATS_WORDS + self.COV_WORDS + 2 * self.max_cells}
        self._q_parity = 0
        self.handle = None
        if use_peers is None:
            use_peers = (self.world > 1 and dist.is_initialized() and dist.get_backend() == "nccl" and self.device.type == "cuda"
                         and os.environ.get("DFCSR_FUSED_GATHER", "1") != "0"
                         and int(os.environ.get("LOCAL_WORLD_SIZE", self.world)) == self.world and self.world <= 8)
        ok = 1
        if use_peers:
            try:
                import torch.distributed._symmetric_memory as symm
                self.buf = symm.empty(words, dtype=torch.float64, device=self.device)
                self.buf.zero_()
                self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
                self._bases = [int(p) for p in self.handle.buffer_ptrs]
            except Exception as e:     # noqa: BLE001 - any failure means "no peer memory": NCCL takes over (collectively)
                ok = 0
                if self.rank == 0:
                    import sys
                    print(f"[pydfcsr_b200] peer memory unavailable for particle shards ({type(e).__name__}: {e}); using NCCL",
                          file=sys.stderr)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag[0]) != 1:
                self.handle = None
        if self.handle is None:
            self.buf = torch.zeros(words, dtype=torch.float64, device=self.device)
            self._bases = [self.buf.data_ptr()]
        self.mode = "peers" if self.handle is not None else ("nccl" if self.world > 1 else "single")

    def local_slice(self, a):
        """This rank's particles of a full-bunch array (last axis = particle index)."""
        return a[..., self.lo:self.hi]

    def _view(self, key, words):
        o = self._offsets[key]
        return self.buf[o:o + words]

    def _peer_ptrs(self, key):
        import ctypes as C
        if self.mode != "peers":
            return None
        return (C.c_uint64 * self.world)(*[b + 8 * self._offsets[key] for b in self._bases])

    def stats_table(self, p):
        t = self._view(f"stats{p}", self.STATS_WORDS)
        if self.mode == "nccl":
            t.zero_()            # rows of the other ranks must be exactly zero for the all-reduce below
        return t, self._peer_ptrs(f"stats{p}")

    def cov_table(self):
        self.barrier()           # nothing but this separates two covariance passes: the peers' readers must be done
        t = self._view("cov", self.COV_WORDS)
        if self.mode == "nccl":
            t.zero_()
        return t, self._peer_ptrs("cov")

    def exchange(self, table):
        """Make every rank's table complete: a barrier after the peer stores, or an all-reduce (rows are disjoint and the
        others zero, so the sum is exact)."""
        if self.mode == "peers":
            self.handle.barrier(channel=0)
        elif self.mode == "nccl":
            dist.all_reduce(table)

    def barrier(self):
        if self.mode == "peers":
            self.handle.barrier(channel=0)

    def q_buffer(self, cells):
        """(int64 view of 2*cells words, addresses of that buffer on all ranks) for this deposit; two buffers alternate."""
        import ctypes as C
        if 2 * cells > 2 * self.max_cells:
            raise ValueError(f"deposit grid of {cells} cells exceeds the {self.max_cells} cells the shard buffers were sized for")
        key = f"q{self._q_parity}"
        self._q_parity ^= 1
        q = self._view(key, 2 * cells).view(torch.int64)
        if self.mode == "peers":
            return q, self._peer_ptrs(key)
        return q, (C.c_uint64 * 1)(q.data_ptr())

    def reduce_q(self, q):
        if self.mode == "peers":
            self.handle.barrier(channel=0)
        elif self.mode == "nccl":
            dist.all_reduce(q)

    def gather(self, t: torch.Tensor) -> torch.Tensor:
        """All particles of a sharded 1-D tensor, on every rank (output / debugging; not on the hot path)."""
        if self.world == 1:
            return t
        sizes = []
        for r in range(self.world):
            f, c = shard_blocks(r, self.world)
            sizes.append(min((f + c) * self.chunk, self.n_total) - min(f * self.chunk, self.n_total))
        pad = max(sizes)
        send = torch.zeros(pad, dtype=t.dtype, device=t.device)
        send[:t.numel()] = t
        recv = torch.empty(self.world * pad, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(recv, send)
        return torch.cat([recv[r * pad:r * pad + sizes[r]] for r in range(self.world)])
