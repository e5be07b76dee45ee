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
