"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL (NVLink 5 / NVSwitch).

The only exchange on the path is the gather of the per-rank wake blocks — the reference's two
``comm.Allgatherv`` calls (CSR.py:447-448).  Here it is ONE all-gather of a padded ``[2, ceil(N/P)]``
buffer per rank (<= 512 KiB in total: latency-bound, so one collective instead of two), un-padded
with the reference's count/displ rule (CSR.py:121-125, test/test_mpi.py:14-17).
The same code runs on the gloo backend with CPU tensors for the world_size-2 tests.
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
