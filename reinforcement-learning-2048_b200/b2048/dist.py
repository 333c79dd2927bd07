"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on GPUs, gloo in CPU tests).

Environments are independent, so boards are sharded by contiguous global index with no data-path
collective; the only exchange is the sum-allreduce of the online Q-network's
gradient (one flat float64 buffer: 268 KB conv / 3.2 MB dense) and a weight broadcast at start.
With MSELoss(reduction='sum') the global-batch gradient is exactly the sum of the local ones, so
no averaging is applied.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise the default process group from torchrun's environment.  -> (rank, world, local)."""
    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if size > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, size, local


def shard(n_total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous global-index range of `rank`: (index_base, n_local).  Bases are kept multiples
    of 8 so that every shard takes the streaming kernel's aligned Philox path (one call per eight
    consecutive global board indices)."""
    per = (n_total // world_size) & ~7
    base = rank * per
    n_local = per if rank < world_size - 1 else n_total - base
    return base, n_local


class FlatGrads:
    """One contiguous gradient buffer for a module; every parameter's .grad is a view into it, so
    a single collective moves the whole gradient."""

    def __init__(self, module: torch.nn.Module, tail: int = 0):
        """`tail` extra elements are allocated behind the gradients in the SAME allocation
        (`self.tail`): the peer-memory exchange keeps its flags there so that one IPC handle
        covers everything a peer needs."""
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        self.params = params
        n = sum(p.numel() for p in params)
        self.buffer = torch.zeros(n + tail, dtype=params[0].dtype, device=params[0].device)
        self.flat, self.tail = self.buffer[:n], self.buffer[n:]
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self) -> None:
        self.flat.zero_()

    def allreduce_(self) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


class FlatParams:
    """One contiguous parameter buffer for a module; every parameter's .data is a view into it, so
    a single fused optimizer kernel updates the whole network (the module keeps working as is)."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        self.params = params
        self.flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        off = 0
        for p in params:
            p.data = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s weights (and buffers)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
