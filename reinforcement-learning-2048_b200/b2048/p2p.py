"""NVLink peer-memory gradient exchange: every rank maps its peers' flat gradient buffers (CUDA IPC)
and one kernel per rank sums them in rank order and applies Adam (`p2p_allreduce_adam_f64`).

Replaces `dist.all_reduce(flat_grads)` + a separate optimizer kernel for the data-parallel update:
the messages are tiny (268 KB conv / 3.2 MB dense), so the cost is launch + synchronisation
latency, which one fused kernel with two flag barriers keeps to a few microseconds.  All ranks add
the W buffers in the same order, so their parameter replicas stay bit-identical.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from .env import _ptr, _stream


_opened: dict[bytes, int] = {}          # IPC handle -> mapped base address (a handle opens once per process)


def _export(t: torch.Tensor):
    """(64-byte cudaIpcMemHandle of the allocation holding `t`, byte offset of `t` inside it)."""
    import ctypes
    st = t.untyped_storage()._share_cuda_()      # torch's IPC export: st[3] = offset of the storage in its allocation
    buf = ctypes.create_string_buffer(64)
    with torch.cuda.device(t.device):
        _lib.check(_lib.lib().p2p_get_ipc_handle(_ptr(t), buf), "p2p_get_ipc_handle")
    return buf.raw, int(st[3]) + t.storage_offset() * t.element_size()


def _import(desc) -> int:
    handle, byte_offset = desc
    if handle not in _opened:
        out = _lib.c_void_p()
        _lib.check(_lib.lib().p2p_open_ipc_handle(handle, out), "p2p_open_ipc_handle")
        _opened[handle] = int(out.value)
    return _opened[handle] + byte_offset


class PeerGradExchange:
    """`grads` = FlatGrads created with tail >= 2 * world: gradients and flag block share one allocation."""

    def __init__(self, grads):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGradExchange needs an initialised process group (one process per GPU)")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        flat = grads.flat
        self.device = flat.device
        if grads.tail.numel() < 2 * self.world or grads.tail.element_size() != 8:
            raise ValueError("FlatGrads needs tail >= 2 * world_size float64 slots for the flag block")
        self.flags = grads.tail[:2 * self.world].view(torch.int64)
        self.flags.zero_()
        self.sync = torch.zeros(4, dtype=torch.int64, device=self.device)       # epoch, blocks done, error, -
        torch.cuda.synchronize(self.device)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (_export(flat), _export(self.flags)))
        g_ptrs, f_ptrs = [], []
        with torch.cuda.device(self.device):          # map the peers' memory for access from MY device
            for r, (g_desc, f_desc) in enumerate(everyone):
                if r == self.rank:
                    g_ptrs.append(flat.data_ptr()); f_ptrs.append(self.flags.data_ptr())
                else:
                    g_ptrs.append(_import(g_desc)); f_ptrs.append(_import(f_desc))
        self.peer_grads = torch.tensor(g_ptrs, dtype=torch.int64, device=self.device)
        self.peer_flags = torch.tensor(f_ptrs, dtype=torch.int64, device=self.device)
        dist.barrier()                                # nobody starts signalling before everyone has mapped everyone

    def allreduce_adam(self, opt) -> None:
        """flat gradient sum over ranks + Adam on `opt` (a b2048.ddqn.FusedAdam), one kernel."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().p2p_allreduce_adam_f64(
                _ptr(self.peer_grads), _ptr(self.peer_flags), _ptr(self.sync), self.rank, self.world, _ptr(opt.p),
                _ptr(opt.exp_avg), _ptr(opt.exp_avg_sq), _ptr(opt.step_count), opt.p.numel(), opt.lr, opt.betas[0],
                opt.betas[1], opt.eps, _stream(opt.p)), "p2p_allreduce_adam_f64")

    def timed_out(self) -> bool:
        """True if a bounded wait inside the kernel expired (a peer never arrived); synchronises."""
        return bool(self.sync[2].item())

    def check(self) -> None:
        """Raise if an exchange gave up waiting for a peer (the kernel then skipped the update, so this
        rank's replica is intact but behind).  The trainers call this wherever they read a scalar back."""
        if self.timed_out():
            raise RuntimeError(f"rank {self.rank}: NVLink gradient exchange timed out waiting for a peer; "
                               "the update was skipped and the replicas are no longer in step")
