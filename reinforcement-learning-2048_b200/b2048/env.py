"""Batched Board2048 operations on CUDA tensors (thin wrappers over include/b2048.h).

Boards are ``torch.int64`` CUDA tensors holding the packed-u64 bit pattern (nibble 4r+c = tile
exponent).  All functions launch on torch's current stream of the tensors' device and never
synchronise; outputs can be preallocated (``out=`` arguments) for CUDA-graph capture.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

P4_TEN_PERCENT = 0x1999999A      # north_star: 2 at 90 %, 4 at 10 %
P4_FIFTY_PERCENT = 0x80000000    # the reference: np.random.choice([2, 4]) (src/board.py:12,49)

FLAG_LEGAL = 0x0F
FLAG_DONE, FLAG_CHANGED, FLAG_OVERFLOW, FLAG_BADSPAWN = 0x10, 0x20, 0x40, 0x80
SPAWN_NONE = 0xFF
SPAWN_SKIP = 0xFE

_U64 = (1 << 64) - 1


def p4_threshold(p: float) -> int:
    """Probability that a spawned tile is a 4 -> 32-bit threshold."""
    return max(0, min(0xFFFFFFFF, int(round(p * 4294967296.0))))


def _dev(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise _lib.B2048Error("b2048 needs CUDA tensors: there is no CPU fallback")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _stream(t: torch.Tensor):
    return _lib.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return None if t is None else _lib.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, dtype, n=None, name="tensor"):
    if t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    if n is not None and t.numel() != n:
        raise ValueError(f"{name}: expected {n} elements, got {t.numel()}")
    return t


def step(boards, actions, seed=0, step_index=0, index_base=0, p4=P4_TEN_PERCENT, spawn_override=None,
         out=None):
    """One action per board -> (next int64[n], reward int32[n], flags uint8[n]).

    = Board2048.peek_action + merge-score reward + legal mask/done of the input board
    (src/board.py:185-202, src/dqn_lib.py:17-18, 87-88)."""
    n = boards.numel()
    dev = _dev(boards)
    _lib.init(dev)
    _chk(boards, torch.int64, name="boards")
    _chk(actions, torch.uint8, n, "actions")
    if out is None:
        out = (torch.empty_like(boards), torch.empty(n, dtype=torch.int32, device=boards.device),
               torch.empty(n, dtype=torch.uint8, device=boards.device))
    nxt, reward, flags = out
    _chk(nxt, torch.int64, n, "next"); _chk(reward, torch.int32, n, "reward"); _chk(flags, torch.uint8, n, "flags")
    if spawn_override is not None:
        _chk(spawn_override, torch.uint8, n, "spawn_override")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_step(_ptr(boards), _ptr(actions), _ptr(nxt), _ptr(reward), _ptr(flags), n,
                                         seed & _U64, step_index & _U64, index_base & _U64, p4,
                                         _ptr(spawn_override), _stream(boards)), "b2048_step")
    return nxt, reward, flags


def step_all4(boards, seed=0, step_index=0, index_base=0, p4=P4_TEN_PERCENT, spawn_override4=None, out=None):
    """All four successors -> (next4 int64[n,4], reward4 int32[n,4], flags uint8[n]).
    = Board2048.available_moves (src/board.py:138-145)."""
    n = boards.numel()
    dev = _dev(boards)
    _lib.init(dev)
    _chk(boards, torch.int64, name="boards")
    if out is None:
        out = (torch.empty((n, 4), dtype=torch.int64, device=boards.device),
               torch.empty((n, 4), dtype=torch.int32, device=boards.device),
               torch.empty(n, dtype=torch.uint8, device=boards.device))
    nxt, reward, flags = out
    _chk(nxt, torch.int64, 4 * n, "next4"); _chk(reward, torch.int32, 4 * n, "reward4"); _chk(flags, torch.uint8, n, "flags")
    if spawn_override4 is not None:
        _chk(spawn_override4, torch.uint8, 4 * n, "spawn_override4")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_step_all4(_ptr(boards), _ptr(nxt), _ptr(reward), _ptr(flags), n,
                                              seed & _U64, step_index & _U64, index_base & _U64, p4,
                                              _ptr(spawn_override4), _stream(boards)), "b2048_step_all4")
    return nxt, reward, flags


def legal_mask(boards, out=None):
    """flags uint8[n]: bits 0-3 legal [up,down,left,right], bit 4 done (src/board.py:128-135)."""
    n = boards.numel()
    dev = _dev(boards)
    _lib.init(dev)
    _chk(boards, torch.int64, name="boards")
    flags = out if out is not None else torch.empty(n, dtype=torch.uint8, device=boards.device)
    _chk(flags, torch.uint8, n, "flags")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_legal_mask(_ptr(boards), _ptr(flags), n, _stream(boards)), "b2048_legal_mask")
    return flags


def reset(boards, seed=0, step_index=0, index_base=0, p4=P4_TEN_PERCENT, where_flags=None):
    """In place: fresh boards (zeros + two spawns, src/board.py:10-20); with `where_flags` only
    the boards whose DONE bit is set."""
    n = boards.numel()
    dev = _dev(boards)
    _lib.init(dev)
    _chk(boards, torch.int64, name="boards")
    if where_flags is not None:
        _chk(where_flags, torch.uint8, n, "where_flags")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_reset(_ptr(boards), n, seed & _U64, step_index & _U64, index_base & _U64, p4,
                                          _ptr(where_flags), _stream(boards)), "b2048_reset")
    return boards


def spawn(boards, seed=0, step_index=0, index_base=0, p4=P4_TEN_PERCENT, where_flags=None):
    """In place: one new tile per board with an empty cell (= Board2048._populate_empty_cell,
    src/board.py:41-51); with `where_flags` only where the CHANGED bit is set."""
    n = boards.numel()
    dev = _dev(boards)
    _lib.init(dev)
    _chk(boards, torch.int64, name="boards")
    if where_flags is not None:
        _chk(where_flags, torch.uint8, n, "where_flags")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_spawn(_ptr(boards), n, seed & _U64, step_index & _U64, index_base & _U64, p4,
                                          _ptr(where_flags), _stream(boards)), "b2048_spawn")
    return boards


def episode_end(nxt, prev, reward, flags, max_q, ep_score, ep_moves, ep_qsum, totals, qmean_sum, max_tile_hist,
                seed=0, step_index=0, index_base=0, p4=P4_TEN_PERCENT):
    """Fused per-game bookkeeping + auto-reset of finished games (b2048_episode_end)."""
    n = nxt.numel()
    dev = _dev(nxt)
    _lib.init(dev)
    _chk(nxt, torch.int64, n, "next"); _chk(prev, torch.int64, n, "prev"); _chk(reward, torch.int32, n, "reward")
    _chk(flags, torch.uint8, n, "flags"); _chk(ep_score, torch.int64, n, "ep_score")
    _chk(ep_moves, torch.int32, n, "ep_moves"); _chk(totals, torch.int64, 4, "totals")
    _chk(max_tile_hist, torch.int64, 16, "max_tile_hist")
    if max_q is not None:
        _chk(max_q, torch.float64, n, "max_q")
    if ep_qsum is not None:
        _chk(ep_qsum, torch.float64, n, "ep_qsum"); _chk(qmean_sum, torch.float64, 1, "qmean_sum")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_episode_end(_ptr(nxt), _ptr(prev), _ptr(reward), _ptr(flags), _ptr(max_q),
                                                _ptr(ep_score), _ptr(ep_moves), _ptr(ep_qsum), _ptr(totals),
                                                _ptr(qmean_sum), _ptr(max_tile_hist), n, seed & _U64,
                                                step_index & _U64, index_base & _U64, p4, _stream(nxt)),
                   "b2048_episode_end")


def new_boards(n, device="cuda", **kw):
    return reset(torch.empty(n, dtype=torch.int64, device=device), **kw)


def pack(tiles, check=True):
    """int64 tile values [n,16] / [n,4,4] (reference `state`) -> packed boards int64[n]."""
    tiles = tiles.contiguous()
    n = tiles.numel() // 16
    dev = _dev(tiles)
    _lib.init(dev)
    _chk(tiles, torch.int64, 16 * n, "tiles")
    boards = torch.empty(n, dtype=torch.int64, device=tiles.device)
    bad = torch.empty(n, dtype=torch.uint8, device=tiles.device) if check else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_pack(_ptr(tiles), _ptr(boards), _ptr(bad), n, _stream(tiles)), "b2048_pack")
    if check and bool(bad.any()):
        raise ValueError("pack: tiles must be 0 or powers of two in 2..32768 (4-bit exponents)")
    return boards


def unpack_tiles(boards):
    n = boards.numel()
    dev = _dev(boards)
    _lib.init(dev)
    _chk(boards, torch.int64, name="boards")
    out = torch.empty((n, 16), dtype=torch.int64, device=boards.device)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_unpack_tiles(_ptr(boards), _ptr(out), n, _stream(boards)), "b2048_unpack_tiles")
    return out


def unpack_f64(boards, out=None, conv=False):
    """Network input: exponents as float64 [n,16] (dense) or [n,1,4,4] (conv) — the same bytes
    (src/board.py:224-237, src/dqn_lib.py:8-13)."""
    n = boards.numel()
    dev = _dev(boards)
    _lib.init(dev)
    _chk(boards, torch.int64, name="boards")
    if out is None:
        out = torch.empty((n, 16), dtype=torch.float64, device=boards.device)
    _chk(out, torch.float64, 16 * n, "out")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_unpack_f64(_ptr(boards), _ptr(out), n, _stream(boards)), "b2048_unpack_f64")
    return out.view(n, 1, 4, 4) if conv else out.view(n, 16)


def random_boards(n, seed=2048, index_base=0, p_empty=0.3, max_exp=11, device="cuda", out=None):
    """Synthetic boards of SURVEY.md §8(d): cell empty w.p. p_empty else exponent uniform 1..max_exp."""
    boards = out if out is not None else torch.empty(n, dtype=torch.int64, device=device)
    dev = _dev(boards)
    _lib.init(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_random_boards(_ptr(boards), n, seed & _U64, index_base & _U64,
                                                  p4_threshold(p_empty), max_exp, _stream(boards)),
                   "b2048_random_boards")
    return boards


def random_actions(n, seed=2050, step_index=0, index_base=0, device="cuda", out=None):
    actions = out if out is not None else torch.empty(n, dtype=torch.uint8, device=device)
    dev = _dev(actions)
    _lib.init(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b2048_random_actions(_ptr(actions), n, seed & _U64, step_index & _U64,
                                                   index_base & _U64, _stream(actions)), "b2048_random_actions")
    return actions


def steady_state_boards(n, seed=2049, steps=64, index_base=0, p4=P4_TEN_PERCENT, device="cuda", chunk=1 << 23):
    """The second synthetic distribution of SURVEY.md §8(d), "rollout-steady-state": boards after `steps`
    uniformly random LEGAL moves from reset (dead boards restart).  Built on the device in chunks: legal mask
    -> argmax of random scores over the legal moves -> env step -> masked reset.  Deterministic in
    (seed, index_base).  Benchmark-input generator only (torch ops for the action choice)."""
    boards = torch.empty(n, dtype=torch.int64, device=device)
    gen = torch.Generator(device=boards.device)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        base = index_base + c0
        gen.manual_seed((seed * 1000003 + base) & 0x7FFFFFFFFFFFFFFF)
        b = new_boards(m, device=boards.device, seed=seed, step_index=0, index_base=base, p4=p4)
        nxt, rew, flg = torch.empty_like(b), torch.empty(m, dtype=torch.int32, device=b.device), \
            torch.empty(m, dtype=torch.uint8, device=b.device)
        legal = torch.empty(m, dtype=torch.uint8, device=b.device)
        bits = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, device=b.device)
        for t in range(1, steps + 1):
            legal_mask(b, out=legal)
            score = torch.rand((m, 4), dtype=torch.float32, device=b.device, generator=gen)
            score.masked_fill_((legal[:, None] & bits) == 0, -1.0)
            acts = score.argmax(dim=1).to(torch.uint8)
            step(b, acts, seed=seed, step_index=t, index_base=base, p4=p4, out=(nxt, rew, flg))
            reset(nxt, seed=seed ^ 0x4E57, step_index=t, index_base=base, p4=p4, where_flags=flg)
            b, nxt = nxt, b
        boards[c0:c0 + m] = b
    return boards


def step_host(boards, actions, nxt, reward, flags, seed=0, step_index=0, index_base=0, p4=P4_TEN_PERCENT,
              spawn_override=None, device=0):
    """b2048_step with HOST buffers (numpy arrays or CPU tensors, ideally pinned): the library
    chunks the batch and overlaps H2D, kernel and D2H.  Blocks until the results are in `nxt`,
    `reward`, `flags`."""
    _lib.init(device)

    def hp(a):
        if a is None:
            return None
        if isinstance(a, np.ndarray):
            return _lib.c_void_p(a.ctypes.data)
        return _lib.c_void_p(a.data_ptr())

    n = int(boards.shape[0])
    _lib.check(_lib.lib().b2048_step_host(hp(boards), hp(actions), hp(nxt), hp(reward), hp(flags), n,
                                          seed & _U64, step_index & _U64, index_base & _U64, p4,
                                          hp(spawn_override), device), "b2048_step_host")
    return nxt, reward, flags


class PinnedBuffer:
    """NUMA-local pinned host memory (b2048_host_alloc) viewed as a numpy array / CPU tensor.  Keep the
    object alive as long as the views are used; `free()` (or garbage collection) releases the memory."""

    def __init__(self, n: int, dtype, device: int = 0):
        import ctypes
        self.dtype = np.dtype(dtype)
        self.nbytes = int(n) * self.dtype.itemsize
        ptr, node, bound = _lib.c_void_p(), ctypes.c_int(-1), ctypes.c_int(0)
        _lib.check(_lib.lib().b2048_host_alloc(ctypes.byref(ptr), max(self.nbytes, 1), int(device), ctypes.byref(node),
                                               ctypes.byref(bound)), "b2048_host_alloc")
        self.ptr, self.numa_node, self.bound = int(ptr.value), int(node.value), bool(bound.value)
        raw = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=self.dtype, count=int(n))
        self.tensor = torch.from_numpy(self.array)

    def free(self) -> None:
        if self.ptr:
            self.array = self.tensor = None
            _lib.lib().b2048_host_free(_lib.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def bind_thread_near(device: int = 0) -> bool:
    """Bind the calling thread to the CPUs next to `device`; False if the topology is unknown."""
    return _lib.lib().b2048_bind_thread_near(int(device)) == 0


def row_lut_host() -> np.ndarray:
    """The 65536-entry row table as built on the host (no GPU needed)."""
    out = np.zeros(65536, dtype=np.uint32)
    _lib.check(_lib.lib().b2048_copy_row_lut_host(_lib.c_void_p(out.ctypes.data)), "b2048_copy_row_lut_host")
    return out
