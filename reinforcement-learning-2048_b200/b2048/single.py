"""Single-board engine behind the drop-in `board.Board2048` shim.

Every Board2048 operation that the reference computes with numpy loops (slide/merge, spawn, legal
mask) is ONE kernel launch through `b2048_board_host`: the tile values travel as a kernel argument, the
kernel packs, steps (same arithmetic and Philox lanes as the batched kernels with n = 1), unpacks, and
writes the result to mapped pinned memory.  There is no CPU path; constructing the engine without a GPU
raises.
"""
from __future__ import annotations

import ctypes
import random

import numpy as np
import torch

from . import _lib, env


OP_MOVE, OP_ALL4, OP_LEGAL, OP_SPAWN, OP_FRESH = range(5)


class SingleBoardEngine:
    """One launch + one stream synchronisation per Board2048 operation (`b2048_board_host`): the 16 tile values
    travel as a kernel argument, the result comes back through mapped pinned memory."""

    def __init__(self, device=None, seed=None, p_four=0.5):
        if not torch.cuda.is_available():
            raise _lib.B2048Error("Board2048 needs a CUDA device: this framework has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:0")
        self.dev_index = self.device.index or 0
        _lib.init(self.dev_index)
        # the reference draws the spawn from Python's `random` / numpy's global RNG; seeding
        # `random` before the first board therefore still makes a run reproducible here
        self.seed = seed if seed is not None else random.getrandbits(64)
        self.counter = 0
        self.p4 = env.p4_threshold(p_four)
        self._res = _lib.BoardResult()
        self._res_ptr = ctypes.addressof(self._res)
        self._next = np.ctypeslib.as_array(self._res.next)        # int64 [4, 16] view of the result struct
        self._reward = np.ctypeslib.as_array(self._res.reward)
        self._tiles = np.zeros(16, dtype=np.int64)
        self._tiles_ptr = self._tiles.ctypes.data
        self._call = _lib.lib().b2048_board_host

    def reseed(self, seed: int) -> None:
        self.seed, self.counter = int(seed), 0

    def set_p_four(self, p: float) -> None:
        self.p4 = env.p4_threshold(p)

    def _tick(self) -> int:
        self.counter += 1
        return self.counter

    def _run(self, op: int, state, action: int = 0, spawn: bool = True, tick: bool = True) -> None:
        if state is not None:
            s = np.asarray(state)
            if s.shape != (4, 4):
                raise NotImplementedError("the CUDA board is 4x4 (packed 64-bit); other sizes are not supported")
            self._tiles[:] = s.reshape(16)
        step = self._tick() if tick else 0
        _lib.check(self._call(op, self._tiles_ptr, int(action) & 3, 1 if spawn else 0, self.seed & env._U64, step, self.p4,
                              self._res_ptr, self.dev_index), "b2048_board_host")
        if self._res.bad:
            raise ValueError("tiles must be 0 or powers of two in 2..32768")

    def move(self, state, action: int, spawn: bool = True):
        """-> (next state int[4,4], reward int, flags int)."""
        self._run(OP_MOVE, state, action, spawn)
        return self._next[0].reshape(4, 4).copy(), int(self._reward[0]), int(self._res.flags)

    def all4(self, state):
        """-> (next states int[4,4,4], rewards int[4], flags int) with a spawn in every changed successor."""
        self._run(OP_ALL4, state)
        return self._next.reshape(4, 4, 4).copy(), self._reward.copy(), int(self._res.flags)

    def legal(self, state) -> int:
        self._run(OP_LEGAL, state, tick=False)
        return int(self._res.flags)

    def spawn(self, state):
        self._run(OP_SPAWN, state)
        return self._next[0].reshape(4, 4).copy()

    def fresh(self):
        self._run(OP_FRESH, None)
        return self._next[0].reshape(4, 4).copy()


_engine = None


def engine() -> SingleBoardEngine:
    global _engine
    if _engine is None:
        _engine = SingleBoardEngine()
    return _engine
