"""Single-board engine behind the drop-in `board.Board2048` shim.

Every Board2048 operation that the reference computes with numpy loops (slide/merge, spawn, legal
mask) is one launch of the batched CUDA kernels with n = 1 (or 4): tile values go up as 16 int64,
are packed on the device, stepped, unpacked on the device and come back as 16 int64.  There is no
CPU path; constructing the engine without a GPU raises.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from . import _lib, env


class SingleBoardEngine:
    def __init__(self, device=None, seed=None, p_four=0.5):
        if not torch.cuda.is_available():
            raise _lib.B2048Error("Board2048 needs a CUDA device: this framework has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:0")
        _lib.init(self.device.index or 0)
        # the reference draws the spawn from Python's `random` / numpy's global RNG; seeding
        # `random` before the first board therefore still makes a run reproducible here
        self.seed = seed if seed is not None else random.getrandbits(64)
        self.counter = 0
        self.p4 = env.p4_threshold(p_four)
        kw = dict(device=self.device)
        self.d_tiles = torch.zeros((4, 16), dtype=torch.int64, **kw)
        self.d_actions = torch.zeros(4, dtype=torch.uint8, **kw)
        self.d_override = torch.full((4,), env.SPAWN_SKIP, dtype=torch.uint8, **kw)
        self.h_tiles = torch.zeros((4, 16), dtype=torch.int64).pin_memory()
        self.h_out = torch.zeros((4, 16), dtype=torch.int64).pin_memory()
        self.h_reward = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.h_flags = torch.zeros(4, dtype=torch.uint8).pin_memory()

    def reseed(self, seed: int) -> None:
        self.seed, self.counter = int(seed), 0

    def set_p_four(self, p: float) -> None:
        self.p4 = env.p4_threshold(p)

    def _tick(self) -> int:
        self.counter += 1
        return self.counter

    def _upload(self, state) -> torch.Tensor:
        s = np.asarray(state)
        if s.shape != (4, 4):
            raise NotImplementedError("the CUDA board is 4x4 (packed 64-bit); other sizes are not supported")
        self.h_tiles[0].copy_(torch.from_numpy(np.ascontiguousarray(s, dtype=np.int64).reshape(16)))
        self.d_tiles[0].copy_(self.h_tiles[0], non_blocking=True)
        return env.pack(self.d_tiles[:1])          # raises ValueError on non power-of-two tiles

    def _download(self, boards, reward=None, flags=None, n=1):
        self.h_out[:n].copy_(env.unpack_tiles(boards), non_blocking=True)
        if reward is not None:
            self.h_reward[:n].copy_(reward, non_blocking=True)
        if flags is not None:
            self.h_flags[:n].copy_(flags, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.h_out[:n].numpy().reshape(n, 4, 4).copy()

    def move(self, state, action: int, spawn: bool = True):
        """-> (next state int[4,4], reward int, flags int)."""
        with torch.cuda.device(self.device):
            b = self._upload(state)
            self.d_actions[0] = int(action) & 3
            nxt, rew, flg = env.step(b, self.d_actions[:1], seed=self.seed, step_index=self._tick(), p4=self.p4,
                                     spawn_override=None if spawn else self.d_override[:1])
            out = self._download(nxt, rew, flg)
        return out[0], int(self.h_reward[0]), int(self.h_flags[0])

    def all4(self, state):
        """-> (next states int[4,4,4], rewards int[4], flags int) with a spawn in every changed successor."""
        with torch.cuda.device(self.device):
            b = self._upload(state)
            nxt4, rew4, flg = env.step_all4(b, seed=self.seed, step_index=self._tick(), p4=self.p4)
            out = self._download(nxt4.reshape(4), rew4.reshape(4), flg, n=4)
        return out, self.h_reward[:4].numpy().copy(), int(self.h_flags[0])

    def legal(self, state) -> int:
        with torch.cuda.device(self.device):
            b = self._upload(state)
            flg = env.legal_mask(b)
            self.h_flags[:1].copy_(flg, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return int(self.h_flags[0])

    def spawn(self, state):
        with torch.cuda.device(self.device):
            b = self._upload(state)
            env.spawn(b, seed=self.seed, step_index=self._tick(), p4=self.p4)
            return self._download(b)[0]

    def fresh(self):
        with torch.cuda.device(self.device):
            b = env.new_boards(1, device=self.device, seed=self.seed, step_index=self._tick(), p4=self.p4)
            return self._download(b)[0]


_engine = None


def engine() -> SingleBoardEngine:
    global _engine
    if _engine is None:
        _engine = SingleBoardEngine()
    return _engine
