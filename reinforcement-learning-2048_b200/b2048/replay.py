"""GPU-resident ring replay buffer (K2) — the deque(maxlen) of src/dqn_lib.py:172 as packed
struct-of-arrays tensors with device-side head/size counters (CUDA-graph friendly)."""
from __future__ import annotations

import ctypes

import numpy as np

import torch

from . import _lib
from .env import _chk, _dev, _ptr, _stream, _U64


class ReplayRing:
    """Ring of `capacity` transitions (state, action, reward, next_state, done) on one GPU.

    Logical index 0 is the oldest entry, exactly like indexing the reference's deque, so the
    reference's ``np.random.randint(len(buf), size=B)`` draw can be replayed via `idx_override`.
    """

    def __init__(self, capacity: int, device="cuda"):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.init(dev)
        self._dev = dev
        kw = dict(device=self.device)
        self.s = torch.zeros(self.capacity, dtype=torch.int64, **kw)
        self.s2 = torch.zeros(self.capacity, dtype=torch.int64, **kw)
        self.r = torch.zeros(self.capacity, dtype=torch.int32, **kw)
        self.a = torch.zeros(self.capacity, dtype=torch.uint8, **kw)
        self.d = torch.zeros(self.capacity, dtype=torch.uint8, **kw)
        self.head_size = torch.zeros(4, dtype=torch.int64, **kw)   # head, size, auto sample counter, reserved
        self._ring = _lib.Ring(self.s.data_ptr(), self.s2.data_ptr(), self.r.data_ptr(), self.a.data_ptr(),
                               self.d.data_ptr(), self.head_size.data_ptr(), self.capacity)

    def __len__(self) -> int:          # synchronises (reads the device counter)
        return int(self.head_size[1].item())

    def append(self, s, a, r, s2, done_flags, done_is_bool: bool = False) -> None:
        """Append n transitions (src/dqn_lib.py:106).  `done_flags`: the step kernel's flags bytes
        (done = byte & 0x10); pass done_is_bool=True for a 0/1 array."""
        n = s.numel()
        if done_is_bool:
            done_flags = (done_flags != 0).to(torch.uint8) * 0x10
        _chk(s, torch.int64, name="s"); _chk(s2, torch.int64, n, "s2"); _chk(a, torch.uint8, n, "a")
        _chk(r, torch.int32, n, "r"); _chk(done_flags, torch.uint8, n, "done")
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().replay_append(ctypes.byref(self._ring), _ptr(s), _ptr(a), _ptr(r), _ptr(s2),
                                                _ptr(done_flags), n, _stream(s)), "replay_append")

    CTR_AUTO = (1 << 64) - 1   # use (and bump) the ring's device-side counter: CUDA-graph friendly

    def sample(self, batch_size: int, seed=2051, ctr=0, idx_override=None, out=None, return_idx=False):
        """Fused sample + gather + unpack (src/dqn_lib.py:33-84).
        -> states f64[B,16], actions i64[B], rewards i64[B], next_states f64[B,16], dones i64[B]
        (reference order: states, actions, rewards, next_states, dones)."""
        B = int(batch_size)
        kw = dict(device=self.device)
        if out is None:
            out = (torch.empty((B, 16), dtype=torch.float64, **kw), torch.empty(B, dtype=torch.int64, **kw),
                   torch.empty(B, dtype=torch.int64, **kw), torch.empty((B, 16), dtype=torch.float64, **kw),
                   torch.empty(B, dtype=torch.int64, **kw))
        states, actions, rewards, next_states, dones = out
        idx_out = torch.empty(B, dtype=torch.int64, **kw) if return_idx else None
        if idx_override is not None:
            _chk(idx_override, torch.int64, B, "idx_override")
        with torch.cuda.device(self._dev):
            _lib.check(_lib.lib().replay_sample(ctypes.byref(self._ring), B, seed & _U64, ctr & _U64,
                                                _ptr(idx_override), _ptr(states), _ptr(next_states),
                                                _ptr(actions), _ptr(rewards), _ptr(dones), _ptr(idx_out),
                                                _stream(states)), "replay_sample")
        if return_idx:
            return states, actions, rewards, next_states, dones, idx_out
        return states, actions, rewards, next_states, dones

    def clear(self) -> None:
        self.head_size.zero_()


class ReplayDeque:
    """`collections.deque(maxlen=...)` look-alike for `dqn_lib` whose contents live in a GPU ring.

    ``append((board, action, reward, next_board, done))`` takes the reference's 5-tuples
    (src/dqn_lib.py:106); boards are staged as tile values in pinned host memory and flushed to the
    device ring in batches (packed on the GPU), so sampling never touches Python objects.
    ``len()`` and the oldest-first logical indexing match the deque the reference uses."""

    STAGE = 2048

    def __init__(self, iterable=(), maxlen=None, device="cuda"):
        if maxlen is None:
            raise ValueError("ReplayDeque needs a maxlen (the replay_buffer_length)")
        self.maxlen = int(maxlen)
        self.ring = ReplayRing(self.maxlen, device=device)
        self._n_total = 0
        self._h_s = torch.zeros((self.STAGE, 16), dtype=torch.int64).pin_memory()
        self._h_s2 = torch.zeros((self.STAGE, 16), dtype=torch.int64).pin_memory()
        self._h_a = torch.zeros(self.STAGE, dtype=torch.uint8).pin_memory()
        self._h_r = torch.zeros(self.STAGE, dtype=torch.int32).pin_memory()
        self._h_d = torch.zeros(self.STAGE, dtype=torch.uint8).pin_memory()
        self._staged = 0
        for item in iterable:
            self.append(item)

    def __len__(self) -> int:
        return min(self._n_total, self.maxlen)

    def append(self, experience) -> None:
        board, action, reward, next_board, done = experience
        i = self._staged
        self._h_s[i] = torch.from_numpy(np.ascontiguousarray(board.state, dtype=np.int64).reshape(16))
        self._h_s2[i] = torch.from_numpy(np.ascontiguousarray(next_board.state, dtype=np.int64).reshape(16))
        self._h_a[i] = int(action)
        self._h_r[i] = int(reward)
        self._h_d[i] = 0x10 if bool(done) else 0            # B2048_FLAG_DONE
        self._staged += 1
        self._n_total += 1
        if self._staged == self.STAGE:
            self.flush()

    def flush(self) -> None:
        """Upload the staged transitions: H2D of tile values, pack on the GPU, ring append."""
        n = self._staged
        if n == 0:
            return
        from . import env
        dev = self.ring.device
        with torch.cuda.device(dev):
            s = env.pack(self._h_s[:n].to(dev, non_blocking=True))
            s2 = env.pack(self._h_s2[:n].to(dev, non_blocking=True))
            self.ring.append(s, self._h_a[:n].to(dev, non_blocking=True), self._h_r[:n].to(dev, non_blocking=True),
                             s2, self._h_d[:n].to(dev, non_blocking=True))
            torch.cuda.current_stream(dev).synchronize()     # staging buffers are reused
        self._staged = 0

    def sample(self, batch_size, indices=None, **kw):
        self.flush()
        idx = None
        if indices is not None:
            idx = torch.as_tensor(np.asarray(indices, dtype=np.int64)).to(self.ring.device)
        return self.ring.sample(batch_size, idx_override=idx, **kw)
