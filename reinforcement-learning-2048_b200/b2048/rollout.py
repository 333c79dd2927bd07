"""Batched rollouts: thousands to millions of concurrent Board2048 games on one GPU.

The reference plays one game at a time (src/dqn_lib.py:174-205, src/player.py:40-64); here every
step of `VectorEnv` advances all boards with one launch each of: legal mask -> (Q-network forward)
-> batched epsilon-greedy -> env step -> replay append -> masked reset of finished games.  Episode
statistics (the fields the reference passes to Experiment.add_episode, src/experiments.py:112-122)
are accumulated per board on the device.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ddqn, env
from .qfused import FusedConvQ
from .replay import ReplayRing

_U64 = (1 << 64) - 1


class VectorEnv:
    def __init__(self, n: int, device="cuda", seed: int = 0, index_base: int = 0, p_four: float = 0.5,
                 conv: bool = True):
        self.n, self.device, self.seed, self.index_base = int(n), torch.device(device), int(seed), int(index_base)
        self.p4 = env.p4_threshold(p_four)
        self.conv = conv
        self.t = 0                                           # global step counter (Philox counter)
        kw = dict(device=self.device)
        self.boards = env.new_boards(self.n, device=self.device, seed=self.seed, step_index=0,
                                     index_base=self.index_base, p4=self.p4)
        self.next_boards = torch.empty_like(self.boards)
        self.reward = torch.empty(self.n, dtype=torch.int32, **kw)
        self.flags = torch.empty(self.n, dtype=torch.uint8, **kw)
        self.legal = torch.empty(self.n, dtype=torch.uint8, **kw)
        self.actions = torch.empty(self.n, dtype=torch.uint8, **kw)
        self.max_q = torch.empty(self.n, dtype=torch.float64, **kw)
        self.obs = torch.empty((self.n, 16), dtype=torch.float64, **kw)
        self.q = torch.empty((self.n, 4), dtype=torch.float64, **kw)
        # per-game accumulators and totals over finished games (all on the device)
        self.ep_score = torch.zeros(self.n, dtype=torch.int64, **kw)
        self.ep_moves = torch.zeros(self.n, dtype=torch.int32, **kw)
        self.ep_qsum = torch.zeros(self.n, dtype=torch.float64, **kw)
        self.totals = torch.zeros(4, dtype=torch.int64, **kw)          # games, merge-score sum, moves sum
        self.qmean_sum = torch.zeros(1, dtype=torch.float64, **kw)     # sum over games of mean max-Q
        self.max_tile_hist = torch.zeros(16, dtype=torch.int64, **kw)
        # step() is launch-bound on the host below ~1 Mi games (4 096 games: 75 us of Python per step against ~25 us
        # of kernels): it calls the C-ABI directly with raw addresses of these persistent buffers, validated here once,
        # instead of going through the checking wrappers of env.py five times per step
        self._ever_greedy = False
        self._dev = env._dev(self.boards)
        _lib.init(self._dev)
        self._L = _lib.lib()
        self._p = {name: getattr(self, name).data_ptr() for name in
                   ("boards", "next_boards", "reward", "flags", "legal", "actions", "max_q", "q", "ep_score", "ep_moves",
                    "ep_qsum", "totals", "qmean_sum", "max_tile_hist")}

    def _rebind(self) -> None:
        """Re-validate and re-read the buffer addresses after a caller assigned new tensors."""
        for name, dtype in (("boards", torch.int64), ("next_boards", torch.int64)):
            env._chk(getattr(self, name), dtype, self.n, name)
            if env._dev(getattr(self, name)) != self._dev:
                raise ValueError(f"{name} moved to another device")
        for name in self._p:
            self._p[name] = getattr(self, name).data_ptr()

    def observe(self) -> torch.Tensor:
        """Network input of the current boards: float64 [n,1,4,4] (conv) or [n,16] (dense)."""
        env.unpack_f64(self.boards, out=self.obs)
        return self.obs.view(self.n, 1, 4, 4) if self.conv else self.obs

    @torch.no_grad()
    def step(self, model=None, epsilon: float = 1.0, replay: ReplayRing | None = None):
        """One step of every game.  model=None or epsilon>=1 plays the uniformly random policy of
        the reference's epsilon branch (illegal no-op moves included, src/dqn_lib.py:20-21)."""
        if torch.cuda.current_device() != self._dev:
            with torch.cuda.device(self._dev):
                return self._step(model, epsilon, replay)
        return self._step(model, epsilon, replay)

    def _step(self, model, epsilon, replay):
        self.t += 1
        L, p, n, t, base, chk = self._L, self._p, self.n, self.t, self.index_base & _U64, _lib.check
        if p["boards"] != self.boards.data_ptr() or p["next_boards"] != self.next_boards.data_ptr():
            self._rebind()                                   # a caller replaced the board tensors (e.g. a restored checkpoint)
        st = torch.cuda.current_stream(self.device).cuda_stream
        greedy = model is not None and epsilon < 1.0
        if greedy:
            chk(L.b2048_legal_mask(p["boards"], p["legal"], n, st), "b2048_legal_mask")
            if isinstance(model, FusedConvQ) and model.device == self.boards.device:
                model.forward_boards_raw(p["boards"], p["q"], n, st)      # one fused kernel straight from the packed boards
                q_ptr = p["q"]
            else:
                q = model(self.observe()).contiguous()
                if q.dtype != torch.float64 or q.numel() != 4 * n or q.device != self.boards.device:
                    raise ValueError("the model must return float64 Q-values [n, 4] on the environment's device")
                q_ptr = q.data_ptr()
            chk(L.egreedy_select(q_ptr, p["legal"], float(epsilon), (self.seed ^ 0x5EED) & _U64, t, base, None,
                                 p["actions"], p["max_q"], n, st), "egreedy_select")
        else:
            chk(L.b2048_random_actions(p["actions"], n, (self.seed ^ 0xAC71) & _U64, t, base, st), "b2048_random_actions")
        chk(L.b2048_step(p["boards"], p["actions"], p["next_boards"], p["reward"], p["flags"], n, self.seed & _U64, t, base,
                         self.p4, None, st), "b2048_step")
        if replay is not None:
            if replay._dev != self._dev:
                raise ValueError("the replay ring lives on another device")
            chk(L.replay_append(ctypes.byref(replay._ring), p["boards"], p["actions"], p["reward"], p["next_boards"],
                                p["flags"], n, st), "replay_append")
        # one fused pass: per-game accumulators, totals of finished games, fresh boards for them
        # (the max-Q accumulators stay out of the pass until the first model-driven step: 16 of 45 bytes per game)
        self._ever_greedy = self._ever_greedy or greedy
        chk(L.b2048_episode_end(p["next_boards"], p["boards"], p["reward"], p["flags"], p["max_q"] if greedy else None,
                                p["ep_score"], p["ep_moves"], p["ep_qsum"] if self._ever_greedy else None, p["totals"],
                                p["qmean_sum"] if self._ever_greedy else None, p["max_tile_hist"], n,
                                (self.seed ^ 0x4E57) & _U64, t, base, self.p4, st), "b2048_episode_end")
        self.boards, self.next_boards = self.next_boards, self.boards
        p["boards"], p["next_boards"] = p["next_boards"], p["boards"]
        return self.flags

    def stats(self) -> dict:
        games, score, moves, _ = self.totals.tolist()
        hist = {int(2 ** e): int(c) for e, c in enumerate(self.max_tile_hist.tolist()) if c}
        return {"games": games, "mean_merge_score": score / max(games, 1), "mean_moves": moves / max(games, 1),
                "mean_max_q": float(self.qmean_sum.item()) / max(games, 1), "max_tile_hist": hist,
                "steps": self.t * self.n}
