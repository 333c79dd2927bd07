"""Batched rollouts: thousands to millions of concurrent Board2048 games on one GPU.

The reference plays one game at a time (src/dqn_lib.py:174-205, src/player.py:40-64); here every
step of `VectorEnv` advances all boards with one launch each of: legal mask -> (Q-network forward)
-> batched epsilon-greedy -> env step -> replay append -> masked reset of finished games.  Episode
statistics (the fields the reference passes to Experiment.add_episode, src/experiments.py:112-122)
are accumulated per board on the device.
"""
from __future__ import annotations

import torch

from . import ddqn, env
from .replay import ReplayRing


class VectorEnv:
    def __init__(self, n: int, device="cuda", seed: int = 0, index_base: int = 0, p_four: float = 0.1,
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
        # per-game accumulators
        self.ep_score = torch.zeros(self.n, dtype=torch.int64, **kw)
        self.ep_moves = torch.zeros(self.n, dtype=torch.int64, **kw)
        self.ep_qsum = torch.zeros(self.n, dtype=torch.float64, **kw)
        # totals over finished games
        self.finished = torch.zeros((), dtype=torch.int64, **kw)
        self.sum_score = torch.zeros((), dtype=torch.int64, **kw)
        self.sum_moves = torch.zeros((), dtype=torch.int64, **kw)
        self.max_tile_hist = torch.zeros(16, dtype=torch.int64, **kw)

    def observe(self) -> torch.Tensor:
        """Network input of the current boards: float64 [n,1,4,4] (conv) or [n,16] (dense)."""
        env.unpack_f64(self.boards, out=self.obs)
        return self.obs.view(self.n, 1, 4, 4) if self.conv else self.obs

    @torch.no_grad()
    def step(self, model=None, epsilon: float = 1.0, replay: ReplayRing | None = None):
        """One step of every game.  model=None or epsilon>=1 plays the uniformly random policy of
        the reference's epsilon branch (illegal no-op moves included, src/dqn_lib.py:20-21)."""
        self.t += 1
        env.legal_mask(self.boards, out=self.legal)
        if model is not None and epsilon < 1.0:
            q = model(self.observe()).contiguous()
            ddqn.egreedy_select(q, self.legal, epsilon, seed=self.seed ^ 0x5EED, ctr=self.t,
                                index_base=self.index_base, out=(self.actions, self.max_q))
            self.ep_qsum += self.max_q
        else:
            env.random_actions(self.n, seed=self.seed ^ 0xAC71, step_index=self.t, index_base=self.index_base,
                               out=self.actions)
        env.step(self.boards, self.actions, seed=self.seed, step_index=self.t, index_base=self.index_base,
                 p4=self.p4, out=(self.next_boards, self.reward, self.flags))
        if replay is not None:
            replay.append(self.boards, self.actions, self.reward, self.next_boards, self.flags)
        done = (self.flags & env.FLAG_DONE) != 0
        self.ep_score += self.reward
        self.ep_moves += 1
        # finished games: fold their statistics into the totals, then start fresh boards
        nd = done.sum()
        self.finished += nd
        self.sum_score += torch.where(done, self.ep_score, 0).sum()
        self.sum_moves += torch.where(done, self.ep_moves, 0).sum()
        shifts = 4 * torch.arange(16, device=self.device, dtype=torch.int64)
        max_exp = ((self.boards[:, None] >> shifts) & 0xF).amax(dim=1)
        self.max_tile_hist += torch.bincount(max_exp[done], minlength=16)[:16]
        self.ep_score.masked_fill_(done, 0)
        self.ep_moves.masked_fill_(done, 0)
        self.ep_qsum.masked_fill_(done, 0.0)
        self.boards, self.next_boards = self.next_boards, self.boards
        env.reset(self.boards, seed=self.seed ^ 0x4E57, step_index=self.t, index_base=self.index_base, p4=self.p4,
                  where_flags=self.flags)
        return self.flags

    def stats(self) -> dict:
        f = int(self.finished.item())
        hist = {int(2 ** e): int(c) for e, c in enumerate(self.max_tile_hist.tolist()) if c}
        return {"games": f, "mean_merge_score": float(self.sum_score.item()) / max(f, 1),
                "mean_moves": float(self.sum_moves.item()) / max(f, 1), "max_tile_hist": hist, "steps": self.t * self.n}
