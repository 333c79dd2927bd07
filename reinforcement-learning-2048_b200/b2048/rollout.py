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
from .qfused import FusedConvQ
from .replay import ReplayRing


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

    def observe(self) -> torch.Tensor:
        """Network input of the current boards: float64 [n,1,4,4] (conv) or [n,16] (dense)."""
        env.unpack_f64(self.boards, out=self.obs)
        return self.obs.view(self.n, 1, 4, 4) if self.conv else self.obs

    @torch.no_grad()
    def step(self, model=None, epsilon: float = 1.0, replay: ReplayRing | None = None):
        """One step of every game.  model=None or epsilon>=1 plays the uniformly random policy of
        the reference's epsilon branch (illegal no-op moves included, src/dqn_lib.py:20-21)."""
        self.t += 1
        greedy = model is not None and epsilon < 1.0
        if greedy:
            env.legal_mask(self.boards, out=self.legal)
            if isinstance(model, FusedConvQ):                # one fused kernel straight from the packed boards
                q = model.forward_boards(self.boards, out=self.q)
            else:
                q = model(self.observe()).contiguous()
            ddqn.egreedy_select(q, self.legal, epsilon, seed=self.seed ^ 0x5EED, ctr=self.t,
                                index_base=self.index_base, out=(self.actions, self.max_q))
        else:
            env.random_actions(self.n, seed=self.seed ^ 0xAC71, step_index=self.t, index_base=self.index_base,
                               out=self.actions)
        env.step(self.boards, self.actions, seed=self.seed, step_index=self.t, index_base=self.index_base,
                 p4=self.p4, out=(self.next_boards, self.reward, self.flags))
        if replay is not None:
            replay.append(self.boards, self.actions, self.reward, self.next_boards, self.flags)
        # one fused pass: per-game accumulators, totals of finished games, fresh boards for them
        env.episode_end(self.next_boards, self.boards, self.reward, self.flags, self.max_q if greedy else None,
                        self.ep_score, self.ep_moves, self.ep_qsum, self.totals, self.qmean_sum,
                        self.max_tile_hist, seed=self.seed ^ 0x4E57, step_index=self.t, index_base=self.index_base,
                        p4=self.p4)
        self.boards, self.next_boards = self.next_boards, self.boards
        return self.flags

    def stats(self) -> dict:
        games, score, moves, _ = self.totals.tolist()
        hist = {int(2 ** e): int(c) for e, c in enumerate(self.max_tile_hist.tolist()) if c}
        return {"games": games, "mean_merge_score": score / max(games, 1), "mean_moves": moves / max(games, 1),
                "mean_max_q": float(self.qmean_sum.item()) / max(games, 1), "max_tile_hist": hist,
                "steps": self.t * self.n}
