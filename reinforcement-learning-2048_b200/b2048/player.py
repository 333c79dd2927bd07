"""Batched players: the reference's `player.py` policies over many concurrent games.

`Player.play_game(random_policy=True)` picks uniformly among the LEGAL moves (argmax of
mask * rand, src/player.py:46-57 — unlike the epsilon branch of dqn_lib, which ignores legality)
and `basic_upleft_algorithm` cycles up, left and falls back to down, right when neither changed the
board (src/player.py:66-84).  Both run here on thousands of boards at once with the CUDA env step;
per-game statistics are accumulated on the device by the fused episode kernel.
"""
from __future__ import annotations

import torch

from . import env
from .qfused import FusedConvQ, accelerate_inference


class BatchedPlayer:
    def __init__(self, n_envs: int = 1 << 16, device="cuda", seed: int = 0, p_four: float = 0.5):
        self.n, self.device, self.seed = int(n_envs), torch.device(device), int(seed)
        self.p4 = env.p4_threshold(p_four)
        kw = dict(device=self.device)
        self.boards = env.new_boards(self.n, device=self.device, seed=self.seed, step_index=0, p4=self.p4)
        self.next = torch.empty_like(self.boards)
        self.reward = torch.empty(self.n, dtype=torch.int32, **kw)
        self.flags = torch.empty(self.n, dtype=torch.uint8, **kw)
        self.legal = torch.empty(self.n, dtype=torch.uint8, **kw)
        self.ep_score = torch.zeros(self.n, dtype=torch.int64, **kw)
        self.ep_moves = torch.zeros(self.n, dtype=torch.int32, **kw)
        self.totals = torch.zeros(4, dtype=torch.int64, **kw)
        self.hist = torch.zeros(16, dtype=torch.int64, **kw)
        self.t = 0
        self.games_done = torch.zeros(self.n, dtype=torch.int64, **kw)
        self._bits = torch.tensor([1, 2, 4, 8], dtype=torch.uint8, **kw)

    def _finish(self, done_flags, quota: int):
        # every board plays exactly `quota` complete games (stopping at a total count instead would
        # over-represent short games); boards that are through keep stepping but are not counted
        active = self.games_done < quota
        done = ((done_flags & env.FLAG_DONE) != 0) & active
        self.games_done += done
        done_flags = torch.where(done, done_flags | env.FLAG_DONE, done_flags & ~env.FLAG_DONE & 0xFF)
        env.episode_end(self.next, self.boards, self.reward, done_flags, None, self.ep_score, self.ep_moves, None,
                        self.totals, None, self.hist, seed=self.seed ^ 0x4E57, step_index=self.t, p4=self.p4)
        self.boards, self.next = self.next, self.boards

    def _stats(self) -> dict:
        games, score, moves, _ = self.totals.tolist()
        return {"games": games, "mean_merge_score": score / max(games, 1), "mean_moves": moves / max(games, 1),
                "max_tile_hist": {int(2 ** e): int(c) for e, c in enumerate(self.hist.tolist()) if c}}

    @torch.no_grad()
    def random_baseline(self, n_games: int) -> dict:
        """Uniformly random LEGAL moves until `n_games` games have ended (src/player.py:40-64).
        Like the reference, the dead board gets one final no-op move before the game is closed."""
        g = torch.Generator(device=self.device).manual_seed(self.seed)
        quota = -(-n_games // self.n)
        while int(self.games_done.min().item()) < quota:
            for _ in range(32):
                self.t += 1
                env.legal_mask(self.boards, out=self.legal)
                q = torch.rand((self.n, 4), device=self.device, generator=g)
                mask = (self.legal[:, None] & self._bits) != 0
                actions = torch.argmax(q * mask, dim=1).to(torch.uint8)      # dead board -> action 0, a no-op
                env.step(self.boards, actions, seed=self.seed, step_index=self.t, p4=self.p4,
                         out=(self.next, self.reward, self.flags))
                self._finish(self.flags, quota)
        return self._stats()

    def _observe(self, scaling: str, conv: bool) -> torch.Tensor:
        if scaling == "log2":                                    # what training feeds the network (dqn_lib.py:8-13)
            x = env.unpack_f64(self.boards)
        elif scaling == "normalized":                            # what player.py feeds it (board.py:218-222)
            t = env.unpack_tiles(self.boards).to(torch.float64)
            x = t / t.max(dim=1, keepdim=True).values
        else:
            raise ValueError(f"unknown input scaling {scaling!r}")
        return x.view(self.n, 1, 4, 4) if conv else x

    @torch.no_grad()
    def _model_actions(self, net, scaling: str, conv: bool) -> torch.Tensor:
        env.legal_mask(self.boards, out=self.legal)
        if isinstance(net, FusedConvQ):
            q = net.forward_boards(self.boards, scaling=scaling)
        else:
            q = net(self._observe(scaling, conv))
        mask = ((self.legal[:, None] & self._bits) != 0).to(q.dtype)
        return torch.argmax(mask * q, dim=1).to(torch.uint8)   # dead board -> action 0, a no-op

    @torch.no_grad()
    def model_policy(self, model: torch.nn.Module, n_games: int, scaling: str = "normalized", conv: bool = True) -> dict:
        """Greedy play with a Q-network: action = argmax(legal_mask * Q(board)) with first-index ties
        (src/player.py:48-53), every board until it has finished its share of `n_games` games."""
        net = model if isinstance(model, FusedConvQ) else accelerate_inference(model)
        stalled = torch.zeros(self.n, dtype=torch.int64, device=self.device)
        quota = -(-n_games // self.n)
        while int(self.games_done.min().item()) < quota:
            for _ in range(32):
                self.t += 1
                actions = self._model_actions(net, scaling, conv)
                env.step(self.boards, actions, seed=self.seed, step_index=self.t, p4=self.p4,
                         out=(self.next, self.reward, self.flags))
                # argmax(mask * Q) picks an ILLEGAL move (value 0) when every legal Q is negative; the
                # policy is deterministic, so such a game never moves again — the reference loops
                # forever there (src/player.py:43-59).  Close it and count it.
                stuck = (self.flags & (env.FLAG_CHANGED | env.FLAG_DONE)) == 0
                stalled += stuck & (self.games_done < quota)
                self._finish(torch.where(stuck, self.flags | env.FLAG_DONE, self.flags), quota)
        out = self._stats()
        out["stalled_games"] = int(stalled.sum().item())
        return out

    @torch.no_grad()
    def upleft_baseline(self, n_games: int) -> dict:
        """up, left, up, left, ...; when neither changed the board: down, right; when those did not
        either, the game ends (src/player.py:66-84)."""
        kw = dict(device=self.device)
        phase = torch.zeros(self.n, dtype=torch.int64, **kw)                 # 0 up, 1 left, 2 down, 3 right
        moved = torch.zeros(self.n, dtype=torch.bool, **kw)
        action_of = torch.tensor([0, 2, 1, 3], dtype=torch.uint8, **kw)      # phase -> action index
        done_bit = torch.tensor(env.FLAG_DONE, dtype=torch.uint8, **kw)
        quota = -(-n_games // self.n)
        while int(self.games_done.min().item()) < quota:
            for _ in range(32):
                self.t += 1
                env.step(self.boards, action_of[phase], seed=self.seed, step_index=self.t, p4=self.p4,
                         out=(self.next, self.reward, self.flags))
                changed = (self.flags & env.FLAG_CHANGED) != 0
                moved = torch.where((phase == 0) | (phase == 2), changed, moved | changed)
                over = (phase == 3) & ~moved
                restart = (phase == 1) & moved
                phase = torch.where(over | restart | ((phase == 3) & moved), 0, phase + 1)
                self._finish(torch.where(over, done_bit, torch.zeros_like(done_bit)).expand(self.n).contiguous(), quota)
        return self._stats()
