"""Batched Double-DQN training loop: the reference's `training_loop` semantics on thousands of
concurrent games per GPU.

The reference (src/dqn_lib.py:167-233) plays one game at a time and, after every finished episode
`ep`: computes epsilon from the episode index, runs one `train_step` once `ep >
no_episodes_before_training`, and copies the online network into the target network every
`no_episodes_before_updating_target` episodes.  Here every step of `VectorEnv` advances n games at
once; the number of games that have ended is copied to the host asynchronously after every step and
the same per-episode rules are applied to that running episode counter (a few steps late, never
blocking the GPU):

    epsilon(ep)      = max((E - ep) / E, min_epsilon)                 (:184-185), ep = episodes finished so far
    updates owed     = one per finished episode beyond the warm-up     (:213)
    target sync      = whenever the episode counter crosses a multiple of K   (:227-228)

Updates are the real ones (zero_grad -> backward -> allreduce -> Adam, `DDQNUpdater`), not the
reference's no-op order.  With several ranks every rank steps its own shard of
games and replay ring; gradients are summed across ranks.  The schedule is driven by the MINIMUM over
the ranks of the local episode counters, obtained by one asynchronous NCCL all-reduce per step on the
device stream, so every rank launches the same number of updates per step (the update contains a
collective) without any host-side collective.
"""
from __future__ import annotations

import dataclasses

import torch
import torch.distributed as dist

from . import dist as bdist
from .replay import ReplayRing
from .rollout import VectorEnv
from .trainer import DDQNUpdater


@dataclasses.dataclass
class TrainConfig:
    """Field names follow the reference's config modules (src/configs/double_dqn_conv.py:33-47)."""
    n_envs: int = 4096
    replay_buffer_length: int = 15000
    batch_size: int = 5000
    discount_factor: float = 0.80
    learning_rate: float = 1e-2
    no_episodes: int = 30000
    no_episodes_to_reach_epsilon: int = 1000
    min_epsilon: float = 0.01
    no_episodes_before_training: int = 700
    no_episodes_before_updating_target: int = 100
    use_double_dqn: bool = True
    conv: bool = True
    p_four: float = 0.5
    seed: int = 0
    max_updates_per_step: int = 8
    use_graph: bool = True
    schedule_lag: int = 4          # the update schedule follows the episode counter of this many env steps ago


def epsilon_for(ep: int, cfg: TrainConfig) -> float:
    return max((cfg.no_episodes_to_reach_epsilon - ep) / cfg.no_episodes_to_reach_epsilon, cfg.min_epsilon)


class _LaggedCounter:
    """Host view of a device counter WITHOUT stalling the launch queue.

    After every environment step the finished-episode counter (with several ranks: its minimum over
    the ranks, one tiny NCCL all-reduce enqueued on the stream) is copied to pinned host memory
    asynchronously; the host reads the copy made `lag` steps earlier, whose event has normally long
    completed.  Every rank therefore sees the SAME sequence of values (they come out of the collective),
    so all ranks derive the same update schedule with no per-step host synchronisation and no blocking
    collective on the host path (round 1: one .item() plus one all_reduce(MIN).item() per step)."""

    def __init__(self, device, world: int, lag: int = 4, err_source=None):
        self.world, self.lag, self.err_source = world, max(1, lag), err_source
        self.slots = self.lag + 2
        self.dev = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(self.slots)]
        self.host = [torch.zeros(2, dtype=torch.int64).pin_memory() for _ in range(self.slots)]
        self.events = [torch.cuda.Event() for _ in range(self.slots)]
        self.dev0 = [d[:1] for d in self.dev]             # views made once: the loop is launch-bound on the host
        self.dev1 = [d[1:] for d in self.dev]
        self.host0 = [h[:1] for h in self.host]
        self.t = 0

    def push(self, counter: torch.Tensor) -> None:
        """`counter`: a one-element int64 device tensor (a persistent view, e.g. venv.totals[:1])."""
        i = self.t % self.slots
        if self.world == 1 and self.err_source is None:
            self.host0[i].copy_(counter, non_blocking=True)   # one 8-byte copy; the error word of the slot stays 0
        else:
            d = self.dev[i]
            self.dev0[i].copy_(counter)
            if self.err_source is not None:               # K5's sticky error flag rides along, negated for MIN
                self.dev1[i].copy_(-self.err_source)
            else:
                self.dev1[i].zero_()
            if self.world > 1:
                dist.all_reduce(d, op=dist.ReduceOp.MIN)  # asynchronous w.r.t. the host: enqueued on the stream
            self.host[i].copy_(d, non_blocking=True)
        self.events[i].record()
        self.t += 1

    def pop(self, final: bool = False):
        """(value, error) of the copy made `lag` pushes ago (the latest one with final=True), or None while
        the pipeline fills."""
        j = self.t - 1 if final else self.t - 1 - self.lag
        if j < 0:
            return None
        i = j % self.slots
        self.events[i].synchronize()
        v, e = self.host[i].tolist()
        return int(v), bool(e)


def train_batched(model: torch.nn.Module, cfg: TrainConfig, device="cuda", log_every: int = 0, on_log=None) -> dict:
    """Runs until `cfg.no_episodes` games have finished on this rank (with several ranks: on every
    rank).  Returns the rollout statistics plus the update / target-sync counts.

    The schedule is driven by the finished-episode counter as seen `cfg.schedule_lag` environment steps
    ago (see _LaggedCounter): the number of updates still equals the number of finished episodes beyond
    the warm-up, each is merely issued a few steps later, and the GPU never waits for the host."""
    rank, world = bdist.world()
    dev = torch.device(device)
    # every rank simulates cfg.n_envs games: its global index range is [rank * n_envs, (rank + 1) * n_envs),
    # never overlapping a neighbour's (overlapping ranges would share spawn / epsilon-greedy / reset streams)
    base = rank * cfg.n_envs
    venv = VectorEnv(cfg.n_envs, device=dev, seed=cfg.seed, index_base=base, p_four=cfg.p_four, conv=cfg.conv)
    ring = ReplayRing(cfg.replay_buffer_length, device=dev)
    updater = DDQNUpdater(model, ring, batch_size=cfg.batch_size, gamma=cfg.discount_factor, lr=cfg.learning_rate,
                          use_double=cfg.use_double_dqn, conv=cfg.conv, use_graph=cfg.use_graph, seed=cfg.seed + 1)
    err = updater.exchange.sync[2] if updater.exchange is not None else None
    counter = _LaggedCounter(dev, world, lag=cfg.schedule_lag, err_source=err)
    episodes = updates = syncs = owed = steps = 0
    last_loss = None
    k = cfg.no_episodes_before_updating_target

    def account(finished: int) -> int:
        """Apply the reference's per-episode rules to the episodes that finished since the last look;
        returns how many target syncs they imply."""
        nonlocal episodes, owed
        # one update per finished episode once past the warm-up (src/dqn_lib.py:213)
        owed += max(0, finished - max(episodes, cfg.no_episodes_before_training + 1))
        # target sync whenever the counter crosses a multiple of K (src/dqn_lib.py:227)
        crossings = (finished - 1) // k - (episodes - 1) // k      # episode indices in [episodes, finished) divisible by K
        episodes = finished
        return crossings

    finished_games = venv.totals[:1]                       # persistent view of the device counter
    while episodes < cfg.no_episodes:
        venv.step(model=updater.i_model, epsilon=epsilon_for(episodes, cfg), replay=ring)
        steps += 1
        counter.push(finished_games)
        seen = counter.pop()
        if seen is None:
            continue
        finished, failed = seen
        if failed:
            updater.check()                                # raises: a peer was lost in the gradient exchange
        crossings = account(finished)
        n_upd = min(owed, cfg.max_updates_per_step)
        for _ in range(n_upd):
            last_loss = updater.update()
        owed -= n_upd
        updates += n_upd
        if crossings > 0:
            updater.sync_target()
            syncs += 1
        if log_every and on_log is not None and steps % log_every == 0:
            on_log({"step": steps, "episodes": episodes, "updates": updates, "epsilon": epsilon_for(episodes, cfg),
                    "loss": None if last_loss is None else float(last_loss.item()), **venv.stats()})
    # drain: the episodes of the last `lag` steps are settled too, so that updates == episodes beyond the warm-up
    seen = counter.pop(final=True)
    if seen is not None:
        if seen[1]:
            updater.check()
        if account(seen[0]) > 0:
            updater.sync_target()
            syncs += 1
        for _ in range(owed):
            last_loss = updater.update()
        updates += owed
        owed = 0
    updater.check()
    out = venv.stats()
    out.update(updates=updates, target_syncs=syncs, env_steps_per_game=out["mean_moves"], steps=steps,
               final_loss=None if last_loss is None else float(last_loss.item()))
    return out
