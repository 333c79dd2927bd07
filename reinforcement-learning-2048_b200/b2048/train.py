"""Batched Double-DQN training loop: the reference's `training_loop` semantics on thousands of
concurrent games per GPU.

The reference (src/dqn_lib.py:167-233) plays one game at a time and, after every finished episode
`ep`: computes epsilon from the episode index, runs one `train_step` once `ep >
no_episodes_before_training`, and copies the online network into the target network every
`no_episodes_before_updating_target` episodes.  Here every step of `VectorEnv` advances n games at
once; the number of games that ended in that step is read back once per step and the same
per-episode rules are applied to the running episode counter:

    epsilon(ep)      = max((E - ep) / E, min_epsilon)                 (:184-185), ep = episodes finished so far
    updates owed     = one per finished episode beyond the warm-up     (:213)
    target sync      = whenever the episode counter crosses a multiple of K   (:227-228)

Updates are the real ones (zero_grad -> backward -> allreduce -> Adam, `DDQNUpdater`), not the
reference's no-op order.  With several ranks every rank steps its own shard of
games and replay ring; gradients are summed across ranks, so the schedule is driven by the LOCAL
episode count and `max_updates_per_step` keeps all ranks launching the same number of updates per
step (the update contains a collective).
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


def epsilon_for(ep: int, cfg: TrainConfig) -> float:
    return max((cfg.no_episodes_to_reach_epsilon - ep) / cfg.no_episodes_to_reach_epsilon, cfg.min_epsilon)


def train_batched(model: torch.nn.Module, cfg: TrainConfig, device="cuda", log_every: int = 0, on_log=None) -> dict:
    """Runs until `cfg.no_episodes` games have finished on this rank (with several ranks: on every
    rank).  Returns the rollout statistics plus the update / target-sync counts."""
    rank, world = bdist.world()
    dev = torch.device(device)
    base, _ = bdist.shard(cfg.n_envs * world, rank, world)
    venv = VectorEnv(cfg.n_envs, device=dev, seed=cfg.seed, index_base=base, p_four=cfg.p_four, conv=cfg.conv)
    ring = ReplayRing(cfg.replay_buffer_length, device=dev)
    updater = DDQNUpdater(model, ring, batch_size=cfg.batch_size, gamma=cfg.discount_factor, lr=cfg.learning_rate,
                          use_double=cfg.use_double_dqn, conv=cfg.conv, use_graph=cfg.use_graph, seed=cfg.seed + 1)
    episodes = updates = syncs = owed = steps = 0
    last_loss = None
    running = True
    while running:
        venv.step(model=updater.i_model, epsilon=epsilon_for(episodes, cfg), replay=ring)
        steps += 1
        finished = int(venv.totals[0].item())            # one scalar read-back per step
        # one update per finished episode once past the warm-up (src/dqn_lib.py:213)
        owed += max(0, finished - max(episodes, cfg.no_episodes_before_training + 1))
        # target sync whenever the counter crosses a multiple of K (src/dqn_lib.py:227)
        k = cfg.no_episodes_before_updating_target
        crossings = (finished - 1) // k - (episodes - 1) // k          # episode indices in [episodes, finished) divisible by K
        episodes = finished
        n_upd = min(owed, cfg.max_updates_per_step)
        running = episodes < cfg.no_episodes
        if world > 1:        # same number of collective launches on every rank, and a collective stop:
            t = torch.tensor([n_upd, -int(running)], device=dev)     # ranks that are through keep playing
            dist.all_reduce(t, op=dist.ReduceOp.MIN)                 # until the slowest one is (MIN of -running)
            n_upd, running = int(t[0].item()), bool(-int(t[1].item()))
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
    out = venv.stats()
    out.update(updates=updates, target_syncs=syncs, env_steps_per_game=out["mean_moves"], steps=steps,
               final_loss=None if last_loss is None else float(last_loss.item()))
    return out
