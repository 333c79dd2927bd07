"""End-to-end batched Double-DQN training (b2048.train.train_batched) with the reference's conv config:
episodes/s, env steps/s and updates/s over a fixed number of episodes."""
import sys, time, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048.train import TrainConfig, train_batched
from bench import conv_qnet
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = conv_qnet().to(dev)
cfg = TrainConfig(n_envs=int(sys.argv[1]) if len(sys.argv) > 1 else 4096, no_episodes=int(sys.argv[2]) if len(sys.argv) > 2 else 30000,
                  no_episodes_before_training=700, no_episodes_to_reach_epsilon=1000, batch_size=5000, learning_rate=1e-4,
                  max_updates_per_step=8)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = train_batched(model, cfg, device=dev)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"n_envs {cfg.n_envs}: {out['games']} episodes, {out['steps']} vector steps, {out['updates']} updates in {dt:.2f} s -> "
      f"{out['games'] / dt:.0f} episodes/s, {out['steps'] * cfg.n_envs / dt:.3e} env steps/s, {out['updates'] / dt:.0f} updates/s; "
      f"mean moves {out['mean_moves']:.1f}, mean merge score {out['mean_merge_score']:.0f}, final loss {out['final_loss']}")
