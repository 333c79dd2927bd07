"""Eager (no CUDA graph) DDQN updates for an ncu launch list: which kernels make up one update."""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
import b2048
from b2048.rollout import VectorEnv
from b2048.trainer import DDQNUpdater
from bench import conv_qnet, dense_qnet
kind = sys.argv[1] if len(sys.argv) > 1 else "conv"
dev = torch.device("cuda:0")
ve = VectorEnv(1 << 16, device=dev, seed=3)
ring = b2048.ReplayRing(15000, device=dev)
ve.step(replay=ring)
torch.manual_seed(0)
up = DDQNUpdater((conv_qnet() if kind == "conv" else dense_qnet()).to(dev), ring, batch_size=5000, conv=kind == "conv", use_graph=False)
for _ in range(3):
    up.update()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("update")
up.update()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    up.update()
e1.record(); torch.cuda.synchronize()
print(kind, "eager ms/update", e0.elapsed_time(e1) / 20)
