"""DDQN updates/s at batch 5000 (CUDA graph), conv and dense — the bench.py extras, stand-alone."""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
import b2048
from b2048.rollout import VectorEnv
from b2048.trainer import DDQNUpdater
from bench import conv_qnet, dense_qnet
dev = torch.device("cuda:0")
ve = VectorEnv(1 << 16, device=dev, seed=3)
ring = b2048.ReplayRing(15000, device=dev)
ve.step(replay=ring)
for kind in sys.argv[1:] or ("conv", "dense"):
    torch.manual_seed(0)
    up = DDQNUpdater((conv_qnet() if kind == "conv" else dense_qnet()).to(dev), ring, batch_size=5000, conv=kind == "conv")
    for _ in range(5):
        up.update()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        up.update()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print(f"{kind}: {ms:.4f} ms/update  {1e3 / ms:.1f} updates/s  loss {float(up.loss):.6e}", flush=True)
