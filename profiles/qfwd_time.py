"""Time the conv Q-network forward (cuBLAS path vs fused kernel when present) and model-driven rollout steps."""
import sys, time, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import b2048
from b2048 import env
from b2048.qnet import accelerate
from b2048.rollout import VectorEnv
from torch import nn
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = nn.Sequential(nn.Conv2d(1, 64, 2), nn.ReLU(), nn.Conv2d(64, 64, 2), nn.ReLU(), nn.Flatten(), nn.Linear(256, 64), nn.ReLU(),
                    nn.Linear(64, 4)).double().to(dev)
fast = accelerate(net)
def timeit(f, k=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
for n in (4096, 5000, 65536, 1 << 20):
    b = env.random_boards(n, device=dev)
    x = env.unpack_f64(b, conv=True)
    with torch.no_grad():
        ms = timeit(lambda: fast(x))
        line = f"n={n}: cuBLAS-path forward {ms:.3f} ms ({n / ms * 1e3:.3e} boards/s)"
        if hasattr(b2048, "qfused"):
            fq = b2048.qfused.FusedConvQ(net)
            ms2 = timeit(lambda: fq.forward_boards(b))
            err = (fq.forward_boards(b) - fast(x)).abs().max().item()
            line += f" | fused {ms2:.3f} ms ({n / ms2 * 1e3:.3e} boards/s), max |dQ| {err:.2e}"
    print(line, flush=True)
for n in (4096, 65536, 1 << 20):
    for name, m in (("GEMM path", fast), ("fused K6", b2048.qfused.FusedConvQ(net))):
        ve = VectorEnv(n, device=dev, seed=1)
        ms = timeit(lambda: ve.step(model=m, epsilon=0.1), k=10)
        print(f"epsilon-greedy rollout step ({name}), n={n}: {ms:.3f} ms/step ({n / ms * 1e3:.3e} steps/s)", flush=True)
