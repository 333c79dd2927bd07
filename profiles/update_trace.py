"""Per-kernel device times of graph-replayed DDQN updates (torch profiler / CUPTI), hot cache."""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
import b2048
from b2048.rollout import VectorEnv
from b2048.trainer import DDQNUpdater
from bench import conv_qnet, dense_qnet
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
ve = VectorEnv(1 << 16, device=dev, seed=3)
ring = b2048.ReplayRing(15000, device=dev)
ve.step(replay=ring)
kind = sys.argv[1] if len(sys.argv) > 1 else "conv"
torch.manual_seed(0)
up = DDQNUpdater((conv_qnet() if kind == "conv" else dense_qnet()).to(dev), ring, batch_size=5000, conv=kind == "conv")
for _ in range(5):
    up.update()
torch.cuda.synchronize()
N = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        up.update()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = {}
for e in evs:
    k = e.name[:90]
    t = tot.setdefault(k, [0, 0.0]); t[0] += 1; t[1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
span = (max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)) / N
busy = sum(v[1] for v in tot.values()) / N
print(f"{kind}: wall span per update {span:.1f} us, sum of kernel times {busy:.1f} us")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t / N:8.2f} us/update  x{n / N:4.1f}  {k}")

# one update, kernel by kernel in start order (last replay)
t0 = max(e.time_range.start for e in evs if "ring_sample" in e.name)
last = sorted((e for e in evs if e.time_range.start >= t0), key=lambda e: e.time_range.start)
print("--- last update, in start order: start(us) duration(us) name")
for e in last:
    d = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    print(f"{e.time_range.start - t0:8.1f} {d:8.2f}  {e.name[:100]}")
