import sys, torch, time
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048 import env
dev = torch.device('cuda:0')
for n in (1<<20, 1<<24, 1<<26):
    b = env.random_boards(n, device=dev); a = env.random_actions(n, device=dev)
    out = (torch.empty_like(b), torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
    for _ in range(3): env.step(b, a, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for i in range(K): env.step(b, a, step_index=i, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"step n={n}: {ms:.4f} ms  {n/ms*1e3:.3e} steps/s  {n*22/ms/1e6:.1f} GB/s algorithmic", flush=True)
n = 1<<20
b = env.random_boards(n, device=dev)
out4 = (torch.empty((n,4), dtype=torch.int64, device=dev), torch.empty((n,4), dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
for _ in range(3): env.step_all4(b, out=out4)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20): env.step_all4(b, step_index=i, out=out4)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/20
print(f"all4 n={n}: {ms:.4f} ms {4*n/ms*1e3:.3e} board-actions/s {n*57/ms/1e6:.1f} GB/s")
import os
for n in (1 << 20, 1 << 22):
    b = env.random_boards(n, device=dev)
    out4 = (torch.empty((n,4), dtype=torch.int64, device=dev), torch.empty((n,4), dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
    for mode in ("smem", "L2"):
        if mode == "L2": os.environ["B2048_ALL4_FROM_L2"] = "1"
        else: os.environ.pop("B2048_ALL4_FROM_L2", None)
        for _ in range(3): env.step_all4(b, out=out4)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50): env.step_all4(b, step_index=i, out=out4)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/50
        print(f"all4[{mode}] n={n}: {ms:.4f} ms {4*n/ms*1e3:.3e} board-actions/s {n*57/ms/1e6:.1f} GB/s", flush=True)
