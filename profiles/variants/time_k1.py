#!/usr/bin/env python3
"""Time the 64 Mi-board env step (K1) for the default build and for every variant .so given on the
command line (each in its own process through B2048_LIB).  Timing only: variants may be wrong."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, torch
sys.path.insert(0, "reinforcement-learning-2048_b200"); sys.path.insert(0, ".")
from b2048 import env
dev = torch.device("cuda:0")
n = 1 << 26
b = env.random_boards(n, device=dev); a = env.random_actions(n, device=dev)
out = (torch.empty_like(b), torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
for _ in range(5): env.step(b, a, out=out)
torch.cuda.synchronize()
res = []
for rep in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): env.step(b, a, step_index=i, out=out)
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 20)
res.sort()
chk = int(out[0][:1 << 20].sum().item()) & 0xFFFFFFFF
m = 1 << 22
w = torch.arange(1, m + 1, device=dev, dtype=torch.int64)
chk2 = int(((out[2][:m].to(torch.int64) * w).sum() + (out[1][:m].to(torch.int64) * w).sum()).item()) & 0xFFFFFFFF
print(f"min {res[0]:.4f} med {res[2]:.4f} ms  ({n * 22 / res[0] / 1e6:.0f} GB/s)  chk {chk:08x} flags+reward {chk2:08x}")
'''
def run(lib):
    envv = dict(os.environ)
    if lib: envv["B2048_LIB"] = os.path.abspath(lib)
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=ROOT, env=envv, capture_output=True, text=True)
    return (r.stdout.strip().splitlines() or ["FAILED: " + r.stderr.strip()[-300:]])[-1]
print(f"{'default':40s} {run(None)}", flush=True)
for so in sys.argv[1:]:
    print(f"{os.path.basename(so):40s} {run(so)}", flush=True)
print(f"{'default (again)':40s} {run(None)}", flush=True)
