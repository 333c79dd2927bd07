"""The bench's e2e leg alone: b2048_step_host on 64Mi boards with pinned host buffers."""
import sys, time, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048 import env
dev = torch.device('cuda:0')
n = 1 << 26
boards = env.random_boards(n, device=dev); actions = env.random_actions(n, device=dev)
hb, ha = boards.cpu().pin_memory(), actions.cpu().pin_memory()
hn = torch.empty(n, dtype=torch.int64).pin_memory(); hr = torch.empty(n, dtype=torch.int32).pin_memory()
hf = torch.empty(n, dtype=torch.uint8).pin_memory()
env.step_host(hb, ha, hn, hr, hf, seed=1, step_index=0, device=0)
torch.cuda.synchronize(); t0 = time.perf_counter()
K = 5
for k in range(K):
    env.step_host(hb, ha, hn, hr, hf, seed=1, step_index=1 + k, device=0)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
print(f"e2e: {dt * 1e3:.2f} ms per 64Mi boards = {n / dt:.3e} steps/s  (H2D {9 * n / dt / 1e9:.1f} GB/s, D2H {13 * n / dt / 1e9:.1f} GB/s)")
