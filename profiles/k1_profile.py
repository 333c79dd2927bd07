"""ncu target: 4 launches of the env-step kernel at the bench size (64 Mi boards), nothing else.
    ncu --set full --clock-control none --import-source on -k regex:step_stream -s 1 -c 2 -o gpurun_out/prof python profiles/k1_profile.py [steady]
`steady`: boards after 64 random-policy steps from reset (SURVEY 8d "rollout-steady-state", seed 2049)."""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048 import env
dev = torch.device('cuda:0')
n = 1 << 26
if len(sys.argv) > 1 and sys.argv[1] == "steady":
    b = env.steady_state_boards(n, seed=2049, device=dev)
else:
    b = env.random_boards(n, seed=2048, device=dev)
a = env.random_actions(n, seed=2050, device=dev)
out = (torch.empty_like(b), torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.uint8, device=dev))
for i in range(4):
    env.step(b, a, seed=7, step_index=i, out=out)
torch.cuda.synchronize()
