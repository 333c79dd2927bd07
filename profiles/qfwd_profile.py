"""A few launches of the fused conv Q-network forward on 1Mi boards (target of the ncu capture)."""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
import b2048
from b2048 import env
from torch import nn
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = nn.Sequential(nn.Conv2d(1, 64, 2), nn.ReLU(), nn.Conv2d(64, 64, 2), nn.ReLU(), nn.Flatten(), nn.Linear(256, 64), nn.ReLU(),
                    nn.Linear(64, 4)).double().to(dev)
fq = b2048.qfused.FusedConvQ(net)
b = env.random_boards(1 << 20, device=dev)
for _ in range(5):
    q = fq.forward_boards(b)
torch.cuda.synchronize()
print(float(q.abs().max()))
