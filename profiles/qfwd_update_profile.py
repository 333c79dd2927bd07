"""The conv update's merged forward (Q(s) saving + Q_online(s') + Q_target(s'), 15 000 boards, one K6 launch) a few
times -- target of an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:qconv_forward -s 2 -c 1 -o gpurun_out/qfwd_update python profiles/qfwd_update_profile.py"""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048 import qfused
from bench import conv_qnet
dev = torch.device('cuda:0')
torch.manual_seed(0)
net, tgt = conv_qnet().to(dev), conv_qnet().to(dev)
a, b = qfused.TrainableConvQ(net), qfused.TrainableConvQ(tgt)
sn = torch.randint(0, 12, (10000, 16), device=dev).double()
x, xn = sn[:5000], sn[5000:]
for _ in range(4):
    out = a.forward_update(x, xn, b, True)
torch.cuda.synchronize()
print(float(out[0].abs().max()))
