"""A few launches of the training-size forwards of the conv update (5 000 states: the no-grad and the saving
instantiation of K6) — target of an ncu capture."""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
import b2048
from b2048 import qfused
from bench import conv_qnet
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = conv_qnet().to(dev)
params = [p.detach() for p in net.parameters()]
x = torch.randint(0, 12, (5000, 16), device=dev).double()
fq = qfused.FusedConvQ(net)
for _ in range(4):
    q0 = fq.forward_states(x) if hasattr(fq, "forward_states") else fq(x.view(-1, 1, 4, 4))
    q, saved = qfused.conv_q_forward_saving(x, params)
torch.cuda.synchronize()
print(float(q.abs().max()), float((q - q0).abs().max()))
