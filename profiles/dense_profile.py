"""ncu target: the K8 layer kernels at the dense config's big shapes (batch 5000).
   ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -c 6 -o gpurun_out/prof_dense python profiles/dense_profile.py"""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048 import _lib
from b2048.env import _ptr, _stream
dev = torch.device('cuda:0'); L = _lib.lib(); _lib.init(0)
rows, n_in, n_out = 5000, 512, 512
x = torch.randn(rows, n_in, dtype=torch.float64, device=dev); w = torch.randn(n_out, n_in, dtype=torch.float64, device=dev)
b = torch.randn(n_out, dtype=torch.float64, device=dev); g = torch.randn(rows, n_out, dtype=torch.float64, device=dev)
out = torch.empty(rows, n_out, dtype=torch.float64, device=dev); dz = torch.empty(rows, n_in, dtype=torch.float64, device=dev)
dw = torch.empty(n_out, n_in, dtype=torch.float64, device=dev); db = torch.empty(n_out, dtype=torch.float64, device=dev)
sc = torch.empty(int(L.dense_linear_wgrad_scratch_elems(rows, n_in, n_out)), dtype=torch.float64, device=dev)
st = _stream(x)
for _ in range(2):
    L.dense_linear_forward_f64(_ptr(x), _ptr(w), _ptr(b), _ptr(out), rows, n_in, n_out, 1, st)
    L.dense_linear_dgrad_f64(_ptr(g), _ptr(w), _ptr(x), _ptr(dz), rows, n_in, n_out, st)
    L.dense_linear_wgrad_f64(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), _ptr(sc), rows, n_in, n_out, st)
torch.cuda.synchronize()
