"""The conv update's backward kernels one at a time at batch 5000 (CUDA events, hot cache): fused conv2-dgrad +
conv1 backward, the 64 x 256 weight gradients, fc1's regrouped input gradient."""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048 import _lib
from b2048.env import _ptr, _stream
dev = torch.device("cuda:0")
_lib.init(0)
L = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
kw = dict(dtype=torch.float64, device=dev)
torch.manual_seed(0)
x = torch.randint(0, 12, (n, 16), device=dev).double()
w2 = torch.randn(64, 64, 2, 2, **kw)
p2 = torch.relu(torch.randn(4 * n, 256, **kw))
g2 = torch.randn(4 * n, 64, **kw)
g3, w3, a2 = torch.randn(n, 64, **kw), torch.randn(64, 256, **kw), torch.relu(torch.randn(n, 256, **kw))
dw1, db1 = torch.empty(64, 4, **kw), torch.empty(64, **kw)
dw2, db2 = torch.empty(64, 256, **kw), torch.empty(64, **kw)
s_f = torch.empty(L.conv2_dgrad_conv1_wgrad_scratch_elems(n), **kw)
s_w = torch.empty(L.layer_wgrad64_scratch_elems(4 * n, 256), **kw)
s_w3 = torch.empty(L.layer_wgrad64_scratch_elems(n, 256), **kw)
g2o = torch.empty(4 * n, 64, **kw)
s_k8 = torch.empty(max(L.dense_linear_wgrad_scratch_elems(4 * n, 256, 64), L.dense_linear_wgrad_scratch_elems(n, 256, 64)), **kw)
st = _stream(x)
cases = {
    "conv2_dgrad_conv1_wgrad": lambda: L.conv2_dgrad_conv1_wgrad_f64(_ptr(g2), _ptr(w2), _ptr(p2), _ptr(x), _ptr(dw1), _ptr(db1), _ptr(s_f), n, st),
    "layer_wgrad64 conv2 (20000 rows)": lambda: L.layer_wgrad64_f64(_ptr(g2), _ptr(p2), _ptr(dw2), _ptr(db2), _ptr(s_w), 4 * n, 256, st),
    "layer_wgrad64 fc1 (5000 rows)": lambda: L.layer_wgrad64_f64(_ptr(g3), _ptr(a2), _ptr(dw2), _ptr(db2), _ptr(s_w3), n, 256, st),
    "K8 dense_linear_wgrad conv2 (20000 rows)": lambda: L.dense_linear_wgrad_f64(_ptr(g2), _ptr(p2), _ptr(dw2), _ptr(db2), _ptr(s_k8), 4 * n, 256, 64, st),
    "K8 dense_linear_wgrad fc1 (5000 rows)": lambda: L.dense_linear_wgrad_f64(_ptr(g3), _ptr(a2), _ptr(dw2), _ptr(db2), _ptr(s_k8), n, 256, 64, st),
    "dgrad_regroup fc1": lambda: L.dense_linear_dgrad_regroup_f64(_ptr(g3), _ptr(w3), _ptr(a2), _ptr(g2o), n, 256, 64, 4, st),
}
for name, fn in cases.items():
    for _ in range(3):
        assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:36s} {e0.elapsed_time(e1) / reps * 1e3:8.2f} us per call (incl. its reduce kernel)", flush=True)
