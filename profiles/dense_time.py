"""Time the dense Double-DQN update (batch 5000) and its K8 layers; compare with the plain torch / cuBLAS path."""
import sys, torch, time
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
import b2048
from b2048 import _lib
from b2048.env import _ptr, _stream
from b2048.rollout import VectorEnv
from b2048.trainer import DDQNUpdater
from torch import nn
dev = torch.device('cuda:0')
def dense():
    return nn.Sequential(nn.Linear(16, 512), nn.ReLU(), nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, 4)).double()
def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
L = _lib.lib(); _lib.init(0)
rows = 5000
for n_in, n_out in ((512, 512), (512, 256), (16, 512), (256, 4)):
    x = torch.randn(rows, n_in, dtype=torch.float64, device=dev); w = torch.randn(n_out, n_in, dtype=torch.float64, device=dev)
    b = torch.randn(n_out, dtype=torch.float64, device=dev); g = torch.randn(rows, n_out, dtype=torch.float64, device=dev)
    out = torch.empty(rows, n_out, dtype=torch.float64, device=dev); dz = torch.empty(rows, n_in, dtype=torch.float64, device=dev)
    dw = torch.empty(n_out, n_in, dtype=torch.float64, device=dev); db = torch.empty(n_out, dtype=torch.float64, device=dev)
    sc = torch.empty(int(L.dense_linear_wgrad_scratch_elems(rows, n_in, n_out)), dtype=torch.float64, device=dev)
    st = _stream(x)
    gf = 2.0 * rows * n_in * n_out / 1e9
    t1 = timeit(lambda: L.dense_linear_forward_f64(_ptr(x), _ptr(w), _ptr(b), _ptr(out), rows, n_in, n_out, 0 if n_out == 4 else 1, st))
    t2 = timeit(lambda: L.dense_linear_dgrad_f64(_ptr(g), _ptr(w), _ptr(x), _ptr(dz), rows, n_in, n_out, st))
    t3 = timeit(lambda: L.dense_linear_wgrad_f64(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), _ptr(sc), rows, n_in, n_out, st))
    c1 = timeit(lambda: torch.relu(torch.addmm(b, x, w.t())))
    c2 = timeit(lambda: (g @ w) * (x > 0))
    c3 = timeit(lambda: (g.t() @ x, g.sum(0)))
    print(f"layer {n_in:3d}->{n_out:3d}: fwd {t1*1e3:6.1f} us ({gf/t1:5.1f} TF/s; torch {c1*1e3:6.1f})  dgrad {t2*1e3:6.1f} us ({gf/t2:5.1f}; torch {c2*1e3:6.1f})"
          f"  wgrad {t3*1e3:6.1f} us ({gf/t3:5.1f}; torch {c3*1e3:6.1f})", flush=True)
ve = VectorEnv(1 << 16, device=dev, seed=3)
ring = b2048.ReplayRing(15000, device=dev)
for _ in range(3): ve.step(replay=ring)
torch.manual_seed(0)
up = DDQNUpdater(dense().to(dev), ring, batch_size=5000, gamma=0.8, lr=1e-2, conv=False, use_graph=True)
print(type(up.f_model).__name__, type(up.i_model).__name__)
ms = timeit(lambda: up.update(), 100)
print(f"dense update (K8): {ms:.4f} ms = {1e3/ms:.0f} updates/s = {20.1216/ms:.1f} TFLOP/s", flush=True)
