"""K8 forward GEMM (5000 rows x 512 outputs) against the reduction length: fixed cost per launch (prologue, epilogue,
launch gap) vs slope (the DMMA main loop).  B2048_LIB selects a diagnostic build (e.g. without the epilogue stores)."""
import sys, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048 import _lib
from b2048.env import _ptr, _stream
dev = torch.device('cuda:0')
L = _lib.lib(); _lib.init(0)
def timeit(fn, iters=100):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
rows, n_out = 5000, 512
res = []
for n_in in (64, 128, 256, 512, 1024, 2048):
    x = torch.randn(rows, n_in, dtype=torch.float64, device=dev); w = torch.randn(n_out, n_in, dtype=torch.float64, device=dev)
    b = torch.randn(n_out, dtype=torch.float64, device=dev); out = torch.empty(rows, n_out, dtype=torch.float64, device=dev)
    st = _stream(x)
    t = timeit(lambda: L.dense_linear_forward_f64(_ptr(x), _ptr(w), _ptr(b), _ptr(out), rows, n_in, n_out, 1, st))
    res.append((n_in, t))
    print(f"K={n_in:5d}: {t:7.2f} us  {2.0 * rows * n_in * n_out / t / 1e6:5.1f} TFLOP/s", flush=True)
(k0, t0), (k1, t1) = res[-3], res[-1]
slope = (t1 - t0) / (k1 - k0)
print(f"slope {slope * 512:.2f} us per 512 of K ({2.0 * rows * 512 * n_out / (slope * 512) / 1e6:.1f} TFLOP/s in the main loop), fixed {t0 - slope * k0:.2f} us per launch")
# the same for the input gradient (reduction over n_out, B operand reduction-strided, ReLU-mask epilogue) and the
# weight gradient (reduction over the rows, split over the SMs, + the fixed-order reduce pass)
print("dgrad  dz[5000 x 512] = (G[5000 x K] W[K x 512]) * (H > 0)")
res = []
for K in (128, 256, 512, 1024, 2048):
    g = torch.randn(rows, K, dtype=torch.float64, device=dev); w = torch.randn(K, 512, dtype=torch.float64, device=dev)
    h = torch.randn(rows, 512, dtype=torch.float64, device=dev); dz = torch.empty(rows, 512, dtype=torch.float64, device=dev)
    st = _stream(g)
    t = timeit(lambda: L.dense_linear_dgrad_f64(_ptr(g), _ptr(w), _ptr(h), _ptr(dz), rows, 512, K, st))
    res.append((K, t))
    print(f"K={K:5d}: {t:7.2f} us  {2.0 * rows * K * 512 / t / 1e6:5.1f} TFLOP/s", flush=True)
(k0, t0), (k1, t1) = res[-3], res[-1]
slope = (t1 - t0) / (k1 - k0)
print(f"slope {slope * 512:.2f} us per 512 of K, fixed {t0 - slope * k0:.2f} us per launch")
print("wgrad  dW[512 x 512] = G[R x 512]^T X[R x 512]  (+ reduce pass)")
res = []
for R in (2500, 5000, 10000, 20000):
    g = torch.randn(R, 512, dtype=torch.float64, device=dev); x = torch.randn(R, 512, dtype=torch.float64, device=dev)
    dw = torch.empty(512, 512, dtype=torch.float64, device=dev); db = torch.empty(512, dtype=torch.float64, device=dev)
    sc = torch.empty(int(L.dense_linear_wgrad_scratch_elems(R, 512, 512)), dtype=torch.float64, device=dev)
    st = _stream(g)
    t = timeit(lambda: L.dense_linear_wgrad_f64(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), _ptr(sc), R, 512, 512, st))
    res.append((R, t))
    print(f"R={R:5d}: {t:7.2f} us  {2.0 * R * 512 * 512 / t / 1e6:5.1f} TFLOP/s", flush=True)
(k0, t0), (k1, t1) = res[-3], res[-1]
slope = (t1 - t0) / (k1 - k0)
print(f"slope {slope * 5000:.2f} us per 5000 rows, fixed {t0 - slope * k0:.2f} us per call (GEMM + reduce)")
