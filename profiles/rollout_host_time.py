"""Host vs device time of one VectorEnv.step (wall clock of the Python call sequence against the CUDA-event span of the
same steps): is the rollout / training loop launch-bound on the host?"""
import sys, time, torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
import b2048
from b2048.rollout import VectorEnv
from b2048.qfused import accelerate_inference
from bench import conv_qnet
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = accelerate_inference(conv_qnet().to(dev))
for n in (4096, 1 << 16, 1 << 22):
    for greedy in (False, True):
        ve = VectorEnv(n, device=dev, seed=1)
        ring = b2048.ReplayRing(15000, device=dev)
        kw = dict(model=net, epsilon=0.1) if greedy else {}
        for _ in range(20): ve.step(replay=ring, **kw)
        torch.cuda.synchronize()
        iters = 200 if n <= (1 << 16) else 30
        # host time: enqueue only (the queue never fills at these sizes if the device is faster)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(iters): ve.step(replay=ring, **kw)
        t_host = (time.perf_counter() - t0) / iters
        e1.record(); torch.cuda.synchronize()
        t_wall = (time.perf_counter() - t0) / iters
        t_dev = e0.elapsed_time(e1) / iters * 1e-3
        print(f"n={n:8d} {'egreedy' if greedy else 'random ':7s}: host enqueue {t_host*1e6:7.1f} us/step, device span {t_dev*1e6:7.1f} us/step, "
              f"wall {t_wall*1e6:7.1f} us/step -> {n / t_wall:.3e} env steps/s", flush=True)
