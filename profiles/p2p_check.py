"""torchrun --nproc-per-node 2 profiles/p2p_check.py : fused NVLink allreduce+Adam vs NCCL path + timing."""
import copy, os, sys, time, traceback
import torch, torch.distributed as dist
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import b2048
from b2048.rollout import VectorEnv
from b2048.trainer import DDQNUpdater
from bench import conv_qnet, dense_qnet
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
try:
    ve = VectorEnv(4096, device=dev, seed=5, index_base=rank * 4096)
    ring = b2048.ReplayRing(15000, device=dev)
    for _ in range(6):
        ve.step(replay=ring)
    for name, net, conv in (("conv", conv_qnet, True), ("dense", dense_qnet, False)):
        torch.manual_seed(0)
        base = net().to(dev)
        for mode in ("nccl", "p2p"):
            up = DDQNUpdater(copy.deepcopy(base), ring, batch_size=5000, lr=1e-3, conv=conv, use_graph=True, seed=11, exchange=mode)
            for _ in range(5):
                up.update()
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                up.update()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 200
            chk = up.params.flat.double().sum().item()
            flag = up.exchange.timed_out() if up.exchange is not None else None
            print(f"rank {rank} {name} {mode}: {ms:.4f} ms/update  param-sum {chk:.12e} timed_out={flag}", flush=True)
            dist.barrier()
except Exception:
    traceback.print_exc()
    os._exit(1)
dist.barrier()
dist.destroy_process_group()
