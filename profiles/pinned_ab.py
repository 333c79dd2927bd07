"""Host<->device link rates for three kinds of pinned host memory: torch's pin_memory(), b2048_host_alloc bound
to the GPU's local CPUs, and b2048_host_alloc unbound (B2048_NO_NUMA_BIND=1).  Prints the topology the box exposes."""
import os, subprocess, sys, time
import torch
sys.path.insert(0, 'reinforcement-learning-2048_b200'); sys.path.insert(0, '.')
from b2048 import env, _lib
dev = torch.device("cuda:0")
_lib.init(0)
for cmd in ("nvidia-smi topo -m", "lscpu | head -25", "cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c",
            "nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader",
            "for d in /sys/bus/pci/devices/*; do if [ -e $d/local_cpulist ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/local_cpulist) node $(cat $d/numa_node); fi; done",
            "cat /sys/devices/system/node/online 2>/dev/null; taskset -p $$"):
    print("$", cmd); print(subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout[-1500:], flush=True)
n = 1 << 26
d_a = torch.empty(n, dtype=torch.int64, device=dev); d_b = torch.empty(n, dtype=torch.int64, device=dev)
side = torch.cuda.Stream(device=dev)

def rates(h_in, h_out, tag):
    res = {}
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            if mode in ("h2d", "both"): d_a.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(side): h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize(); res[mode] = 3 * n * 8 / (time.perf_counter() - t0) / 1e9
    print(f"{tag:28s} h2d {res['h2d']:6.1f}  d2h {res['d2h']:6.1f}  both (each way) {res['both']:6.1f} GB/s", flush=True)

a, b = torch.empty(n, dtype=torch.int64).pin_memory(), torch.empty(n, dtype=torch.int64).pin_memory()
rates(a, b, "torch pin_memory"); rates(a, b, "torch pin_memory (again)")
p1, p2 = env.PinnedBuffer(n, "int64", 0), env.PinnedBuffer(n, "int64", 0)
print("bound:", p1.bound, "numa node:", p1.numa_node)
rates(p1.tensor, p2.tensor, "b2048_host_alloc bound")
os.environ["B2048_NO_NUMA_BIND"] = "1"
q1, q2 = env.PinnedBuffer(n, "int64", 0), env.PinnedBuffer(n, "int64", 0)
rates(q1.tensor, q2.tensor, "b2048_host_alloc unbound")
rates(a, b, "torch pin_memory (last)")
