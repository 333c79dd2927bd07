#!/usr/bin/env python3
"""bench.py — headline benchmark of the hot path: batched 2048 env steps/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA path)
    python bench.py --impl reference [--steps K] [--warmup W]       # CPU arm: the oracle port

A "step" is one pass of the env-step kernel over one batch: 64 Mi synthetic boards per GPU
(cell empty w.p. 0.3 else exponent uniform 1..11; uniform random actions incl.
illegal ones), one action each, Philox spawns at 10 % fours.  Inputs and outputs (1.4 GB) are far
larger than the 126 MB L2, so nothing is cache-resident between steps.  N > 1 runs one process per
GPU (torchrun); boards are sharded by contiguous global index, there is no data-path collective,
`value` is the whole-job aggregate and scaling is weak.

One JSON line is printed by rank 0 (see the bench contract in the task statement): `value` is
measured with inputs resident in HBM, `e2e` through the host-buffer C-ABI call (b2048_step_host)
with pinned host memory and both copies inside the timed region, `roofline` against the measured
HBM copy bandwidth in MEASURED_PEAKS.json (and, as `frac_vs_spec`, against the 8 TB/s of north_star),
`cpu_baseline` the C port of the reference algorithm (oracle/board_oracle.c) on the host cores of the
same box.  BASELINE.json's second metric, Double-DQN updates/sec at batch 5000, is the `ddqn` block
(conv + dense, fraction of the measured FP64 tensor-core peak, the unmodified reference's train_step on
this host beside it, replica identity at N > 1); `extra` carries config 2, config 5, the steady-state
distribution, rollouts, the end-to-end trainer, the link probe and the reference's own CPU figures.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "reinforcement-learning-2048_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

BOARDS_PER_GPU = 1 << 26           # 64 Mi boards = 512 MiB packed
BYTES_PER_STEP = 22                # board u64 in + action u8 in + board u64 out + reward i32 + flags u8
SEED_BOARDS, SEED_ACTIONS, SEED_SPAWN = 2048, 2050, 7
WORKLOAD = "env_step: 64Mi random boards/GPU x 1 action (roofline run of SURVEY 8d; streams >> L2)"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


HBM_SPEC_GBS = 8000.0              # north_star's denominator ("~8 TB/s"); reported beside the measured copy peak


def f64_peak_tflops():
    """FP64 tensor-core peak measured on this pool's B200 with profiles/ubench/dmma_peak.cu
    (mma.sync.m8n8k4.f64 = DMMA.8x8x4, the instruction K6 / K7 use; output committed beside the source)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ubench", "dmma_peak_b200.json")) as f:
            return float(json.load(f)["f64_dmma_peak_tflops"]), "measured: profiles/ubench/dmma_peak.cu -> dmma_peak_b200.txt"
    except Exception:
        return 37.2, "derived: 148 SMs x 128 flop/clk x 1.965 GHz"


def kernel_source_sha16():
    """Hash of the dominant kernel's machine code (the SASS of step_stream_kernel<false> in the built library, addresses
    and encodings included): ties the committed ncu traffic figure to a kernel version.  Falls back to a hash of the
    kernel's source files where cuobjdump is missing."""
    lib = os.path.join(ROOT, "reinforcement-learning-2048_b200", "b2048", "libb2048.so")
    try:
        import subprocess
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, timeout=120).stdout
        for block in sass.split("Function : ")[1:]:
            if "step_stream_kernelILb0" in block.split("\n", 1)[0]:
                return "sass:" + hashlib.sha256(block.split("\n", 1)[1].encode()).hexdigest()[:16]
    except Exception:
        pass
    h = hashlib.sha256()
    for name in ("env_kernels.cu", "b2048_common.cuh"):
        with open(os.path.join(ROOT, "reinforcement-learning-2048_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return "src:" + h.hexdigest()[:16]


def ncu_traffic_per_launch():
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture, and whether that capture
    was taken on the kernel sources of this tree (profiles/step_stream_traffic.json records their hash)."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_stream_traffic.json")) as f:
            d = json.load(f)
        return (float(d["dram_bytes_per_launch"]) * (BOARDS_PER_GPU / float(d["boards_per_launch"])),
                d.get("kernel_source_sha16") == kernel_source_sha16(), d.get("source"))
    except Exception:
        return None, False, None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples = []          # (t, sm_mhz, reasons_bitmask)
        self.stop_flag = False
        self.max_mhz = None
        self.ok = False
        self.ready = threading.Event()        # set after the first sample (the first NVML query is slow)
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            self.ready.set()
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            self.ready.set()
            time.sleep(self.period)

    def summary(self, t0: float, t1: float):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        win = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        bits = 0
        for s in win:
            bits |= s[2]
        reasons = [n for b, n in names.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(s[1] for s in win), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(win)}


def run_reference(args):
    """CPU arm: the reference algorithm (oracle C port of src/board.py) on all host cores."""
    import numpy as np
    from oracle import board_oracle as bo
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = bo.num_threads()
    sample = 1 << 21                      # boards per step (bounded sample of the 64Mi workload)
    boards = bo.random_boards(sample, seed=SEED_BOARDS)
    actions = np.random.default_rng(SEED_ACTIONS).integers(0, 4, size=sample, dtype=np.uint8)
    # size the sample so the whole run ends within a few minutes
    t = time.perf_counter()
    bo.step_packed(boards[: 1 << 17], actions[: 1 << 17], threads=threads)
    rate = (1 << 17) / (time.perf_counter() - t)
    budget_s = 60.0
    while sample > (1 << 16) and sample * (args.steps + args.warmup) / rate > budget_s:
        sample >>= 1
    boards, actions = boards[:sample], actions[:sample]
    for w in range(args.warmup):
        bo.step_packed(boards, actions, seed=SEED_SPAWN, step=w, threads=threads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        bo.step_packed(boards, actions, seed=SEED_SPAWN, step=args.warmup + k, threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} boards/step x {args.steps} steps of the same synthetic distribution, {threads} pthreads"
    print(json.dumps({
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": desc,
                   "note": "reference is single-process Python (~6e2 steps/s, BASELINE.md); this arm is its "
                           "algorithm restated in C (oracle/board_oracle.c) on all host cores"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def cpu_baseline_sample():
    import numpy as np
    from oracle import board_oracle as bo
    threads = bo.num_threads()
    n = 1 << 20
    boards = bo.random_boards(n, seed=SEED_BOARDS)
    actions = np.random.default_rng(SEED_ACTIONS).integers(0, 4, size=n, dtype=np.uint8)
    bo.step_packed(boards[: 1 << 16], actions[: 1 << 16], threads=threads)
    t0 = time.perf_counter()
    reps = 0
    while True:
        bo.step_packed(boards, actions, seed=SEED_SPAWN, step=reps, threads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > 10.0 or reps >= 64:
            break
    return {"value": n * reps / dt, "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": f"{reps} passes over {n} boards of the same synthetic distribution "
                      f"({dt:.1f} s wall, {threads} pthreads, oracle/board_oracle.c)"}


def reference_cpu_figures():
    """The reference's OWN code (oracle/_ref, unmodified) timed on this host: C1 Player.play_game(random) and
    C2 dqn_lib.train_step at batch 5000 (conv + dense); plus C1 again with this repo's drop-in
    `board` module on the GPU (BASELINE config 1 through the per-call shim).  Each runs in its own process."""
    import subprocess
    script = os.path.join(ROOT, "oracle", "ref_bench.py")
    cwd = os.path.join(ROOT, "gpurun_out")
    os.makedirs(os.path.join(cwd, ".git"), exist_ok=True)        # the reference's Experiment wants a git root (unused here)
    out = {}

    def run(tag, extra_args, timeout):
        try:
            r = subprocess.run([sys.executable, script] + extra_args, cwd=cwd, capture_output=True, text=True,
                               timeout=timeout, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            out[tag] = json.loads(line[-1]) if line else {"unavailable": (r.stderr or r.stdout)[-300:]}
        except Exception as e:          # noqa: BLE001 - a baseline that cannot run is reported, not fatal
            out[tag] = {"unavailable": repr(e)[:300]}

    run("reference_on_host_cpu", ["--device", "cpu", "--c1-seconds", "5", "--c2-calls", "2"], 240)
    run("b2048_shim_player_loop", ["--engine", "b2048", "--device", "cuda", "--c1-seconds", "5", "--skip-c2"], 240)
    out["kind"] = "reference"
    out["cores"] = os.cpu_count()
    return out


def conv_qnet():
    from torch import nn   # reference configs/double_dqn_conv.py:19-28 (33 476 parameters, float64)
    return nn.Sequential(nn.Conv2d(1, 64, kernel_size=2), nn.ReLU(), nn.Conv2d(64, 64, kernel_size=2), nn.ReLU(),
                         nn.Flatten(), nn.Linear(2 * 2 * 64, 64), nn.ReLU(), nn.Linear(64, 4)).double()


def dense_qnet():
    from torch import nn   # reference configs/double_dqn_dense.py:7-15 (403 716 parameters, float64)
    return nn.Sequential(nn.Linear(16, 512), nn.ReLU(), nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, 256),
                         nn.ReLU(), nn.Linear(256, 4)).double()


def secondary_metrics(args, dev, rank, world, barrier):
    """The other numbers BASELINE.json asks for, reported beside the headline (not the roofline
    kernel): config 2 (1Mi boards x 4 actions, L2-resident), DDQN updates/sec at batch 5000 for the
    conv and dense Q-networks (real updates: sample -> 3 forwards -> fused target/loss -> backward
    -> NCCL allreduce -> Adam, one CUDA graph), and a random-policy rollout with replay append."""
    import torch
    import torch.distributed as dist
    import b2048
    from b2048 import env
    from b2048.rollout import VectorEnv
    from b2048.trainer import DDQNUpdater

    def timed(fn, iters):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / iters

    out = {}
    n2 = 1 << 20
    b2 = env.random_boards(n2, seed=SEED_BOARDS, index_base=rank * n2, device=dev)
    o2 = (torch.empty((n2, 4), dtype=torch.int64, device=dev), torch.empty((n2, 4), dtype=torch.int32, device=dev),
          torch.empty(n2, dtype=torch.uint8, device=dev))
    def all4_ms():
        for w in range(3):
            env.step_all4(b2, seed=SEED_SPAWN, step_index=w, index_base=rank * n2, out=o2)
        return timed(lambda i: env.step_all4(b2, seed=SEED_SPAWN, step_index=3 + i, index_base=rank * n2, out=o2), 50)
    ms = all4_ms()                                       # persistent kernel, row table staged in shared memory
    os.environ["B2048_ALL4_FROM_L2"] = "1"               # A/B: one board per thread, row table gathered from L2
    ms_l2 = all4_ms()
    os.environ.pop("B2048_ALL4_FROM_L2", None)
    out["config2_all4"] = {"workload": "1Mi random boards x 4 actions per GPU (57 MB/launch: L2-resident, not a roofline run)",
                           "board_actions_per_sec": world * 4 * n2 / (ms * 1e-3), "ms_per_launch": ms,
                           "algorithmic_GBps": n2 * 57 / (ms * 1e-3) / 1e9,
                           "kernel": "step_all4_stream_kernel (row table in shared memory, 8 boards per thread and Philox call)",
                           "ab_table_from_l2": {"kernel": "step_all4_kernel (16 L2 gathers per board)", "ms_per_launch": ms_l2,
                                                "board_actions_per_sec": world * 4 * n2 / (ms_l2 * 1e-3)}}

    # rollout: random policy incl. legal mask, replay append and masked reset (5 launches + torch bookkeeping / step)
    nv = 1 << 22
    ve = VectorEnv(nv, device=dev, seed=3, index_base=rank * nv, p_four=0.1)
    ring = b2048.ReplayRing(15000, device=dev)
    for _ in range(3):
        ve.step(replay=ring)
    ms = timed(lambda i: ve.step(replay=ring), 20)
    out["rollout_random_policy"] = {"workload": "4Mi concurrent games per GPU, random policy, replay ring 15000, auto-reset",
                                    "env_steps_per_sec": world * nv / (ms * 1e-3), "ms_per_step": ms}

    # K6: the conv Q-network's no-gradient forward as one fused FP64 tensor-core kernel, alone and inside
    # an epsilon-greedy rollout step (legal mask -> Q -> action -> env step -> replay append -> reset)
    torch.manual_seed(0)
    qnet = conv_qnet().to(dev)
    fq = b2048.qfused.FusedConvQ(qnet)
    nq = 1 << 20
    bq = env.random_boards(nq, seed=SEED_BOARDS, index_base=rank * nq, device=dev)
    qo = torch.empty((nq, 4), dtype=torch.float64, device=dev)
    for _ in range(3):
        fq.forward_boards(bq, out=qo)
    ms = timed(lambda i: fq.forward_boards(bq, out=qo), 10)
    out["qnet_forward_conv_fused"] = {
        "workload": "conv Q-net float64 forward of 1Mi packed boards per GPU, one kernel (DMMA), 168 960 flop/board",
        "boards_per_sec": world * nq / (ms * 1e-3), "ms_per_launch": ms, "fp64_TFLOPs_per_gpu": nq * 168960 / (ms * 1e-3) / 1e12}
    ng = 1 << 20
    vg = VectorEnv(ng, device=dev, seed=5, index_base=rank * ng, p_four=0.1)
    for _ in range(3):
        vg.step(model=fq, epsilon=0.1, replay=ring)
    ms = timed(lambda i: vg.step(model=fq, epsilon=0.1, replay=ring), 10)
    out["rollout_egreedy_conv"] = {"workload": "1Mi concurrent games per GPU, epsilon-greedy (0.1) on the conv Q-net via the fused "
                                               "forward, replay ring 15000, auto-reset",
                                   "env_steps_per_sec": world * ng / (ms * 1e-3), "ms_per_step": ms}
    del vg, bq, qo

    # K1 on the second distribution of SURVEY 8(d): boards after 64 random legal moves from reset (seed 2049).
    # The shared-memory bank-conflict rate depends on the distribution; ncu counters for both are in profiles/.
    ns = BOARDS_PER_GPU
    bs = env.steady_state_boards(ns, seed=2049, index_base=rank * ns, device=dev)
    as_ = env.random_actions(ns, seed=SEED_ACTIONS, index_base=rank * ns, device=dev)
    os_ = (torch.empty_like(bs), torch.empty(ns, dtype=torch.int32, device=dev), torch.empty(ns, dtype=torch.uint8, device=dev))
    for w in range(3):
        env.step(bs, as_, seed=SEED_SPAWN, step_index=w, index_base=rank * ns, out=os_)
    ms = timed(lambda i: env.step(bs, as_, seed=SEED_SPAWN, step_index=3 + i, index_base=rank * ns, out=os_), 20)
    out["env_step_steady_state"] = {
        "workload": "64Mi boards/GPU after 64 random legal moves from reset (SURVEY 8d 'rollout-steady-state', seed 2049) x 1 action",
        "env_steps_per_sec": world * ns / (ms * 1e-3), "ms_per_launch": ms,
        "algorithmic_GBps": ns * BYTES_PER_STEP / (ms * 1e-3) / 1e9}
    del bs, as_, os_

    # ---- DDQN updates/sec at batch 5000 (BASELINE.json's second metric): see the top-level "ddqn" block ----------
    peak_tf, peak_src = f64_peak_tflops()
    ddqn_block = {"metric": "ddqn_updates_per_sec", "batch_per_gpu": 5000, "global_batch": 5000 * world, "dtype": "f64",
                  "f64_peak_tflops": peak_tf, "f64_peak_source": peak_src,
                  "update": "sample(K2) + 3 forwards + fused target/loss(K3) + backward + gradient exchange + Adam, one CUDA graph"}
    checks = {}
    for name, net, conv, gflop in (("conv", conv_qnet, True, 4.2240), ("dense", dense_qnet, False, 20.1216)):
        torch.manual_seed(0)
        up = DDQNUpdater(net().to(dev), ring, batch_size=5000, gamma=0.8, lr=1e-2, conv=conv, use_graph=True)
        for _ in range(3):
            up.update()
        ms = timed(lambda i: up.update(), 100)
        up.check()
        tf = gflop / ms                                           # GFLOP / ms = TFLOP/s (per GPU)
        ddqn_block[name] = {"updates_per_sec": 1e3 / ms, "ms_per_update": ms, "algorithmic_gflop_per_update": gflop,
                            "fp64_tflops_per_gpu": tf, "frac_of_f64_peak": tf / peak_tf}
        out[f"ddqn_updates_{name}"] = {
            "workload": f"{name} Q-net float64, batch 5000/GPU, replay 15000, gamma 0.8 (f32), Double DQN, Adam; "
                        "sample+3 fwd+fused loss+bwd+allreduce+Adam in one CUDA graph",
            "updates_per_sec": 1e3 / ms, "ms_per_update": ms, "global_batch": 5000 * world}
        if world > 1:                # replicas must be bit-identical after the fused NVLink allreduce + Adam (K5)
            flat = up.params.flat.detach()
            bits = flat.view(torch.int64)
            mine = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=dev)).sum(),
                                flat.sum().view(torch.int64)])
            everyone = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(everyone, mine)
            same = all(torch.equal(everyone[0], e) for e in everyone)
            checks[name] = {"replicas_bit_identical": bool(same), "params_sum": float(flat.sum().item()),
                            "hash": [int(v) for v in mine.tolist()][:2]}
            if not same:
                raise RuntimeError(f"{name}: parameter replicas differ across ranks after the gradient exchange")
    if checks:
        ddqn_block["replica_check"] = checks
    out["__ddqn__"] = ddqn_block

    # ---- BASELINE config 5 as stated: 8 Mi concurrent games per GPU (64 Mi on 8), per-GPU replay shard, conv Double DQN
    # with the gradient exchange -- one combined step = one env step of every game (epsilon-greedy on the conv Q-net,
    # replay append, auto-reset) + one update at batch 5000/GPU
    n5 = 1 << 23
    torch.manual_seed(0)
    net5 = conv_qnet().to(dev)
    ring5 = b2048.ReplayRing(15000, device=dev)
    up5 = DDQNUpdater(net5, ring5, batch_size=5000, gamma=0.95, lr=1e-4, conv=True, use_graph=True)
    v5 = VectorEnv(n5, device=dev, seed=17, index_base=rank * n5, p_four=0.1)
    for _ in range(2):
        v5.step(model=up5.i_model, epsilon=0.1, replay=ring5)
        up5.update()
    ms_env = timed(lambda i: v5.step(model=up5.i_model, epsilon=0.1, replay=ring5), 5)
    ms_rand = timed(lambda i: v5.step(replay=ring5), 10)
    ms_upd = timed(lambda i: up5.update(), 50)

    def combined(i):
        v5.step(model=up5.i_model, epsilon=0.1, replay=ring5)
        up5.update()
    ms_both = timed(combined, 5)
    up5.check()
    out["config5"] = {
        "workload": f"{n5} concurrent games per GPU ({world * n5} total), epsilon-greedy (0.1) on the conv Q-net (K6) + K1 + replay "
                    "shard (K2 append) + auto-reset, and one conv Double-DQN update per step at batch 5000/GPU, gamma 0.95, "
                    "gradient exchange over NVLink (K5) at N > 1",
        "combined_step_ms": ms_both, "env_steps_per_sec": world * n5 / (ms_both * 1e-3), "updates_per_sec": 1e3 / ms_both,
        "env_step_alone_ms": ms_env, "env_steps_per_sec_alone": world * n5 / (ms_env * 1e-3),
        "random_policy_step_alone_ms": ms_rand, "random_policy_env_steps_per_sec": world * n5 / (ms_rand * 1e-3),
        "update_alone_ms": ms_upd, "updates_per_sec_alone": 1e3 / ms_upd, "global_batch": 5000 * world}
    del v5, up5, ring5
    # the whole loop: batched Double-DQN training with the reference's conv config and schedule
    # (train_batched: epsilon-greedy rollouts on K6 -> replay ring -> one update per finished episode)
    import time
    from b2048.train import TrainConfig, train_batched
    torch.manual_seed(0)
    cfg = TrainConfig(n_envs=4096, no_episodes=8000, no_episodes_before_training=700, no_episodes_to_reach_epsilon=1000,
                      batch_size=5000, learning_rate=1e-4, max_updates_per_step=8, seed=11)
    barrier()
    t0 = time.perf_counter()
    st = train_batched(conv_qnet().to(dev), cfg, device=dev)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    out["train_batched_conv"] = {
        "workload": "reference conv config end to end: 4096 concurrent games per GPU, 8000 episodes per GPU, one update "
                    "(batch 5000/GPU) per finished episode after 700, target sync every 100; wall clock incl. graph capture",
        "episodes_per_sec": world * st["games"] / dt, "updates_per_sec": st["updates"] / dt,
        "env_steps_per_sec": world * st["steps"] * cfg.n_envs / dt, "seconds": dt}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from b2048 import env

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this framework has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug lines go to stderr
        # (NCCL ignores NCCL_DEBUG_FILE at level VERSION, so that level is raised to WARN)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n = BOARDS_PER_GPU
    base = rank * n                                     # contiguous global index shard
    boards = env.random_boards(n, seed=SEED_BOARDS, index_base=base, device=dev)
    actions = env.random_actions(n, seed=SEED_ACTIONS, index_base=base, device=dev)
    out = (torch.empty_like(boards), torch.empty(n, dtype=torch.int32, device=dev),
           torch.empty(n, dtype=torch.uint8, device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local, period_s=0.002)
    sampler.start()
    sampler.ready.wait(timeout=5.0)
    for w in range(args.warmup):
        env.step(boards, actions, seed=SEED_SPAWN, step_index=w, index_base=base, out=out)
    barrier()
    # Two events around the K back-to-back launches (none in between: the kernel uses programmatic dependent
    # launch, and an event record between two launches would serialise them); one launch per step, so the
    # kernel's average launch duration over the timed region is total / K.
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        env.step(boards, actions, seed=SEED_SPAWN, step_index=args.warmup + k, index_base=base, out=out)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    total_ms = ev0.elapsed_time(ev1)
    kernel_ms = [total_ms / args.steps]
    clocks = None
    if rank == 0:
        clocks = sampler.summary(t_wall0, t_wall1)

    if args.headline_only:
        sampler.stop_flag = True
        if rank == 0:
            print(json.dumps({"metric": "env_steps_per_sec", "value": world * n * args.steps / (total_ms * 1e-3),
                              "ms_per_step": total_ms / args.steps, "headline_only": True}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region ----------------------
    # Pinned buffers come from b2048_host_alloc: placed on the NUMA node next to this rank's GPU, and the calling
    # thread is bound there too, so that N ranks do not stage through the far socket.
    env.bind_thread_near(local)
    pb = {k: env.PinnedBuffer(n, dt, device=local) for k, dt in
          (("boards", "int64"), ("actions", "uint8"), ("next", "int64"), ("reward", "int32"), ("flags", "uint8"))}
    hb, ha, hn, hr, hf = (pb[k].tensor for k in ("boards", "actions", "next", "reward", "flags"))
    hb.copy_(boards)
    ha.copy_(actions)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 5))
    env.step_host(hb, ha, hn, hr, hf, seed=SEED_SPAWN, step_index=0, index_base=base, device=local)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        env.step_host(hb, ha, hn, hr, hf, seed=SEED_SPAWN, step_index=1 + k, index_base=base, device=local)
    barrier()
    e2e_s = time.perf_counter() - t0
    # what the link itself gives, all ranks at once: plain copies of the same buffers, one direction at a time
    # and both together (the ceiling of the e2e figure: 9 B/step up, 13 B/step down)
    link = {}
    scratch = torch.empty(n, dtype=torch.int64, device=dev)
    side = torch.cuda.Stream(device=dev)
    for tag in ("h2d", "d2h", "both"):
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            if tag in ("h2d", "both"):
                boards.copy_(hb, non_blocking=True)
            if tag in ("d2h", "both"):
                with torch.cuda.stream(side):
                    hn.copy_(scratch, non_blocking=True)
        barrier()
        link[tag] = time.perf_counter() - t0
    del scratch
    link_t = torch.tensor([link["h2d"], link["d2h"], link["both"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(link_t, op=dist.ReduceOp.MAX)
    gb = 3 * n * 8 / 1e9
    e2e_link = {"h2d_GBps_per_gpu": gb / float(link_t[0]), "d2h_GBps_per_gpu": gb / float(link_t[1]),
                "bidirectional_GBps_per_gpu_each_way": gb / float(link_t[2]), "concurrent_ranks": world,
                "pinned_numa_node": pb["boards"].numa_node, "numa_bound": pb["boards"].bound,
                "what": "cudaMemcpyAsync of 512 MiB pinned <-> device on every rank at once (max time over ranks)"}
    e2e_link["e2e_ceiling_steps_per_sec"] = world * min(e2e_link["bidirectional_GBps_per_gpu_each_way"] * 1e9 / 13.0,
                                                        e2e_link["h2d_GBps_per_gpu"] * 1e9 / 9.0)
    sampler.stop_flag = True

    extra = secondary_metrics(args, dev, rank, world, barrier)
    ddqn_block = extra.pop("__ddqn__")
    extra["e2e_link"] = e2e_link

    times = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = times.tolist()
    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = world * n * args.steps / (total_ms * 1e-3)
        peak, peak_src = measured_peak_gbs()
        k_ms = statistics.mean(kernel_ms)
        achieved = n * BYTES_PER_STEP / (k_ms * 1e-3) / 1e9
        traffic, traffic_ok, traffic_src = ncu_traffic_per_launch()
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "boards_per_gpu": n, "p_four": 0.1,
                       "l2": "inputs+outputs 1.4 GB/step >> 126 MB L2 (no flush needed)",
                       "parallelism": f"env-shard x{world} (no data-path collective)",
                       "timing": "two CUDA events around the K back-to-back launches on torch's current stream (the launch stream); max over ranks"},
            "clocks": clocks,
            "e2e": {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": "steps/s",
                    "h2d_bytes_per_step": n * 9, "d2h_bytes_per_step": n * 13, "steps": e2e_steps,
                    "api": "b2048_step_host (NUMA-local pinned host buffers from b2048_host_alloc, 3-slot H2D|kernel|D2H pipeline)",
                    "link_ceiling_steps_per_sec": e2e_link["e2e_ceiling_steps_per_sec"]},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_from_this_kernel_version": traffic_ok,
                         "traffic_source": traffic_src,
                         "kernel": "step_stream_kernel<false>", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": n * BYTES_PER_STEP, "peak_source": peak_src,
                         "peak_spec": HBM_SPEC_GBS, "frac_vs_spec": achieved / HBM_SPEC_GBS,
                         "steady_state_distribution": {"achieved": extra["env_step_steady_state"]["algorithmic_GBps"],
                                                       "frac": extra["env_step_steady_state"]["algorithmic_GBps"] / peak}},
            "ddqn": ddqn_block,
        }
        line["extra"] = extra
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_sample()
            ref = reference_cpu_figures()
            extra["reference_cpu"] = ref
            host = ref.get("reference_on_host_cpu", {})
            for name in ("conv", "dense"):
                c2 = host.get(f"c2_train_step_{name}", {})
                if "updates_per_sec" in c2:
                    ddqn_block[name]["reference_cpu_updates_per_sec"] = c2["updates_per_sec"]
                    ddqn_block[name]["reference_cpu"] = (f"dqn_lib.train_step of the unmodified reference on this host, "
                                                         f"{host.get('torch_threads')} torch threads of {host.get('cpu_count')} CPUs")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--headline-only", action="store_true",
                    help="only the device-resident env-step timing (for ncu launch lists); no e2e / extras / CPU baseline")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                       # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
