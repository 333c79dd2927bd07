#!/usr/bin/env python3
"""Time the reference's OWN code on this machine (baseline infrastructure; executed only by bench.py's
reference / cpu_baseline legs and by hand -- nothing in the product imports it).

Runs the unmodified reference from oracle/_ref/src (oracle/make_ref.py), single process, exactly the two
measurements SURVEY.md §8(d) names:
  C1  Player.play_game(random_policy=True)      src/player.py:40-64   -> env steps / s   (BASELINE config 1)
  C2  dqn_lib.train_step at batch 5000          src/dqn_lib.py:119-164 -> updates / s     (configs 3 and 4)
      conv: configs/double_dqn_conv.py (replay 15000); dense: configs/double_dqn_dense.py (replay 100000)
Prints ONE JSON object.  `--device cpu` hides the GPU from the reference (its device/__init__.py then picks
"cpu"); `--device cuda` lets it take its stock `cuda:0` path (model on the GPU, everything else as is).
With `--engine b2048` the same C1 loop runs against THIS repo's drop-in `board` / `dqn_lib` modules
(the per-call shim, BASELINE config 1 on the CUDA engine) -- same script, different sys.path.
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref", "src")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--engine", default="reference", choices=["reference", "b2048"])
    ap.add_argument("--c1-seconds", type=float, default=6.0)
    ap.add_argument("--c2-calls", type=int, default=3)
    ap.add_argument("--skip-c2", action="store_true")
    args = ap.parse_args()
    if not os.path.isfile(os.path.join(REF, "board.py")):
        print(json.dumps({"unavailable": "oracle/_ref is absent (run oracle/make_ref.py where /root/reference exists)"}))
        return
    if args.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""          # before torch is imported: the reference then picks "cpu"
    if args.engine == "b2048":
        sys.path[:0] = [os.path.join(os.path.dirname(HERE), "reinforcement-learning-2048_b200"), REF]
    else:
        sys.path.insert(0, REF)
    import logging
    logging.disable(logging.WARNING)
    import numpy as np
    import torch
    import board as board_mod
    import dqn_lib
    import player as player_mod
    from device import device

    out = {"engine": args.engine, "device": device, "torch_threads": torch.get_num_threads(), "cpu_count": os.cpu_count(),
           "numpy": np.__version__, "torch": torch.__version__, "board_module": os.path.abspath(board_mod.__file__)}

    # ---- C1: the reference's Player loop, random policy (Experiment bypassed: it only stores results) --------------
    pl = object.__new__(player_mod.Player)
    pl.device, pl.model, pl.games_history, pl.reward_func = device, None, [], dqn_lib.reward_func_merge_score
    pl.play_game(random_policy=True)                      # warm-up (imports, CUDA context for the shim)
    steps = games = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < args.c1_seconds:
        steps += len(pl.play_game(random_policy=True))
        games += 1
    dt = time.perf_counter() - t0
    out["c1_player_random"] = {"steps_per_sec": steps / dt, "games": games, "steps": steps, "seconds": dt,
                               "what": "Player.play_game(random_policy=True), src/player.py:40-64, one process"}

    # ---- C2: dqn_lib.train_step, batch 5000 ---------------------------------------------------------------------------
    if not args.skip_c2:
        rng = np.random.default_rng(0)

        def some_board():
            b = board_mod.Board2048()
            e = rng.integers(1, 12, size=(4, 4)) * (rng.random((4, 4)) > 0.3)
            b.state = np.where(e > 0, 2 ** e, 0).astype(np.int64)
            return b

        pool = [some_board() for _ in range(3000)]         # tuples share boards: sampling cost per tuple is unchanged

        def filled(n):
            d = collections.deque(maxlen=n)
            idx = rng.integers(0, len(pool), size=(n, 2))
            for i in range(n):
                d.append((pool[idx[i, 0]], int(rng.integers(4)), int(rng.integers(0, 64)) * 4, pool[idx[i, 1]], bool(rng.random() < 0.01)))
            return d

        for name, cfg_name, to_tensor, extract in (("conv", "configs.double_dqn_conv", dqn_lib.board_as_4d_tensor, dqn_lib.extract_samples_conv),
                                                   ("dense", "configs.double_dqn_dense", dqn_lib.board_as_flattened_tensor, dqn_lib.extract_samples_dense)):
            cfg = __import__(cfg_name, fromlist=["x"])
            buf = filled(cfg.replay_buffer_length)
            times = []
            for i in range(args.c2_calls + 1):
                if device != "cpu":
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
                loss = dqn_lib.train_step(cfg.batch_size, cfg.discount_factor, cfg.model, cfg.target_model, buf, cfg.loss_fn,
                                          cfg.optimizer, device, cfg.use_double_dqn, to_tensor, extract)
                float(loss)
                if device != "cpu":
                    torch.cuda.synchronize()
                if i:                                      # first call is a warm-up
                    times.append(time.perf_counter() - t0)
            out[f"c2_train_step_{name}"] = {"updates_per_sec": len(times) / sum(times), "seconds_per_update": sum(times) / len(times),
                                            "calls": len(times), "batch_size": cfg.batch_size, "replay": cfg.replay_buffer_length,
                                            "what": "dqn_lib.train_step, src/dqn_lib.py:119-164, the reference's own model/loss/optimizer"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
