#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

The reference is imported, never copied.  Its RNG cannot be replayed, so spawns
are captured by diffing the reference's own outputs: the same move is run once with
``_populate_empty_cell`` disabled (slide-only result) and once for real; the single differing cell
is the spawn.  Everything else (row moves, rewards, legal masks, done, sampled tensors, Q-targets,
loss, greedy actions) is recorded straight from reference calls.
"""
from __future__ import annotations

import os
import random
import sys
import warnings
from collections import deque

import numpy as np

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import torch  # noqa: E402

import board as ref_board  # noqa: E402
import dqn_lib as ref_dqn  # noqa: E402

Board2048 = ref_board.Board2048
torch.set_num_threads(4)


def mk(state) -> Board2048:
    b = Board2048(populate_empty_cells=False)
    b.state = np.array(state, dtype=np.int64).reshape(4, 4)
    return b


# --------------------------------------------------------------------------------------------
def gen_rows():
    """All 65536 rows of 4-bit exponents through Board2048._apply_action_to_vector."""
    res = np.zeros((65536, 4), dtype=np.int32)
    rew = np.zeros(65536, dtype=np.int32)
    b = Board2048(populate_empty_cells=False)
    for row in range(65536):
        e = [(row >> (4 * c)) & 0xF for c in range(4)]
        vec = np.array([(1 << x) if x else 0 for x in e], dtype=np.int64)
        b._mergescore = 0
        out = b._apply_action_to_vector(vec)
        res[row] = out
        rew[row] = b._mergescore
    np.savez_compressed(os.path.join(OUT, "rows.npz"), result=res, reward=rew)
    print("rows.npz", res.shape)


# --------------------------------------------------------------------------------------------
def random_state(rng, kind):
    if kind == "iid":          # SURVEY.md §8(d) distribution
        e = rng.integers(1, 12, size=16)
        e[rng.random(16) < 0.3] = 0
    elif kind == "sparse":
        e = rng.integers(1, 6, size=16)
        e[rng.random(16) < 0.8] = 0
    elif kind == "dense":      # full boards, few distinct values: many merges / dead boards
        e = rng.integers(1, 4, size=16)
    elif kind == "checker":    # dead or nearly dead boards
        a, c = rng.integers(1, 8, size=2)
        e = np.array([[a, c][(i // 4 + i % 4) % 2] for i in range(16)])
        if rng.random() < 0.5:
            e[rng.integers(0, 16)] = rng.integers(0, 8)
    elif kind == "high":       # tiles up to 32768 (nibble 14/15 rows: shared-memory table miss path)
        e = rng.integers(9, 16, size=16)
        e[rng.random(16) < 0.25] = 0
    else:
        raise ValueError(kind)
    return np.where(e > 0, 1 << e.astype(np.int64), 0).astype(np.int64)


def move_with_and_without_spawn(state, action):
    """Returns (slide-only state, reward, next state incl. spawn) from reference calls."""
    b = mk(state)
    orig = Board2048._populate_empty_cell
    Board2048._populate_empty_cell = lambda self: self
    try:
        nb = b.peek_action(action)
    finally:
        Board2048._populate_empty_cell = orig
    slide = nb.state.copy()
    reward = ref_dqn.reward_func_merge_score(b, nb, action, 0)
    real = mk(state).peek_action(action).state.copy()
    return slide, int(reward), real


def gen_boards(n_per_kind=600):
    rng = np.random.default_rng(20481)
    random.seed(1)
    np.random.seed(1)
    kinds = ["iid", "sparse", "dense", "checker", "high"]
    states = []
    for k in kinds:
        for _ in range(n_per_kind):
            s = random_state(rng, k)
            if (s != 0).sum() == 0:
                s[0] = 2
            states.append(s)
    # the 3 boards of the reference's tests/test_game_board.py:34-51
    states += [np.array([2, 4, 8, 0, 0, 0, 0, 0, 2, 4, 16, 32, 0, 0, 0, 0]),
               np.array([2, 4, 2, 4] * 4),
               np.array([2, 4, 2, 4, 4, 2, 4, 2, 2, 4, 2, 4, 4, 2, 4, 2])]
    N = len(states)
    inp = np.array(states, dtype=np.int64)
    slide = np.zeros((N, 4, 16), dtype=np.int64)
    nxt = np.zeros((N, 4, 16), dtype=np.int64)
    reward = np.zeros((N, 4), dtype=np.int64)
    legal = np.zeros(N, dtype=np.uint8)
    spawn_cell = np.full((N, 4), -1, dtype=np.int8)
    spawn_val = np.zeros((N, 4), dtype=np.int8)
    for i, s in enumerate(states):
        m = mk(s).available_moves_as_torch_unit_vector(device="cpu")      # src/board.py:128-135
        legal[i] = sum(1 << a for a in range(4) if float(m[a]) != 0)
        for a in range(4):
            sl, r, real = move_with_and_without_spawn(s, a)
            slide[i, a] = sl.reshape(16)
            nxt[i, a] = real.reshape(16)
            reward[i, a] = r
            d = np.nonzero(real.reshape(16) != sl.reshape(16))[0]
            assert len(d) <= 1
            if len(d) == 1:
                assert sl.reshape(16)[d[0]] == 0 and real.reshape(16)[d[0]] in (2, 4)
                spawn_cell[i, a] = d[0]
                spawn_val[i, a] = real.reshape(16)[d[0]]
        if i % 500 == 0:
            print("boards", i, "/", N)
    np.savez_compressed(os.path.join(OUT, "boards.npz"), state=inp, slide=slide, next=nxt,
                        reward=reward, legal=legal, spawn_cell=spawn_cell, spawn_val=spawn_val)
    print("boards.npz", N)


# --------------------------------------------------------------------------------------------
def gen_games(n_games=30, save=True, seed=7):
    """Full reference episodes through dqn_lib.play_one_step with epsilon = 1 (random actions,
    illegal no-ops included, final dead->dead transition with done=1; src/dqn_lib.py:91-107)."""
    random.seed(seed)
    np.random.seed(seed)
    S, A, R, S2, D, G = [], [], [], [], [], []
    buf = deque(maxlen=10 ** 6)
    for g in range(n_games):
        b = Board2048()
        done = False
        while not done:
            nb, action, reward, done, _ = ref_dqn.play_one_step(b, 1.0, None, buf, "cpu")
            S.append(b.state.reshape(16).copy())
            A.append(int(action))
            R.append(int(reward))
            S2.append(nb.state.reshape(16).copy())
            D.append(int(bool(done)))
            G.append(g)
            b = nb
    if not save:
        return buf
    np.savez_compressed(os.path.join(OUT, "games.npz"), state=np.array(S, dtype=np.int64),
                        action=np.array(A, dtype=np.uint8), reward=np.array(R, dtype=np.int64),
                        next=np.array(S2, dtype=np.int64), done=np.array(D, dtype=np.uint8),
                        game=np.array(G, dtype=np.int32))
    print("games.npz", len(S), "transitions")
    return buf


# --------------------------------------------------------------------------------------------
def conv_model():
    from torch import nn
    return nn.Sequential(nn.Conv2d(1, 64, kernel_size=2), nn.ReLU(), nn.Conv2d(64, 64, kernel_size=2),
                         nn.ReLU(), nn.Flatten(), nn.Linear(2 * 2 * 64, 64), nn.ReLU(),
                         nn.Linear(64, 4)).double()


def gen_dqn(buf: deque):
    """sample_experiences / train_step fixtures for the conv and dense configurations."""
    import configs.double_dqn_conv as cconv     # reference config modules (model definitions)
    import configs.double_dqn_dense as cdense
    import copy

    trans = list(buf)[:4000]
    rb = deque(trans, maxlen=len(trans))
    packed = dict(
        buf_state=np.array([t[0].state.reshape(16) for t in trans], dtype=np.int64),
        buf_action=np.array([t[1] for t in trans], dtype=np.uint8),
        buf_reward=np.array([int(t[2]) for t in trans], dtype=np.int64),
        buf_next=np.array([t[3].state.reshape(16) for t in trans], dtype=np.int64),
        buf_done=np.array([int(bool(t[4])) for t in trans], dtype=np.uint8),
    )
    for name, cfg, to_tensor, extract, B in (
            ("conv", cconv, ref_dqn.board_as_4d_tensor, ref_dqn.extract_samples_conv, 5000),
            ("dense", cdense, ref_dqn.board_as_flattened_tensor, ref_dqn.extract_samples_dense, 1000)):
        torch.manual_seed(1234)
        model = copy.deepcopy(cfg.model).cpu()
        for p in model.parameters():            # deterministic non-default weights
            p.data = torch.randn_like(p) * 0.05
        target = copy.deepcopy(model)
        for p in target.parameters():
            p.data = p.data + torch.randn_like(p) * 0.02
        out = dict(packed)
        out["gamma"] = np.float64(cfg.discount_factor)
        # indices the reference will draw (src/dqn_lib.py:68)
        np.random.seed(99)
        out["idx"] = np.random.randint(len(rb), size=B).astype(np.int64)
        np.random.seed(99)
        st, ac, rw, ns, dn = ref_dqn.sample_experiences(B, rb, "cpu", to_tensor, extract)
        out.update(states=st.numpy().reshape(B, 16), actions=ac.numpy(), rewards=rw.numpy(),
                   next_states=ns.numpy().reshape(B, 16), dones=dn.numpy())
        with torch.no_grad():
            out["q_next_online"] = model(ns).numpy()
            out["q_next_target"] = target(ns).numpy()
            out["q_cur"] = model(st).numpy()
        for use_double in (True, False):
            rec = {}
            base_loss = torch.nn.MSELoss(reduction="sum")

            def recording_loss(q, t, rec=rec):
                rec["q"] = q.detach().numpy().copy()
                rec["t"] = t.detach().numpy().copy()
                return base_loss(q, t)

            opt = torch.optim.Adam(model.parameters(), lr=1e-2)
            w_before = [p.detach().clone() for p in model.parameters()]
            np.random.seed(99)
            loss = ref_dqn.train_step(B, cfg.discount_factor, model, target, rb, recording_loss, opt,
                                      "cpu", use_double, to_tensor, extract)
            assert all(torch.equal(a, b) for a, b in zip(w_before, model.parameters()))  # Q1
            tag = "double" if use_double else "single"
            out[f"target_{tag}"] = rec["t"]
            out[f"q_sa_{tag}"] = rec["q"]
            out[f"loss_{tag}"] = np.float64(loss.item())
            for p in list(model.parameters()) + list(target.parameters()):
                p.grad = None
        if name == "conv":      # weights are small enough to ship (268 KB each)
            for k, v in model.state_dict().items():
                out["w_" + k] = v.numpy()
            for k, v in target.state_dict().items():
                out["tw_" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, f"dqn_{name}.npz"), **out)
        print(f"dqn_{name}.npz", "loss", out["loss_double"], out["loss_single"])


def gen_dqn_config3():
    """BASELINE.json config 3 as stated: dense Q-net, batch 5000, gamma 0.95 (and the config module's own
    0.80) through the reference's sample_experiences / train_step.  gamma is applied in float32 by the
    reference: 0.95 and 0.80 round differently, so each needs its own recorded targets.
    The 403 716 weights are not shipped (3.2 MB per network): the Q tensors the reference computed are."""
    import configs.double_dqn_dense as cdense
    import copy
    buf = gen_games(n_games=40, save=False, seed=11)
    trans = list(buf)[:6000]
    rb = deque(trans, maxlen=len(trans))
    B = 5000
    torch.manual_seed(4321)
    model = copy.deepcopy(cdense.model).cpu()
    for p in model.parameters():
        p.data = torch.randn_like(p) * 0.05
    target = copy.deepcopy(model)
    for p in target.parameters():
        p.data = p.data + torch.randn_like(p) * 0.02
    out = dict(
        buf_state=np.array([t[0].state.reshape(16) for t in trans], dtype=np.int64),
        buf_action=np.array([t[1] for t in trans], dtype=np.uint8),
        buf_reward=np.array([int(t[2]) for t in trans], dtype=np.int64),
        buf_next=np.array([t[3].state.reshape(16) for t in trans], dtype=np.int64),
        buf_done=np.array([int(bool(t[4])) for t in trans], dtype=np.uint8),
    )
    np.random.seed(2051)
    out["idx"] = np.random.randint(len(rb), size=B).astype(np.int64)
    np.random.seed(2051)
    st, ac, rw, ns, dn = ref_dqn.sample_experiences(B, rb, "cpu", ref_dqn.board_as_flattened_tensor,
                                                    ref_dqn.extract_samples_dense)
    out.update(states=st.numpy().reshape(B, 16), actions=ac.numpy(), rewards=rw.numpy(),
               next_states=ns.numpy().reshape(B, 16), dones=dn.numpy())
    with torch.no_grad():
        out["q_next_online"] = model(ns).numpy()
        out["q_next_target"] = target(ns).numpy()
        out["q_cur"] = model(st).numpy()
    for gamma, gtag in ((0.95, "g095"), (0.80, "g080")):
        for use_double in (True, False):
            rec = {}
            base_loss = torch.nn.MSELoss(reduction="sum")

            def recording_loss(q, t, rec=rec):
                rec["q"] = q.detach().numpy().copy()
                rec["t"] = t.detach().numpy().copy()
                return base_loss(q, t)

            opt = torch.optim.Adam(model.parameters(), lr=1e-2)
            np.random.seed(2051)
            loss = ref_dqn.train_step(B, gamma, model, target, rb, recording_loss, opt, "cpu", use_double,
                                      ref_dqn.board_as_flattened_tensor, ref_dqn.extract_samples_dense)
            tag = f"{gtag}_{'double' if use_double else 'single'}"
            out[f"target_{tag}"] = rec["t"]
            out[f"q_sa_{tag}"] = rec["q"]
            out[f"loss_{tag}"] = np.float64(loss.item())
            for p in list(model.parameters()) + list(target.parameters()):
                p.grad = None
    np.savez_compressed(os.path.join(OUT, "dqn_dense_b5000.npz"), **out)
    print("dqn_dense_b5000.npz", {k: float(v) for k, v in out.items() if k.startswith("loss_")})


def gen_bench_stream(n=65536):
    """SURVEY 8(d) 'parity subset for every run': the first 65536 boards and actions of the BENCH stream
    (b2048_random_boards seed 2048 / b2048_random_actions seed 2050, restated in oracle/board_oracle.c) through
    the reference's Board2048: slide-only successor, merge-score reward, legal mask.  __graft_entry__.smoke()
    and tests/test_env_gpu.py compare the CUDA path with these, spawn switched off through the override hook."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import board_oracle as bo
    boards = bo.stream_boards(n, seed=2048)
    actions = bo.stream_actions(n, seed=2050, step=0)
    tiles = bo.unpack(boards)
    slide = np.zeros(n, dtype=np.uint64)
    reward = np.zeros(n, dtype=np.int32)
    legal = np.zeros(n, dtype=np.uint8)
    sh = 4 * np.arange(16, dtype=np.uint64)
    for i in range(n):
        s = tiles[i]
        m = mk(s).available_moves_as_torch_unit_vector(device="cpu")      # src/board.py:128-135
        legal[i] = sum(1 << a for a in range(4) if float(m[a]) != 0)
        sl, r, _ = move_with_and_without_spawn(s, int(actions[i]))
        e = np.where(sl.reshape(16) > 0, np.log2(np.maximum(sl.reshape(16), 1)).astype(np.uint64), 0).astype(np.uint64)
        slide[i] = (e << sh).sum(dtype=np.uint64)
        reward[i] = r
        if i % 8192 == 0:
            print("bench stream", i, "/", n)
    np.savez_compressed(os.path.join(OUT, "bench_stream.npz"), boards=boards, actions=actions, slide=slide,
                        reward=reward, legal=legal)
    print("bench_stream.npz", n, "legal histogram", np.bincount(legal, minlength=16).tolist())


def gen_egreedy(n_model=300, n_synth=3000):
    """epsilon_greedy_policy (greedy branch, src/dqn_lib.py:23-30) on boards with real and
    synthetic Q-values (ties, all-negative and all-positive rows, illegal best moves)."""
    rng = np.random.default_rng(5)
    torch.manual_seed(5)
    model = conv_model()
    g = np.load(os.path.join(OUT, "boards.npz"))
    states = g["state"]
    S, Q, A, Dn, M, L = [], [], [], [], [], []
    pick = rng.choice(len(states), size=n_model + n_synth, replace=True)
    for j, i in enumerate(pick):
        b = mk(states[i])
        if j < n_model:
            fn = model
            with torch.no_grad():
                q = model(ref_dqn.board_as_4d_tensor(b, "cpu")).numpy().reshape(4)
        else:
            kind = j % 5
            if kind == 0:
                q = rng.normal(size=4)
            elif kind == 1:
                q = -np.abs(rng.normal(size=4)) - 0.1
            elif kind == 2:
                q = np.abs(rng.normal(size=4)) * 10
            elif kind == 3:
                q = rng.integers(-2, 3, size=4).astype(np.float64)      # ties
            else:
                q = rng.normal(size=4) * 1e3
            qt = torch.tensor(q, dtype=torch.float64).reshape(1, 4)
            fn = lambda state, qt=qt: qt  # noqa: E731
        with torch.no_grad():
            a, done, mq = ref_dqn.epsilon_greedy_policy(b, 0.0, fn, "cpu", ref_dqn.board_as_4d_tensor)
        S.append(states[i]); Q.append(q); A.append(a); Dn.append(int(done)); M.append(float(mq))
        L.append(int(g["legal"][i]))
    np.savez_compressed(os.path.join(OUT, "egreedy.npz"), state=np.array(S, dtype=np.int64),
                        q=np.array(Q, dtype=np.float64), action=np.array(A, dtype=np.uint8),
                        done=np.array(Dn, dtype=np.uint8), max_q=np.array(M, dtype=np.float64),
                        legal=np.array(L, dtype=np.uint8))
    print("egreedy.npz", len(S))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["rows", "boards", "games", "dqn", "egreedy"]
    if "rows" in which:
        gen_rows()
    if "boards" in which:
        gen_boards()
    buf = None
    if "games" in which or "dqn" in which:
        buf = gen_games()
    if "dqn" in which:
        gen_dqn(buf)
    if "egreedy" in which:
        gen_egreedy()
    if "config3" in which:
        gen_dqn_config3()
    if "bench_stream" in which:
        gen_bench_stream()
