"""CPU oracle for the dqn_lib arithmetic (epsilon-greedy, replay extraction, Double-DQN target, loss).

TEST INFRASTRUCTURE ONLY — see oracle/board_oracle.py for who may import this.

numpy restatement of ``src/dqn_lib.py`` with every rounding point written out (the reference
relies on torch type promotion; here the float32 gamma and the float64 products are explicit).
Pinned against ``tests/golden/dqn_*.npz`` (outputs of the reference's own ``sample_experiences``,
``epsilon_greedy_policy`` and ``train_step``) in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np


def egreedy_greedy_branch(q: np.ndarray, legal_bits: int) -> tuple[int, float]:
    """src/dqn_lib.py:23-30 for one board.  q: float64[4].  Returns (action, max_q).

    Q_n = Q - min(Q)*max(Q) - min(Q) evaluated left to right; illegal entries are multiplied by a
    0.0 mask (so they become +-0, NOT -inf); argmax takes the first maximum.
    """
    q = np.asarray(q, dtype=np.float64).reshape(4)
    mn, mx = q.min(), q.max()
    qn = (q - mn * mx) - mn
    mask = np.array([(legal_bits >> j) & 1 for j in range(4)], dtype=np.float64)
    avail = mask * qn
    return int(np.argmax(avail)), float(mx)


def egreedy_batch(q: np.ndarray, flags: np.ndarray, override: np.ndarray):
    """Batched oracle for egreedy_select with an explicit override byte per board
    (0x80 = greedy, 0..3 = that random action; include/b2048.h)."""
    n = q.shape[0]
    actions = np.zeros(n, dtype=np.uint8)
    max_q = np.zeros(n, dtype=np.float64)
    for i in range(n):
        ov = int(override[i])
        if ov & 0x80:
            a, m = egreedy_greedy_branch(q[i], int(flags[i]) & 0xF)
            actions[i], max_q[i] = a, m
        else:
            actions[i], max_q[i] = ov & 3, 0.0  # src/dqn_lib.py:20-21: random action, zeros
    return actions, max_q


def extract_samples(s_exp: np.ndarray, a: np.ndarray, r: np.ndarray, s2_exp: np.ndarray, d: np.ndarray,
                    idx: np.ndarray):
    """src/dqn_lib.py:67-84 + 33-64 given the sampled indices: gather the five fields; boards are
    already in log-scale (exponent) form [N,16] float64."""
    idx = np.asarray(idx, dtype=np.int64)
    return (s_exp[idx].astype(np.float64), a[idx].astype(np.int64), r[idx].astype(np.int64),
            s2_exp[idx].astype(np.float64), d[idx].astype(np.int64))


def ddqn_target_loss(q_next_online, q_next_target, q_cur, actions, rewards, dones, gamma: float,
                     use_double: bool = True):
    """src/dqn_lib.py:125-158.  Returns (target[B], q_sa[B], loss, grad_q_cur[B,4]).

    gamma reaches the product as float32: ``(1-dones) * discount_factor`` is int64-tensor x python
    float = float32 in torch; it is widened to float64 only when multiplied by Q.
    """
    qt = np.asarray(q_next_target, dtype=np.float64)
    qc = np.asarray(q_cur, dtype=np.float64)
    a = np.asarray(actions, dtype=np.int64)
    r = np.asarray(rewards, dtype=np.int64)
    d = np.asarray(dones, dtype=np.int64)
    B = qc.shape[0]
    rows = np.arange(B)
    if use_double:
        qo = np.asarray(q_next_online, dtype=np.float64)
        best = np.argmax(qo, axis=1)              # first index on ties (:127)
        nb = qt[rows, best]                       # one-hot mask + sum (:128-130)
    else:
        nb = qt.max(axis=1)                       # (:138-139)
    g32 = (1 - d).astype(np.float32) * np.float32(gamma)   # float32 product (:131 / :142)
    target = r.astype(np.float64) + g32.astype(np.float64) * nb
    q_sa = qc[rows, a]                            # one-hot mask + sum (:148-155)
    diff = q_sa - target
    loss = float(np.sum(diff * diff))             # MSELoss(reduction='sum') (:158)
    grad = np.zeros_like(qc)
    grad[rows, a] = 2.0 * diff
    return target, q_sa, loss, grad


# ---- the conv Q-network's forward (src/configs/double_dqn_conv.py:19-28) --------------------------------

def conv_q_forward(states: np.ndarray, w1, b1, w2, b2, w3, b3, w4, b4) -> np.ndarray:
    """Q-values of Conv2d(1,64,2) ReLU Conv2d(64,64,2) ReLU Flatten Linear(256,64) ReLU Linear(64,4) in
    float64 numpy: states [n,16] (exponents) -> [n,4].  Cross-correlation with valid padding, exactly as
    torch.nn.Conv2d; nn.Flatten order is (channel, y, x).  Checked against the Q tensors the reference's
    own train_step produced (tests/golden/dqn_conv.npz)."""
    n = states.shape[0]
    x = states.reshape(n, 4, 4)
    # conv1: out1[n, c, y, x] = b1[c] + sum_{ky,kx} x[n, y+ky, x+kx] * w1[c, 0, ky, kx]
    p1 = np.stack([x[:, ky:ky + 3, kx:kx + 3] for ky in range(2) for kx in range(2)], axis=-1)      # [n,3,3,4]
    out1 = np.maximum(p1 @ w1.reshape(64, 4).T + b1, 0.0)                                           # [n,3,3,64]
    # conv2: out2[n, c2, y, x] = b2[c2] + sum_{c1,ky,kx} out1[n, y+ky, x+kx, c1] * w2[c2, c1, ky, kx]
    p2 = np.stack([out1[:, ky:ky + 2, kx:kx + 2, :] for ky in range(2) for kx in range(2)], axis=-1)  # [n,2,2,64,4]
    out2 = np.maximum(p2.reshape(n, 2, 2, 256) @ w2.reshape(64, 256).T + b2, 0.0)                   # [n,2,2,64]
    flat = out2.transpose(0, 3, 1, 2).reshape(n, 256)                                               # (c, y, x)
    h = np.maximum(flat @ w3.T + b3, 0.0)
    return h @ w4.T + b4
