"""CPU: pin the oracle (C and pure-Python restatements) against the reference's own outputs
(tests/golden/*.npz, produced by oracle/gen_golden.py from /root/reference) and against the
known-answer vectors of the reference's tests/test_game_board.py."""
import os

import numpy as np
import pytest

from oracle import board_oracle as bo
from oracle import dqn_oracle as do

# reference tests/test_game_board.py:8-22 — the 15 row vectors, restated
REF_ROW_VECTORS = [
    ([0, 0, 0, 0], [0, 0, 0, 0]), ([0, 0, 0, 2], [2, 0, 0, 0]), ([0, 0, 2, 2], [4, 0, 0, 0]),
    ([2, 0, 0, 0], [2, 0, 0, 0]), ([2, 0, 2, 0], [4, 0, 0, 0]), ([2, 2, 2, 2], [4, 4, 0, 0]),
    ([2, 2, 4, 4], [4, 8, 0, 0]), ([2, 2, 0, 0], [4, 0, 0, 0]), ([2, 0, 0, 2], [4, 0, 0, 0]),
    ([0, 0, 2, 2], [4, 0, 0, 0]), ([2, 4, 2, 4], [2, 4, 2, 4]), ([2, 2, 4, 2], [4, 4, 2, 0]),
    ([2, 4, 4, 2], [2, 8, 2, 0]), ([2, 4, 4, 4], [2, 8, 4, 0]), ([4, 8, 16, 32], [4, 8, 16, 32]),
]
# reference tests/test_game_board.py:34-51 — the 3 legal-move boards, restated
REF_LEGAL_BOARDS = [
    ([[2, 4, 8, 0], [0, 0, 0, 0], [2, 4, 16, 32], [0, 0, 0, 0]], {"up", "down", "right"}),
    ([[2, 4, 2, 4]] * 4, {"up", "down"}),
    ([[2, 4, 2, 4], [4, 2, 4, 2], [2, 4, 2, 4], [4, 2, 4, 2]], set()),
]


@pytest.fixture(scope="module")
def g(golden_dir):
    return {k: np.load(os.path.join(golden_dir, k + ".npz")) for k in
            ("rows", "boards", "games", "dqn_conv", "dqn_dense", "egreedy", "dqn_dense_b5000", "bench_stream")}


def test_reference_row_vectors():
    for vec, want in REF_ROW_VECTORS:
        got_c, _ = bo.row_left(vec)
        got_py, _ = bo.py_row_left(vec)
        assert got_c.tolist() == want
        assert got_py.tolist() == want


def test_reference_legal_boards():
    for state, want in REF_LEGAL_BOARDS:
        s = np.array(state, dtype=np.int64)
        for mask in (bo.legal_mask(s), bo.py_legal_mask(s)):
            assert {bo.ACTIONS[a] for a in range(4) if mask >> a & 1} == want


def test_all_rows_match_reference(g):
    res, rew = g["rows"]["result"], g["rows"]["reward"]
    for row in range(65536):
        vec = [(1 << e) if e else 0 for e in ((row >> (4 * c)) & 0xF for c in range(4))]
        out, r = bo.row_left(vec)
        assert out.tolist() == res[row].tolist() and r == rew[row], row
    for row in np.random.default_rng(0).integers(0, 65536, size=3000):
        vec = [(1 << e) if e else 0 for e in ((int(row) >> (4 * c)) & 0xF for c in range(4))]
        out, r = bo.py_row_left(vec)
        assert out.tolist() == res[row].tolist() and r == rew[row], row


def test_boards_match_reference(g):
    b = g["boards"]
    st = b["state"]
    for i in range(len(st)):
        assert bo.legal_mask(st[i]) == b["legal"][i], i
        for a in range(4):
            out, r, ch = bo.slide(st[i], a)
            assert np.array_equal(out.reshape(16), b["slide"][i, a]), (i, a)
            assert r == b["reward"][i, a]
            assert ch == (b["spawn_cell"][i, a] >= 0)       # spawn iff changed (src/board.py:151-153)
            assert ch == bool(b["legal"][i] >> a & 1)
    for i in np.random.default_rng(1).integers(0, len(st), size=150):
        assert bo.py_legal_mask(st[i]) == b["legal"][i]
        for a in range(4):
            out, r, _ = bo.py_slide(st[i], a)
            assert np.array_equal(out.reshape(16), b["slide"][i, a]) and r == b["reward"][i, a]


def test_packed_step_replays_reference_games(g):
    """Whole reference episodes through the packed oracle with the spawn replayed via the hook:
    next board, reward and done (a property of the pre-action board) all match."""
    t = g["games"]
    S, A, R, S2, D = t["state"], t["action"], t["reward"], t["next"], t["done"]
    slide = np.array([bo.slide(S[i], int(A[i]))[0].reshape(16) for i in range(len(S))])
    diff = (slide != S2)
    assert (diff.sum(axis=1) <= 1).all()
    ovr = np.full(len(S), 0xFF, dtype=np.uint8)
    rows = np.nonzero(diff.any(axis=1))[0]
    cells = diff[rows].argmax(axis=1)
    vals = S2[rows, cells]
    ovr[rows] = (cells | (np.log2(vals).astype(np.int64) << 4)).astype(np.uint8)
    nxt, rew, flg = bo.step_packed(bo.pack(S), A, spawn_override=ovr)
    assert np.array_equal(nxt, bo.pack(S2))
    assert np.array_equal(rew, R.astype(np.int32))
    assert np.array_equal((flg & bo.F_DONE) != 0, D.astype(bool))
    assert np.array_equal((flg & bo.F_CHANGED) != 0, diff.any(axis=1))
    # every episode ends with exactly one done=1 no-op transition (src/dqn_lib.py:99-106)
    for gid in np.unique(t["game"]):
        d = D[t["game"] == gid]
        assert d[-1] == 1 and d[:-1].sum() == 0


def test_pack_unpack_roundtrip(g):
    st = g["boards"]["state"]
    ok = st.max(axis=1) <= 32768
    p = bo.pack(st[ok])
    assert np.array_equal(bo.unpack(p), st[ok])
    assert all(bo.py_pack(st[ok][i]) == int(p[i]) for i in range(0, ok.sum(), 97))


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in bo.philox4x32_10(ctr, key)) == want


def test_philox_spawn_rule_is_uniform_over_empty_cells():
    boards = bo.random_boards(20000, seed=3, p_empty=0.5)
    nxt, _, flg = bo.step_packed(boards, np.full(20000, 2, np.uint8), seed=11, step=5)
    before = bo.slide  # noqa: F841
    ch = (flg & bo.F_CHANGED) != 0
    t_next = bo.unpack(nxt[ch])
    # exactly one new tile (2 or 4) beyond the slid board, ~10 % fours
    slid = np.array([bo.slide(bo.unpack(boards[i:i + 1])[0], 2)[0].reshape(16) for i in np.nonzero(ch)[0][:3000]])
    d = t_next[:3000] - slid
    assert ((d != 0).sum(axis=1) == 1).all()
    vals = d[d != 0]
    assert set(np.unique(vals)) <= {2, 4}
    assert 0.06 < (vals == 4).mean() < 0.14


# ---- dqn_lib arithmetic ---------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["dqn_conv", "dqn_dense"])
def test_extract_samples_matches_reference(g, name):
    d = g[name]
    st, ac, rw, ns, dn = do.extract_samples(bo.exponents(bo.pack(d["buf_state"])), d["buf_action"],
                                            d["buf_reward"], bo.exponents(bo.pack(d["buf_next"])),
                                            d["buf_done"], d["idx"])
    assert np.array_equal(st, d["states"]) and np.array_equal(ns, d["next_states"])
    assert np.array_equal(ac, d["actions"]) and np.array_equal(rw, d["rewards"]) and np.array_equal(dn, d["dones"])


@pytest.mark.parametrize("name", ["dqn_conv", "dqn_dense"])
@pytest.mark.parametrize("tag", ["double", "single"])
def test_ddqn_target_loss_matches_reference(g, name, tag):
    d = g[name]
    target, q_sa, loss, _ = do.ddqn_target_loss(d["q_next_online"], d["q_next_target"], d["q_cur"], d["actions"],
                                                d["rewards"], d["dones"], float(d["gamma"]), tag == "double")
    # tolerance: 1e-9 relative (north_star); the reference's own Q tensors are fed in, so the
    # remaining difference is only summation order
    np.testing.assert_allclose(target, d[f"target_{tag}"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(q_sa, d[f"q_sa_{tag}"], rtol=1e-12, atol=0)
    assert abs(loss - float(d[f"loss_{tag}"])) <= 1e-9 * abs(float(d[f"loss_{tag}"]))
    # exact gamma would miss the 1e-9 gate: proves the float32 rounding is honoured
    g64 = d["rewards"] + (1 - d["dones"]) * float(d["gamma"]) * (
        d["q_next_target"][np.arange(len(d["actions"])), d["q_next_online"].argmax(1)] if tag == "double"
        else d["q_next_target"].max(1))
    t = d[f"target_{tag}"]
    assert (np.abs(g64 - t) > 1e-9 * np.abs(t)).any()


def test_ddqn_oracle_on_config3_batch_5000_gamma_095_and_080(g):
    """BASELINE.json config 3 as stated (dense Q-net, batch 5000, gamma 0.95; and the config module's 0.80):
    the reference rounds gamma to float32, and 0.95 rounds differently from 0.80, so each value has its own
    recorded targets.  Targets bit-identical, loss within 1e-12."""
    d = g["dqn_dense_b5000"]
    assert d["states"].shape == (5000, 16) and d["actions"].shape == (5000,)
    st, ac, rw, ns, dn = do.extract_samples(bo.exponents(bo.pack(d["buf_state"])), d["buf_action"], d["buf_reward"],
                                            bo.exponents(bo.pack(d["buf_next"])), d["buf_done"], d["idx"])
    assert np.array_equal(st, d["states"]) and np.array_equal(ns, d["next_states"]) and np.array_equal(rw, d["rewards"])
    for gamma, gtag in ((0.95, "g095"), (0.80, "g080")):
        for use_double in (True, False):
            tag = f"{gtag}_{'double' if use_double else 'single'}"
            target, q_sa, loss, _ = do.ddqn_target_loss(d["q_next_online"], d["q_next_target"], d["q_cur"], d["actions"],
                                                        d["rewards"], d["dones"], gamma, use_double)
            assert np.array_equal(target, d[f"target_{tag}"]) and np.array_equal(q_sa, d[f"q_sa_{tag}"])
            assert abs(loss - float(d[f"loss_{tag}"])) <= 1e-12 * abs(float(d[f"loss_{tag}"]))
    # float32 rounding of gamma is visible: a float64 gamma would give different targets for 0.95
    live = d["dones"] == 0
    wrong = d["rewards"] + 0.95 * d["q_next_target"].max(axis=1)
    assert not np.array_equal(wrong[live], d["target_g095_single"][live])


def test_bench_stream_head_matches_reference(g):
    """the first 65536 boards / actions of the BENCH stream (restated generators, seeds 2048 /
    2050) went through the reference's Board2048 (oracle/gen_golden.py gen_bench_stream).  The oracle must
    reproduce the inputs bit for bit and agree with the reference on slide-only successor, reward, legal mask."""
    d = g["bench_stream"]
    n = len(d["boards"])
    assert n == 65536
    assert np.array_equal(bo.stream_boards(n, seed=2048), d["boards"])
    assert np.array_equal(bo.stream_actions(n, seed=2050, step=0), d["actions"])
    skip = np.full(n, 0xFE, dtype=np.uint8)
    nxt, rew, flg = bo.step_packed(d["boards"], d["actions"], spawn_override=skip)
    assert np.array_equal(nxt, d["slide"]) and np.array_equal(rew, d["reward"])
    assert np.array_equal(flg & 0x0F, d["legal"]) and np.array_equal((flg & 0x10) != 0, d["legal"] == 0)
    assert np.array_equal(bo.legal_mask_packed(d["boards"]) & 0x0F, d["legal"])


def test_conv_q_forward_matches_reference_q_values(g):
    """The numpy restatement of the conv Q-network against the Q tensors of the reference's train_step
    (its own weights, its own sampled batch): pins the oracle the GPU kernel K6 is compared with."""
    d = g["dqn_conv"]
    on = [d[f"w_{k}"] for k in ("0.weight", "0.bias", "2.weight", "2.bias", "5.weight", "5.bias", "7.weight", "7.bias")]
    tg = [d[f"tw_{k}"] for k in ("0.weight", "0.bias", "2.weight", "2.bias", "5.weight", "5.bias", "7.weight", "7.bias")]
    for got, want in ((do.conv_q_forward(d["states"], *on), d["q_cur"]),
                      (do.conv_q_forward(d["next_states"], *on), d["q_next_online"]),
                      (do.conv_q_forward(d["next_states"], *tg), d["q_next_target"])):
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.abs(want).max())


def test_egreedy_matches_reference(g):
    e = g["egreedy"]
    acts, mq = do.egreedy_batch(e["q"], e["legal"], np.full(len(e["q"]), 0x80, np.uint8))
    assert np.array_equal(acts, e["action"])
    assert np.array_equal(mq, e["max_q"])
    assert np.array_equal(e["done"].astype(bool), (e["legal"] & 0xF) == 0)
    # the reference's quirk is present in the data: some greedy picks are illegal moves
    illegal = ((e["legal"] >> e["action"]) & 1) == 0
    assert illegal[e["legal"] != 0].any()
