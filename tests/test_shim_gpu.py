"""GPU tests of the drop-in `board` / `dqn_lib` modules: the reference's own test vectors
(tests/test_game_board.py restated), Board2048 semantics against the oracle, and the dqn_lib entry
points against the reference's recorded outputs (tests/golden/*.npz)."""
import copy
import os
import types

import numpy as np
import pytest
import torch

from oracle import board_oracle as bo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim(cuda):
    import board
    import dqn_lib
    board.seed(1234)
    return types.SimpleNamespace(board=board, dqn=dqn_lib)


def conv_model():
    from torch import nn   # the architecture of the reference's configs/double_dqn_conv.py:19-28
    return nn.Sequential(nn.Conv2d(1, 64, kernel_size=2), nn.ReLU(), nn.Conv2d(64, 64, kernel_size=2), nn.ReLU(),
                         nn.Flatten(), nn.Linear(2 * 2 * 64, 64), nn.ReLU(), nn.Linear(64, 4)).double()


def dense_model():
    from torch import nn   # configs/double_dqn_dense.py:7-15
    return nn.Sequential(nn.Linear(16, 512), nn.ReLU(), nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, 256),
                         nn.ReLU(), nn.Linear(256, 4)).double()


# ---- the reference's tests/test_game_board.py, restated against the shim -----------------------------

def test_reference_row_vectors_through_shim(shim):
    from test_oracle_golden import REF_ROW_VECTORS
    b = shim.board.Board2048()
    for vec, want in REF_ROW_VECTORS:
        assert np.array_equal(b._apply_action_to_vector(np.array(vec)), want)


def test_reference_available_moves_through_shim(shim):
    from test_oracle_golden import REF_LEGAL_BOARDS
    b = shim.board.Board2048(k=4, populate_empty_cells=False)
    for configuration, possible_moves in REF_LEGAL_BOARDS:
        b.state = np.array(configuration)                    # `state` stays assignable
        assert set(b.available_moves().keys()) == possible_moves
        unit = b.available_moves_as_torch_unit_vector(device="cpu")
        assert unit.dtype == torch.float32 and unit.shape == (4,)
        assert {m for m, u in zip(("up", "down", "left", "right"), unit.tolist()) if u} == possible_moves


# ---- Board2048 semantics ---------------------------------------------------------------------------

def test_board_api_surface(shim):
    B = shim.board.Board2048
    b = B()
    assert b.state.shape == (4, 4) and (b.state != 0).sum() == 2 and set(np.unique(b.state)) <= {0, 2, 4}
    assert b.k == 4 and b._mergescore == 0 and b._action_history == [] and b._empty_spot_numbers == [2, 4]
    e = B(populate_empty_cells=False)
    assert (e.state == 0).all() and e.number_of_empty_cells() == 16
    with pytest.raises(NotImplementedError):
        B(k=5)
    with pytest.raises(ValueError):
        b.peek_action("x")
    c = b.clone()
    assert c == b and c is not b and c.state is not b.state
    assert 2 in b or 4 in b
    for a in ("up", "U", "d", 2, np.int64(3), torch.tensor(1)):      # str / int / 0-d tensor actions
        assert isinstance(b.peek_action(a), B)
    nb = b.peek_action("left")
    assert nb._action_history == ["left"] and b._action_history == []      # boards are immutable by convention
    ls = b.log_scale()
    assert np.array_equal(ls.state, np.where(b.state > 0, np.log2(np.maximum(b.state, 1)), 0).astype(int))
    assert b.flattened_state_as_tensor().dtype == torch.float64 and b.flattened_state_as_tensor().shape == (16,)
    assert b.state_as_4d_tensor().shape == (1, 1, 4, 4)
    assert np.isclose(b.normalized().state.max(), 1.0)
    assert b.simple_score() == b.state.sum()
    p = e._populate_empty_cell()
    assert p is e and (e.state != 0).sum() == 1


def test_random_games_match_oracle_move_by_move(shim):
    """Whole games through the shim: every transition equals the oracle's slide + exactly one
    spawned 2/4 in a previously empty cell; merge score accumulates the oracle's rewards."""
    rng = np.random.default_rng(0)
    n_steps = 0
    for g in range(6):
        b = shim.board.Board2048()
        while True:
            mask = bo.legal_mask(b.state)
            unit = b.available_moves_as_torch_unit_vector()
            assert [int(x) for x in unit.tolist()] == [(mask >> i) & 1 for i in range(4)]
            if mask == 0:
                break
            a = int(rng.integers(0, 4))
            nb = b.peek_action(a)
            slid, reward, changed = bo.slide(b.state, a)
            assert nb.merge_score() - b.merge_score() == reward
            d = nb.state - slid
            if changed:
                assert (d != 0).sum() == 1 and d[d != 0][0] in (2, 4) and slid[d != 0][0] == 0
            else:
                assert (d == 0).all()
            b = nb
            n_steps += 1
    assert n_steps > 300


def test_spawn_four_probability(shim):
    shim.board.set_spawn_four_probability(0.5)
    vals = []
    for _ in range(300):
        b = shim.board.Board2048()
        vals += list(b.state[b.state != 0])
    frac = np.mean(np.array(vals) == 4)
    assert 0.4 < frac < 0.6                                   # the reference's 50/50


# ---- dqn_lib ----------------------------------------------------------------------------------------

class FakeBoard:
    """Minimal stand-in with the attribute dqn_lib reads from replay entries."""
    def __init__(self, state):
        self.state = np.asarray(state).reshape(4, 4)


def test_epsilon_greedy_matches_reference(shim, golden_dir):
    e = np.load(os.path.join(golden_dir, "egreedy.npz"))
    B = shim.board.Board2048
    for i in list(range(0, 300, 7)) + list(range(300, 3300, 41)):
        b = B(populate_empty_cells=False)
        b.state = e["state"][i].reshape(4, 4)
        q = torch.tensor(e["q"][i], dtype=torch.float64, device="cuda").reshape(1, 4)
        action, done, mq = shim.dqn.epsilon_greedy_policy(b, 0.0, lambda s, q=q: q, "cuda:0")
        assert action == int(e["action"][i]) and int(done) == int(e["done"][i])
        assert float(mq) == float(e["max_q"][i])
    # random branch consumes numpy's global RNG exactly like the reference (rand, then randint)
    b = B()
    np.random.seed(5)
    a, done, mq = shim.dqn.epsilon_greedy_policy(b, 1.0, None, "cuda:0")
    np.random.seed(5)
    np.random.rand()
    assert a == np.random.randint(4) and float(mq) == 0.0 and mq.shape == (1,)


@pytest.mark.parametrize("name,B", [("dqn_conv", 5000), ("dqn_dense", 1000)])
def test_sample_experiences_matches_reference(shim, golden_dir, name, B):
    """Same buffer + same numpy seed -> the same five tensors as the reference's
    sample_experiences (values, shapes, dtypes)."""
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    n = len(d["buf_action"])
    buf = shim.dqn.ReplayDeque(maxlen=n)
    for i in range(n):
        buf.append((FakeBoard(d["buf_state"][i]), int(d["buf_action"][i]), int(d["buf_reward"][i]),
                    FakeBoard(d["buf_next"][i]), bool(d["buf_done"][i])))
    assert len(buf) == n
    conv = name == "dqn_conv"
    np.random.seed(99)
    st, ac, rw, ns, dn = shim.dqn.sample_experiences(
        B, buf, "cuda:0", shim.dqn.board_as_4d_tensor if conv else shim.dqn.board_as_flattened_tensor,
        shim.dqn.extract_samples_conv if conv else shim.dqn.extract_samples_dense)
    assert st.shape == ((B, 1, 4, 4) if conv else (B, 16)) and st.dtype == torch.float64
    assert np.array_equal(st.cpu().numpy().reshape(B, 16), d["states"])
    assert np.array_equal(ns.cpu().numpy().reshape(B, 16), d["next_states"])
    assert np.array_equal(ac.cpu().numpy(), d["actions"]) and ac.dtype == torch.int64
    assert np.array_equal(rw.cpu().numpy(), d["rewards"]) and np.array_equal(dn.cpu().numpy(), d["dones"])
    # the plain-deque path (a user-supplied deque of tuples) gives the same tensors
    from collections import deque
    dq = deque([(FakeBoard(d["buf_state"][i]), int(d["buf_action"][i]), int(d["buf_reward"][i]),
                 FakeBoard(d["buf_next"][i]), bool(d["buf_done"][i])) for i in range(n)], maxlen=n)
    np.random.seed(99)
    st2, ac2, rw2, ns2, dn2 = shim.dqn.sample_experiences(
        256, dq, "cuda:0", shim.dqn.board_as_4d_tensor if conv else shim.dqn.board_as_flattened_tensor,
        shim.dqn.extract_samples_conv if conv else shim.dqn.extract_samples_dense)
    assert np.array_equal(st2.cpu().numpy().reshape(256, 16), d["states"][:256])
    assert np.array_equal(ac2.cpu().numpy(), d["actions"][:256])


@pytest.mark.parametrize("use_double", [True, False])
def test_train_step_matches_reference_loss(shim, golden_dir, use_double):
    """End to end on the GPU with the reference's weights, buffer and numpy seed: the loss equals
    the reference train_step's within 1e-9 relative (north_star), and — like the reference — the
    weights do not move unless FIX_UPDATE_ORDER is set."""
    d = np.load(os.path.join(golden_dir, "dqn_conv.npz"))
    model, target = conv_model(), conv_model()
    model.load_state_dict({k[2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("w_")})
    target.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("tw_")})
    model, target = model.cuda(), target.cuda()
    n = len(d["buf_action"])
    buf = shim.dqn.ReplayDeque(maxlen=n)
    for i in range(n):
        buf.append((FakeBoard(d["buf_state"][i]), int(d["buf_action"][i]), int(d["buf_reward"][i]),
                    FakeBoard(d["buf_next"][i]), bool(d["buf_done"][i])))
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    before = [p.detach().clone() for p in model.parameters()]
    np.random.seed(99)
    loss = shim.dqn.train_step(5000, float(d["gamma"]), model, target, buf, torch.nn.MSELoss(reduction="sum"), opt,
                               "cuda:0", use_double, shim.dqn.board_as_4d_tensor, shim.dqn.extract_samples_conv)
    ref = float(d["loss_double" if use_double else "loss_single"])
    assert abs(loss.item() - ref) <= 1e-9 * abs(ref), (loss.item(), ref)
    assert all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
    # a non-MSE loss goes through the generic path with kernel-computed targets
    np.random.seed(99)
    l1 = shim.dqn.train_step(5000, float(d["gamma"]), model, target, buf, torch.nn.L1Loss(reduction="sum"), opt,
                             "cuda:0", use_double, shim.dqn.board_as_4d_tensor, shim.dqn.extract_samples_conv)
    tgt = d["target_double" if use_double else "target_single"]
    want = np.abs(d["q_sa_double" if use_double else "q_sa_single"] - tgt).sum()
    assert abs(l1.item() - want) <= 1e-9 * want
    shim.dqn.FIX_UPDATE_ORDER = True
    try:
        np.random.seed(99)
        shim.dqn.train_step(5000, float(d["gamma"]), model, target, buf, torch.nn.MSELoss(reduction="sum"), opt,
                            "cuda:0", use_double, shim.dqn.board_as_4d_tensor, shim.dqn.extract_samples_conv)
        assert not all(torch.equal(a, b) for a, b in zip(before, model.parameters()))
    finally:
        shim.dqn.FIX_UPDATE_ORDER = False


class FakeExperiment:
    """Duck-typed stand-in for the reference's Experiment (src/experiments.py:112-148)."""
    folder = "fake"

    def __init__(self):
        self.episodes, self.snapshots, self.saves = [], [], 0

    def add_episode(self, board, epsilon, ep, mean_reward, mean_q):
        self.episodes.append((int(np.max(board.state)), board.merge_score(), ep, epsilon, len(board._action_history)))

    def snapshot_game(self, history, ep):
        self.snapshots.append((ep, len(history)))

    def save(self):
        self.saves += 1


@pytest.mark.parametrize("kind", ["conv", "dense"])
def test_training_loop_runs_like_the_drivers_call_it(shim, kind, capsys):
    """The 20 positional arguments of double_dqn_conv.py / double_dqn_dense.py (:42-63)."""
    torch.manual_seed(0)
    np.random.seed(0)
    model = (conv_model() if kind == "conv" else dense_model()).cuda()
    target = copy.deepcopy(model)
    exp = FakeExperiment()
    to_tensor = shim.dqn.board_as_4d_tensor if kind == "conv" else shim.dqn.board_as_flattened_tensor
    extract = shim.dqn.extract_samples_conv if kind == "conv" else shim.dqn.extract_samples_dense
    shim.dqn.training_loop(500, 4, 2, 0, 0.01, model, shim.dqn.reward_func_merge_score, to_tensor, "cuda:0", exp, 2,
                           1, 64, 0.8, target, torch.nn.MSELoss(reduction="sum"),
                           torch.optim.Adam(model.parameters(), lr=1e-2), True, 2, extract)
    assert [e[2] for e in exp.episodes] == [0, 1, 2, 3]
    assert [e[3] for e in exp.episodes] == [1.0, 0.5, 0.01, 0.01]           # epsilon schedule
    assert [s[0] for s in exp.snapshots] == [0, 2] and exp.saves == 2        # ep 0 (ep % 1000) + final
    out = capsys.readouterr().out
    assert "Episode: 0:" in out and "Saved game" in out
    # an exception inside the loop saves the experiment and propagates (src/dqn_lib.py:241-244)
    exp2 = FakeExperiment()
    with pytest.raises(ZeroDivisionError):
        shim.dqn.training_loop(500, 2, 0, 0, 0.01, model, shim.dqn.reward_func_merge_score, to_tensor, "cuda:0", exp2,
                               2, 1, 64, 0.8, target, torch.nn.MSELoss(reduction="sum"),
                               torch.optim.Adam(model.parameters(), lr=1e-2), True, 2, extract)
    assert exp2.saves == 1


def _astar_like_buffer(shim, n=300, maxlen=1000):
    """A deque shaped like generate_replay_buffer_using_A_star's (src/state_space_search.py:104-131): real
    Board2048 objects, the SAME board object as state and next state (:128), rewards computed child-minus-
    parent style (so some are negative), `done` as an int."""
    from collections import deque
    rng = np.random.default_rng(7)
    dq = deque(maxlen=maxlen)
    b = shim.board.Board2048()
    for i in range(n):
        nb = b.peek_action(int(rng.integers(4)))
        if not nb.available_moves():
            nb = shim.board.Board2048()
        reward = int(b.merge_score()) - int(nb.merge_score())          # <= 0, like reward(current, parent)
        dq.append((nb, int(rng.integers(4)), reward, nb, int(i % 50 == 0)))
        b = nb
    return dq


def test_replay_buffer_override_is_ingested_like_the_plain_deque(shim, monkeypatch, capsys):
    """SURVEY 8(f) rank 3 / src/dqn_lib.py:169-170: training_loop(..., replay_buffer_override=deque) takes an
    A*-seeded deque of (Board2048, action, reward, Board2048, done) tuples.  The GPU ring built from it
    samples exactly the tensors the plain deque gives through the same numpy draw, keeps deque(maxlen)
    semantics while the loop appends to it, and feeds train_step."""
    from collections import deque
    dq = _astar_like_buffer(shim)
    assert any(e[2] < 0 for e in dq) and all(e[0] is e[3] for e in dq)
    ring = shim.dqn._make_replay_buffer(15000, dq, True, "cuda:0")
    assert isinstance(ring, shim.dqn.ReplayDeque) and len(ring) == len(dq) and ring.maxlen == dq.maxlen
    for conv in (True, False):
        to_tensor = shim.dqn.board_as_4d_tensor if conv else shim.dqn.board_as_flattened_tensor
        extract = shim.dqn.extract_samples_conv if conv else shim.dqn.extract_samples_dense
        np.random.seed(5)
        got = shim.dqn.sample_experiences(512, ring, "cuda:0", to_tensor, extract)
        np.random.seed(5)
        want = shim.dqn.sample_experiences(512, deque(dq, maxlen=dq.maxlen), "cuda:0", to_tensor, extract)
        for g, w in zip(got, want):
            assert g.dtype == w.dtype and g.shape == w.shape and torch.equal(g, w)
        assert torch.equal(got[0], got[3]) and bool((got[2] < 0).any())       # same board twice, negative rewards kept
    # the whole loop with the override: the ring is what train_step sees, it grows by the played moves
    seen = []
    real_train_step = shim.dqn.train_step

    def spy(batch_size, discount_factor, model, target_model, replay_buffer, *a, **kw):
        seen.append((type(replay_buffer).__name__, len(replay_buffer)))
        return real_train_step(batch_size, discount_factor, model, target_model, replay_buffer, *a, **kw)
    monkeypatch.setattr(shim.dqn, "train_step", spy)
    torch.manual_seed(0)
    np.random.seed(0)
    model = conv_model().cuda()
    exp = FakeExperiment()
    shim.dqn.training_loop(15000, 3, 2, 0, 0.01, model, shim.dqn.reward_func_merge_score, shim.dqn.board_as_4d_tensor,
                           "cuda:0", exp, 2, 0, 128, 0.8, copy.deepcopy(model), torch.nn.MSELoss(reduction="sum"),
                           torch.optim.Adam(model.parameters(), lr=1e-2), True, 2, shim.dqn.extract_samples_conv,
                           replay_buffer_override=dq)
    capsys.readouterr()
    assert [s[0] for s in seen] == ["ReplayDeque", "ReplayDeque"]               # episodes 1 and 2 train
    # the ingested transitions are still there and every played step appended one more (up to maxlen)
    assert len(dq) + 2 <= seen[0][1] <= dq.maxlen and seen[0][1] <= seen[1][1] <= dq.maxlen
    assert seen[1][1] > seen[0][1] or seen[1][1] == dq.maxlen


def test_one_hot_and_helpers(shim):
    t = torch.tensor([0, 3, 1], device="cuda")
    oh = shim.dqn.one_hot(t, 4, "cuda:0")
    assert oh.dtype == torch.float32 and oh.tolist() == [[1, 0, 0, 0], [0, 0, 0, 1], [0, 1, 0, 0]]
    with pytest.raises(AssertionError):
        shim.dqn.one_hot(torch.tensor([4], device="cuda"), 4, "cuda:0")
    b = shim.board.Board2048()
    assert shim.dqn.board_as_4d_tensor(b, "cuda:0").shape == (1, 1, 4, 4)
    assert shim.dqn.board_as_flattened_tensor(b, "cuda:0").shape == (16,)
    assert np.array_equal(shim.dqn.board_as_flattened_tensor(b, "cpu").numpy(), b.log_scale().state.flatten())
