"""GPU parity tests for K2 (replay ring), K3 (fused DDQN target/loss) and K0 (epsilon-greedy)
against the reference's own outputs (tests/golden/dqn_*.npz, egreedy.npz) and the numpy oracle."""
import os

import numpy as np
import pytest
import torch

import b2048
from b2048 import ddqn, env
from oracle import board_oracle as bo
from oracle import dqn_oracle as do

pytestmark = pytest.mark.gpu


def dev_boards(tiles, cuda):
    return torch.from_numpy(bo.pack(tiles).view(np.int64)).to(cuda)


@pytest.fixture(scope="module", params=["dqn_conv", "dqn_dense"])
def dq(request, golden_dir):
    return np.load(os.path.join(golden_dir, request.param + ".npz"))


def fill_ring(d, cuda, capacity=None):
    n = len(d["buf_action"])
    ring = b2048.ReplayRing(capacity or n, device=cuda)
    ring.append(dev_boards(d["buf_state"], cuda), torch.from_numpy(d["buf_action"]).to(cuda),
                torch.from_numpy(d["buf_reward"].astype(np.int32)).to(cuda), dev_boards(d["buf_next"], cuda),
                torch.from_numpy(d["buf_done"]).to(cuda), done_is_bool=True)
    return ring


def test_replay_sample_matches_reference(cuda, dq):
    """Same deque contents + the reference's sampled indices -> identical tensors (values, dtypes,
    layout) to dqn_lib.sample_experiences (src/dqn_lib.py:67-84)."""
    d = dq
    ring = fill_ring(d, cuda)
    assert len(ring) == len(d["buf_action"])
    st, ac, rw, ns, dn = ring.sample(len(d["idx"]), idx_override=torch.from_numpy(d["idx"]).to(cuda))
    assert st.dtype == torch.float64 and ac.dtype == rw.dtype == dn.dtype == torch.int64
    assert np.array_equal(st.cpu().numpy(), d["states"]) and np.array_equal(ns.cpu().numpy(), d["next_states"])
    assert np.array_equal(ac.cpu().numpy(), d["actions"]) and np.array_equal(rw.cpu().numpy(), d["rewards"])
    assert np.array_equal(dn.cpu().numpy(), d["dones"])


def test_replay_ring_wraps_like_deque(cuda, dq):
    """deque(maxlen) semantics: oldest entries fall out, logical index 0 is the oldest survivor,
    also when one append is larger than the capacity and when appends arrive in ragged pieces."""
    d = dq
    n = len(d["buf_action"])
    cap = 1000
    ring = b2048.ReplayRing(cap, device=cuda)
    s, a = dev_boards(d["buf_state"], cuda), torch.from_numpy(d["buf_action"]).to(cuda)
    r, s2 = torch.from_numpy(d["buf_reward"].astype(np.int32)).to(cuda), dev_boards(d["buf_next"], cuda)
    dn = torch.from_numpy(d["buf_done"]).to(cuda)
    cuts = [0, 1, 300, 999, 1000, 1700, n - 1500, n]           # last piece is 1500 > capacity
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        ring.append(s[lo:hi].contiguous(), a[lo:hi].contiguous(), r[lo:hi].contiguous(), s2[lo:hi].contiguous(),
                    dn[lo:hi].contiguous(), done_is_bool=True)
        assert len(ring) == min(hi, cap)
    idx = torch.arange(cap, device=cuda)
    st, ac, rw, ns, dd = ring.sample(cap, idx_override=idx)
    want = slice(n - cap, n)
    assert np.array_equal(st.cpu().numpy(), bo.exponents(bo.pack(d["buf_state"][want])))
    assert np.array_equal(ns.cpu().numpy(), bo.exponents(bo.pack(d["buf_next"][want])))
    assert np.array_equal(ac.cpu().numpy(), d["buf_action"][want]) and np.array_equal(rw.cpu().numpy(), d["buf_reward"][want])
    assert np.array_equal(dd.cpu().numpy(), d["buf_done"][want])


def test_replay_philox_sampling_is_uniform_and_reproducible(cuda, dq):
    ring = fill_ring(dq, cuda)
    n = len(ring)
    a = ring.sample(200000, seed=5, ctr=1, return_idx=True)
    b = ring.sample(200000, seed=5, ctr=1, return_idx=True)
    c = ring.sample(200000, seed=5, ctr=2, return_idx=True)
    assert torch.equal(a[5], b[5]) and not torch.equal(a[5], c[5])
    idx = a[5].cpu().numpy()
    assert idx.min() >= 0 and idx.max() < n
    hist = np.bincount(idx, minlength=n)
    assert abs(hist.mean() - 200000 / n) < 1e-9 and hist.std() < 3 * np.sqrt(200000 / n)   # with replacement
    # flags bytes from the step kernel are accepted as the done field
    ring2 = b2048.ReplayRing(16, device=cuda)
    z = torch.zeros(4, dtype=torch.int64, device=cuda)
    ring2.append(z, torch.zeros(4, dtype=torch.uint8, device=cuda), torch.zeros(4, dtype=torch.int32, device=cuda), z,
                 torch.tensor([0x01, 0x10, 0x2F, 0x30], dtype=torch.uint8, device=cuda))
    assert ring2.sample(4, idx_override=torch.arange(4, device=cuda))[4].tolist() == [0, 1, 0, 1]


@pytest.mark.parametrize("tag", ["double", "single"])
def test_ddqn_target_loss_matches_reference(cuda, dq, tag):
    """Targets, Q(s,a) and summed-MSE loss vs the reference train_step (src/dqn_lib.py:119-164) fed
    the same Q tensors: 1e-9 relative (north_star); targets in fact come out bit-identical."""
    d = dq
    t = lambda k: torch.from_numpy(d[k]).to(cuda)  # noqa: E731
    loss, target, q_sa, grad = ddqn.ddqn_target_loss(t("q_next_online"), t("q_next_target"), t("q_cur"), t("actions"),
                                                     t("rewards"), t("dones"), float(d["gamma"]), tag == "double")
    assert np.array_equal(target.cpu().numpy(), d[f"target_{tag}"])
    assert np.array_equal(q_sa.cpu().numpy(), d[f"q_sa_{tag}"])
    ref_loss = float(d[f"loss_{tag}"])
    assert abs(loss.item() - ref_loss) <= 1e-9 * abs(ref_loss)
    # oracle agreement incl. the gradient seed
    to, qo, lo, go = do.ddqn_target_loss(d["q_next_online"], d["q_next_target"], d["q_cur"], d["actions"],
                                         d["rewards"], d["dones"], float(d["gamma"]), tag == "double")
    assert np.array_equal(grad.cpu().numpy(), go)
    # launches are deterministic (fixed-order reduction)
    loss2 = ddqn.ddqn_target_loss(t("q_next_online"), t("q_next_target"), t("q_cur"), t("actions"), t("rewards"),
                                  t("dones"), float(d["gamma"]), tag == "double")[0]
    assert loss2.item() == loss.item()


def test_config3_dense_batch_5000_gamma_095_and_080(cuda, golden_dir):
    """BASELINE.json config 3 as stated: batch 5000 out of the reference's buffer with the reference's own
    indices (K2), then the fused target/loss (K3) on the Q tensors the reference's dense network produced, for
    gamma = 0.95 and 0.80 (float32-rounded differently), Double DQN on and off."""
    d = np.load(os.path.join(golden_dir, "dqn_dense_b5000.npz"))
    ring = fill_ring(d, cuda)
    st, ac, rw, ns, dn = ring.sample(5000, idx_override=torch.from_numpy(d["idx"]).to(cuda))
    assert np.array_equal(st.cpu().numpy(), d["states"]) and np.array_equal(ns.cpu().numpy(), d["next_states"])
    assert np.array_equal(ac.cpu().numpy(), d["actions"]) and np.array_equal(rw.cpu().numpy(), d["rewards"])
    assert np.array_equal(dn.cpu().numpy(), d["dones"])
    t = lambda k: torch.from_numpy(d[k]).to(cuda)  # noqa: E731
    for gamma, gtag in ((0.95, "g095"), (0.80, "g080")):
        for use_double in (True, False):
            tag = f"{gtag}_{'double' if use_double else 'single'}"
            loss, target, q_sa, _ = ddqn.ddqn_target_loss(t("q_next_online"), t("q_next_target"), t("q_cur"), ac, rw, dn,
                                                          gamma, use_double)
            assert np.array_equal(target.cpu().numpy(), d[f"target_{tag}"])
            assert np.array_equal(q_sa.cpu().numpy(), d[f"q_sa_{tag}"])
            ref = float(d[f"loss_{tag}"])
            assert abs(loss.item() - ref) <= 1e-9 * abs(ref)


def test_ddqn_target_loss_is_reentrant_across_streams(cuda):
    """include/b2048.h promises re-entrancy: the kernel owns no global scratch (one cluster, partial sums
    through distributed shared memory), so many launches in flight on several streams all return the
    bit-identical loss of a lone launch on the same data."""
    B = 5000
    g = torch.Generator(device="cpu").manual_seed(5)
    sets = []
    for k in range(3):
        q = [torch.randn(B, 4, dtype=torch.float64, generator=g).to(cuda) for _ in range(3)]
        a = torch.randint(0, 4, (B,), generator=g).to(cuda)
        r = torch.randint(0, 64, (B,), generator=g).to(cuda)
        dn = torch.randint(0, 2, (B,), generator=g).to(cuda)
        sets.append((q, a, r, dn))
    alone = []
    for q, a, r, dn in sets:
        loss, tgt, _, _ = ddqn.ddqn_target_loss(q[0], q[1], q[2], a, r, dn, 0.95, True)
        torch.cuda.synchronize()
        alone.append((loss.clone(), tgt.clone()))
    streams = [torch.cuda.Stream(device=cuda) for _ in range(3)]
    got = [[] for _ in sets]
    torch.cuda.synchronize()
    for rep in range(40):                      # 120 launches queued back to back on three streams
        for k, (q, a, r, dn) in enumerate(sets):
            with torch.cuda.stream(streams[k]):
                loss, tgt, _, _ = ddqn.ddqn_target_loss(q[0], q[1], q[2], a, r, dn, 0.95, True)
                got[k].append((loss, tgt))
    torch.cuda.synchronize()
    for k in range(3):
        for loss, tgt in got[k]:
            assert torch.equal(loss, alone[k][0]) and torch.equal(tgt, alone[k][1])
    # and against the numpy oracle
    q, a, r, dn = sets[0]
    t_ref, _, l_ref, _ = do.ddqn_target_loss(q[0].cpu().numpy(), q[1].cpu().numpy(), q[2].cpu().numpy(), a.cpu().numpy(),
                                             r.cpu().numpy(), dn.cpu().numpy(), 0.95, True)
    assert np.array_equal(alone[0][1].cpu().numpy(), t_ref)
    np.testing.assert_allclose(float(alone[0][0].item()), l_ref, rtol=1e-12)


@pytest.mark.parametrize("B", [1, 7, 511, 512, 4096, 4097, 70001])
def test_ddqn_target_loss_sizes(cuda, B):
    """Ragged batch sizes around the cluster's 8 x 512 threads, non-double branch included."""
    g = torch.Generator(device="cpu").manual_seed(B)
    q = [torch.randn(B, 4, dtype=torch.float64, generator=g).to(cuda) for _ in range(3)]
    a = torch.randint(0, 4, (B,), generator=g).to(cuda)
    r = torch.randint(0, 2048, (B,), generator=g).to(cuda)
    dn = torch.randint(0, 2, (B,), generator=g).to(cuda)
    for use_double in (True, False):
        loss, tgt, qsa, grad = ddqn.ddqn_target_loss(q[0], q[1], q[2], a, r, dn, 0.8, use_double)
        t_ref, q_ref, l_ref, g_ref = do.ddqn_target_loss(*[t.cpu().numpy() for t in q], a.cpu().numpy(), r.cpu().numpy(),
                                                         dn.cpu().numpy(), 0.8, use_double)
        assert np.array_equal(tgt.cpu().numpy(), t_ref) and np.array_equal(qsa.cpu().numpy(), q_ref)
        assert np.array_equal(grad.cpu().numpy(), g_ref)
        np.testing.assert_allclose(float(loss.item()), l_ref, rtol=1e-12)


def test_ddqn_loss_autograd_matches_torch(cuda):
    """The autograd wrapper gives the online network the same gradient as the reference's
    expression written in torch (one_hot mask, sum, MSELoss(sum))."""
    torch.manual_seed(0)
    B = 5000
    lin = torch.nn.Linear(16, 4).double().to(cuda)
    lin2 = torch.nn.Linear(16, 4).double().to(cuda)
    x = torch.randn(B, 16, dtype=torch.float64, device=cuda)
    a = torch.randint(0, 4, (B,), device=cuda)
    r = torch.randint(0, 64, (B,), device=cuda) * 4
    dn = (torch.rand(B, device=cuda) < 0.1).long()
    # reference expression (src/dqn_lib.py:126-158)
    nq = lin(x)
    best = torch.argmax(nq, axis=1)
    mask = torch.zeros(B, 4, device=cuda); mask[torch.arange(B), best] = 1
    nb = torch.sum(lin2(x) * mask, axis=1)
    tgt = (r + (1 - dn) * 0.8 * nb).double()
    m2 = torch.zeros(B, 4, device=cuda); m2[torch.arange(B), a] = 1
    q = torch.sum(lin(x) * m2, axis=1, keepdim=True, dtype=torch.double)
    q = torch.transpose(q, 0, 1)[0]
    ref = torch.nn.MSELoss(reduction="sum")(q, tgt.detach())
    ref.backward()
    g_ref = [p.grad.clone() for p in lin.parameters()]
    for p in lin.parameters():
        p.grad = None
    loss, target, q_sa = ddqn.ddqn_loss(lin(x), lin(x), lin2(x), a, r, dn, 0.8, True)
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 1e-9 * abs(ref.item())
    np.testing.assert_allclose(target.cpu().numpy(), tgt.detach().cpu().numpy(), rtol=1e-12)
    for p, gr in zip(lin.parameters(), g_ref):
        np.testing.assert_allclose(p.grad.cpu().numpy(), gr.cpu().numpy(), rtol=1e-9, atol=1e-9)


def test_egreedy_matches_reference(cuda, golden_dir):
    """Greedy branch incl. the Q - min*max - min quirk, zeroed illegal entries and first-index
    ties (src/dqn_lib.py:23-30), on the reference's recorded decisions; random branch via the hook."""
    e = np.load(os.path.join(golden_dir, "egreedy.npz"))
    n = len(e["q"])
    q = torch.from_numpy(e["q"]).to(cuda)
    boards = dev_boards(e["state"], cuda)
    flags = env.legal_mask(boards)
    assert np.array_equal(flags.cpu().numpy() & 0xF, e["legal"])
    greedy = torch.full((n,), 0x80, dtype=torch.uint8, device=cuda)
    act, mq = ddqn.egreedy_select(q, flags, 0.0, override=greedy)
    assert np.array_equal(act.cpu().numpy(), e["action"])
    assert np.array_equal(mq.cpu().numpy(), e["max_q"])
    assert np.array_equal((flags.cpu().numpy() & 0x10) != 0, e["done"].astype(bool))
    # epsilon = 0 without override == all greedy; epsilon = 1 == all random with max_q = 0
    act0, mq0 = ddqn.egreedy_select(q, flags, 0.0, seed=1)
    assert torch.equal(act0, act) and torch.equal(mq0, mq)
    act1, mq1 = ddqn.egreedy_select(q, flags, 1.0, seed=1)
    assert (mq1 == 0).all() and set(torch.unique(act1).tolist()) == {0, 1, 2, 3}
    forced = torch.randint(0, 4, (n,), dtype=torch.uint8, device=cuda)
    actf, mqf = ddqn.egreedy_select(q, flags, 0.5, override=forced)
    assert torch.equal(actf, forced) and (mqf == 0).all()
    # mixed epsilon: explore fraction ~ eps, greedy rows equal the greedy answer
    big = 200000
    qb = q.repeat((big + n - 1) // n, 1)[:big].contiguous()
    fb = flags.repeat((big + n - 1) // n)[:big].contiguous()
    actm, mqm = ddqn.egreedy_select(qb, fb, 0.3, seed=9, ctr=4)
    explored = (mqm == 0) & (qb.max(dim=1).values != 0)
    assert 0.29 < explored.float().mean().item() < 0.31
    ao, mo = do.egreedy_batch(qb.cpu().numpy()[:2000], fb.cpu().numpy()[:2000], np.full(2000, 0x80, np.uint8))
    keep = (~explored & (qb.max(dim=1).values != 0))[:2000].cpu().numpy()
    assert np.array_equal(actm.cpu().numpy()[:2000][keep], ao[keep])


# ---- K6: fused conv Q-network forward -----------------------------------------------------------------
def _conv_net(d, prefix, cuda):
    from test_shim_gpu import conv_model
    net = conv_model()
    net.load_state_dict({k[len(prefix):]: torch.from_numpy(d[k]) for k in d.files if k.startswith(prefix)})
    return net.to(cuda)


def test_fused_conv_forward_matches_reference_q_values(cuda, golden_dir):
    """K6 against the Q tensors the reference's train_step computed with its own weights on its own
    sampled batch (tests/golden/dqn_conv.npz): states -> Q(s), next_states -> Q(s') for both networks.
    Tolerance (north_star): 1e-9 RELATIVE, element by element, with an absolute floor of 1e-12 for the
    few Q-values that happen to be ~0 (float64, different summation order; measured ~1e-15)."""
    d = np.load(os.path.join(golden_dir, "dqn_conv.npz"))
    online, target = _conv_net(d, "w_", cuda), _conv_net(d, "tw_", cuda)
    fo, ft = b2048.qfused.FusedConvQ(online), b2048.qfused.FusedConvQ(target)
    s, s2 = torch.from_numpy(d["states"]).to(cuda), torch.from_numpy(d["next_states"]).to(cuda)
    for got, want in ((fo(s), d["q_cur"]), (fo(s2), d["q_next_online"]), (ft(s2), d["q_next_target"])):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-9, atol=1e-12)
    # packed boards in, exponent scaling == the float64 states the reference builds with log_scale()
    packed = dev_boards((2 ** d["next_states"].astype(np.int64)) * (d["next_states"] > 0), cuda)
    np.testing.assert_allclose(fo.forward_boards(packed).cpu().numpy(), d["q_next_online"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("n", [1, 3, 31, 32, 33, 257, 2368, 4736, 5000, 5919, 70001])
def test_fused_conv_forward_sizes_and_scalings(cuda, n):
    """Ragged sizes (partial tiles, partial warp pairs, more tiles than SMs) and both input scalings
    against the torch float64 module on the same inputs."""
    from test_shim_gpu import conv_model
    torch.manual_seed(n)
    net = conv_model().to(cuda)
    fq = b2048.qfused.FusedConvQ(net)
    boards = env.random_boards(n, seed=77, p_empty=0.3, max_exp=15, device=cuda)
    x_log = env.unpack_f64(boards, conv=True)
    tiles = env.unpack_tiles(boards).to(torch.float64)
    x_norm = (tiles / tiles.max(dim=1, keepdim=True).values).view(n, 1, 4, 4)
    with torch.no_grad():
        for scaling, x in (("log2", x_log), ("normalized", x_norm)):
            want = net(x)
            got = fq.forward_boards(boards, scaling=scaling)
            tol = 1e-9 * want.abs() + 1e-12                   # element-wise relative + absolute floor
            assert bool(((got - want).abs() <= tol).all()), (scaling, float((got - want).abs().max()))
            assert bool(((fq(x.contiguous()) - want).abs() <= tol).all())
    if n <= 5000:                                         # and against the numpy oracle (pinned to the reference's Q tensors)
        want = do.conv_q_forward(x_log.reshape(n, 16).cpu().numpy(), *[p.detach().cpu().numpy() for p in net.parameters()])
        np.testing.assert_allclose(fq.forward_boards(boards).cpu().numpy(), want, rtol=1e-9, atol=1e-12)
    # the kernel reads the module's parameters at call time: an in-place weight change is seen
    with torch.no_grad():
        net[7].bias.add_(1.0)
        assert float((fq.forward_boards(boards) - net(x_log)).abs().max()) <= 1e-9 * float(net(x_log).abs().max())


def test_fused_conv_forward_rejects_other_networks(cuda):
    from test_shim_gpu import conv_model, dense_model
    assert not b2048.qfused.matches(dense_model().to(cuda))
    assert not b2048.qfused.matches(conv_model())                      # CPU module
    assert not b2048.qfused.matches(conv_model().float().to(cuda))     # float32
    with pytest.raises(ValueError):
        b2048.qfused.FusedConvQ(dense_model().to(cuda))
    assert isinstance(b2048.qfused.accelerate_inference(conv_model().to(cuda)), b2048.qfused.FusedConvQ)
    fq = b2048.qfused.FusedConvQ(conv_model().to(cuda))
    with pytest.raises(ValueError):
        fq.forward_boards(env.random_boards(8, device=cuda), scaling="sqrt")
    assert fq.forward_boards(torch.empty(0, dtype=torch.int64, device=cuda)).shape == (0, 4)
