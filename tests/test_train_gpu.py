"""GPU tests of the batched rollout (VectorEnv) and the graph-captured Double-DQN update."""
import copy
import os

import numpy as np
import pytest
import torch

import b2048
from b2048 import env
from b2048.rollout import VectorEnv
from b2048.trainer import DDQNUpdater
from oracle import board_oracle as bo
from test_shim_gpu import conv_model, dense_model

pytestmark = pytest.mark.gpu


def test_vector_env_random_policy_and_replay_consistency(cuda):
    n = 8192
    ve = VectorEnv(n, device=cuda, seed=11, p_four=0.5)
    ring = b2048.ReplayRing(200000, device=cuda)
    for _ in range(400):
        ve.step(replay=ring)
    st = ve.stats()
    # the reference's random-policy games: ~100 transitions incl. illegal no-ops, max tile mostly 64/128
    assert st["games"] > 15000 and 80 < st["mean_moves"] < 200, st
    top = max(st["max_tile_hist"], key=st["max_tile_hist"].get)
    assert top in (64, 128), st
    assert len(ring) == 200000
    # every stored transition is a valid 2048 move: oracle slide + one spawned tile iff changed;
    # done transitions are dead->dead no-ops
    idx = torch.randint(0, 200000, (4000,), device=cuda)
    s, a, r, s2, d, _ = ring.sample(4000, idx_override=idx, return_idx=True)
    S = s.cpu().numpy().astype(np.int64)
    S2 = s2.cpu().numpy().astype(np.int64)
    tiles = np.where(S > 0, 1 << S, 0)
    tiles2 = np.where(S2 > 0, 1 << S2, 0)
    A, R, D = a.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
    for i in range(0, 4000, 5):
        slid, rew, changed = bo.slide(tiles[i], int(A[i]))
        diff = tiles2[i].reshape(4, 4) - slid
        assert rew == R[i]
        assert (diff != 0).sum() == (1 if changed else 0)
        assert D[i] == (bo.legal_mask(tiles[i]) == 0)
        if D[i]:
            assert not changed
    assert 0 < D.mean() < 0.05


def test_vector_env_greedy_policy_uses_the_network(cuda):
    torch.manual_seed(0)
    model = conv_model().to(cuda)
    ve = VectorEnv(2048, device=cuda, seed=3, conv=True)
    for _ in range(50):
        ve.step(model=model, epsilon=0.1)
    assert ve.stats()["steps"] == 50 * 2048 and float(ve.ep_qsum.abs().sum()) > 0


def _filled_ring(cuda, n=15000):
    ve = VectorEnv(4096, device=cuda, seed=5)
    ring = b2048.ReplayRing(n, device=cuda)
    for _ in range(8):
        ve.step(replay=ring)
    return ring


@pytest.mark.parametrize("kind", ["conv", "dense"])
def test_updater_matches_plain_torch_update(cuda, kind):
    """One real update (zero_grad -> backward -> Adam) equals the same update written with plain
    torch ops on the same sampled batch: loss within 1e-9 relative, weights within 1e-9."""
    torch.manual_seed(1)
    ring = _filled_ring(cuda)
    model = (conv_model() if kind == "conv" else dense_model()).to(cuda)
    target = copy.deepcopy(model)
    with torch.no_grad():
        for p in target.parameters():
            p.add_(0.01 * torch.randn_like(p))
    ref_model = copy.deepcopy(model)
    up = DDQNUpdater(model, ring, batch_size=5000, gamma=0.8, lr=1e-2, conv=kind == "conv",
                     target_model=copy.deepcopy(target), use_graph=False, seed=77)
    ctr = int(ring.head_size[2].item())
    st, ac, rw, ns, dn = ring.sample(5000, seed=77, ctr=ctr)           # the batch update() will draw
    loss = up.update().item()
    shape = (5000, 1, 4, 4) if kind == "conv" else (5000, 16)
    opt = torch.optim.Adam(ref_model.parameters(), lr=1e-2)
    nq = ref_model(ns.view(shape))
    best = torch.argmax(nq, dim=1)
    nb = target(ns.view(shape)).gather(1, best[:, None])[:, 0]
    tgt = (rw + (1 - dn) * 0.8 * nb).double().detach()                # float32 gamma, like the reference
    q = ref_model(st.view(shape)).gather(1, ac[:, None])[:, 0]
    ref_loss = torch.nn.MSELoss(reduction="sum")(q, tgt)
    opt.zero_grad()
    ref_loss.backward()
    opt.step()
    assert abs(loss - ref_loss.item()) <= 1e-9 * abs(ref_loss.item())
    g_ref = torch.cat([p.grad.flatten() for p in ref_model.parameters()]).cpu().numpy()
    g = up.grads.flat.cpu().numpy()
    np.testing.assert_allclose(g, g_ref, rtol=1e-9, atol=1e-9 * np.abs(g_ref).max())
    # Adam's first step is ~lr*sign(g): elements with |g| ~ eps amplify rounding noise, hence atol
    for p, rp in zip(model.parameters(), ref_model.parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), rp.detach().cpu().numpy(), rtol=1e-9, atol=1e-6)
    assert int(ring.head_size[2].item()) == ctr + 1


def test_updater_graph_equals_eager_and_learns(cuda):
    torch.manual_seed(2)
    ring = _filled_ring(cuda)
    base = conv_model().to(cuda)
    a = DDQNUpdater(copy.deepcopy(base), ring, batch_size=5000, lr=1e-3, use_graph=False, seed=9)
    ring.head_size[2] = 0
    losses_a = [a.update().item() for _ in range(8)]
    # graph mode: the warm-up + capture inside the first update() must leave no trace -- same
    # losses, one optimizer step and one sample-counter bump per update(), from the very first call
    b = DDQNUpdater(copy.deepcopy(base), ring, batch_size=5000, lr=1e-3, use_graph=True, seed=9)
    ring.head_size[2] = 0
    losses_b = [b.update().item() for _ in range(8)]
    np.testing.assert_allclose(losses_a, losses_b, rtol=1e-12)
    assert int(b.opt.step_count.item()) == 8 and int(ring.head_size[2].item()) == 8 and b.updates == 8
    for p, q in zip(a.model.parameters(), b.model.parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().cpu().numpy(), rtol=1e-12, atol=1e-14)
    # the graph path itself: runs, advances the sample counter, changes weights, reduces the loss
    c = DDQNUpdater(copy.deepcopy(base), ring, batch_size=5000, lr=1e-3, use_graph=True, seed=9)
    first = c.update().item()
    ctr0 = int(ring.head_size[2].item())
    for _ in range(60):
        c.update()
    assert int(ring.head_size[2].item()) == ctr0 + 60
    last = c.loss.item()
    assert last < first
    assert not any(torch.equal(p, q) for p, q in zip(c.model.parameters(), base.parameters()))
    c.sync_target()
    assert all(torch.equal(p, q) for p, q in zip(c.model.parameters(), c.target.parameters()))


@pytest.mark.parametrize("rows,n_in,n_out", [(5000, 512, 512), (5000, 512, 256), (5000, 16, 512), (5000, 256, 4),
                                             (1, 512, 256), (7, 16, 512), (137, 64, 128), (4097, 48, 64), (300, 130, 6)])
def test_dense_tensor_core_layers_match_torch(cuda, rows, n_in, n_out):
    """K8 (dense_kernels.cu): nn.Linear forward (+ReLU), input gradient (+ReLU mask) and weight / bias gradient on
    the FP64 tensor cores against torch float64, element-wise 1e-9 relative (+1e-12 absolute: sums of O(100)
    terms of mixed sign; measured agreement ~1e-14), incl. ragged row counts and widths that fill only part of a tile."""
    from b2048 import _lib
    from b2048.env import _ptr, _stream
    g = torch.Generator(device="cpu").manual_seed(rows * 31 + n_in)
    x = torch.randn(rows, n_in, dtype=torch.float64, generator=g).to(cuda)
    w = (torch.randn(n_out, n_in, dtype=torch.float64, generator=g) * 0.1).to(cuda)
    b = torch.randn(n_out, dtype=torch.float64, generator=g).to(cuda)
    gy = torch.randn(rows, n_out, dtype=torch.float64, generator=g).to(cuda)
    h = torch.randn(rows, n_in, dtype=torch.float64, generator=g).to(cuda)     # "output of the layer below" for the mask
    L, st = _lib.lib(), _stream(x)

    def close(got, want):
        tol = 1e-9 * want.abs() + 1e-12 * max(1.0, float(want.abs().max()))
        assert bool(((got - want).abs() <= tol).all()), float((got - want).abs().max())

    for relu in ((0,) if n_out == 4 else (0, 1)):
        out = torch.full((rows, n_out), float("nan"), dtype=torch.float64, device=cuda)
        _lib.check(L.dense_linear_forward_f64(_ptr(x), _ptr(w), _ptr(b), _ptr(out), rows, n_in, n_out, relu, st))
        want = torch.addmm(b, x, w.t())
        close(out, torch.relu(want) if relu else want)
    dz = torch.full((rows, n_in), float("nan"), dtype=torch.float64, device=cuda)
    _lib.check(L.dense_linear_dgrad_f64(_ptr(gy), _ptr(w), _ptr(h), _ptr(dz), rows, n_in, n_out, st))
    close(dz, (gy @ w) * (h > 0))
    scratch = torch.empty(int(L.dense_linear_wgrad_scratch_elems(rows, n_in, n_out)), dtype=torch.float64, device=cuda)
    dw = torch.full((n_out, n_in), float("nan"), dtype=torch.float64, device=cuda)
    db = torch.full((n_out,), float("nan"), dtype=torch.float64, device=cuda)
    _lib.check(L.dense_linear_wgrad_f64(_ptr(gy), _ptr(x), _ptr(dw), _ptr(db), _ptr(scratch), rows, n_in, n_out, st))
    close(dw, gy.t() @ x)
    close(db, gy.sum(dim=0))
    dw2, db2 = torch.empty_like(dw), torch.empty_like(db)                          # bit-reproducible
    _lib.check(L.dense_linear_wgrad_f64(_ptr(gy), _ptr(x), _ptr(dw2), _ptr(db2), _ptr(scratch), rows, n_in, n_out, st))
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


@pytest.mark.parametrize("n", [1, 33, 5000])
def test_dense_q_network_forward_and_backward_match_autograd(cuda, n):
    """DenseQ / TrainableDenseQ on the reference's dense Q-network against the nn.Sequential itself."""
    from b2048.qdense import DenseQ, TrainableDenseQ, matches
    torch.manual_seed(n)
    net = dense_model().to(cuda)
    assert matches(net) and not matches(conv_model().to(cuda))
    x = torch.randn(n, 16, dtype=torch.float64, device=cuda) * 3
    want = net(x)
    got = DenseQ(net)(x)
    assert bool(((got - want).abs() <= 1e-9 * want.abs() + 1e-12).all())
    tq = TrainableDenseQ(net)
    gq = torch.randn(n, 4, dtype=torch.float64, device=cuda)
    want.backward(gq)
    ref = [p.grad.clone() for p in net.parameters()]
    q, saved = tq.forward_saving(x)
    assert torch.equal(q, got)
    grads = [torch.full_like(p, float("nan")) for p in net.parameters()]
    tq.backward_into(saved, gq, grads)
    for a, b in zip(grads, ref):
        assert bool(((a - b).abs() <= 1e-9 * b.abs() + 1e-12 * max(1.0, float(b.abs().max()))).all()), float((a - b).abs().max())
    # the autograd entry point (dqn_lib.train_step with the caller's loss and optimizer)
    for p in net.parameters():
        p.grad = None
    (tq(x) * gq).sum().backward()
    for p, b in zip(net.parameters(), ref):
        assert bool(((p.grad - b).abs() <= 1e-9 * b.abs() + 1e-12 * max(1.0, float(b.abs().max()))).all())


def _other_conv_model():
    from torch import nn   # not the reference's widths: exercises the generic FastQNet path
    return nn.Sequential(nn.Conv2d(1, 32, kernel_size=2), nn.ReLU(), nn.Conv2d(32, 64, kernel_size=2), nn.ReLU(),
                         nn.Flatten(), nn.Linear(2 * 2 * 64, 48), nn.ReLU(), nn.Linear(48, 4)).double()


@pytest.mark.parametrize("which,n", [("reference", 5000), ("reference", 33), ("reference", 6000), ("other", 5000), ("other", 7)])
def test_fast_conv_forward_equals_cudnn_path(cuda, which, n):
    """The conv Q-network with gradients — reference architecture: K6 forward with saved activations +
    hand-built backward (K7, masked col2im, cuBLAS); any other small conv net: FastQNet (patch gathers +
    DGEMMs + K7) — gives the same outputs and parameter gradients as the plain torch float64 module
    (1e-12 / 1e-10 relative), bit-identically from run to run."""
    from b2048.qfused import TrainableConvQ
    from b2048.qnet import FastQNet, accelerate
    torch.manual_seed(3)
    net = (conv_model() if which == "reference" else _other_conv_model()).to(cuda)
    fast = accelerate(net)
    assert isinstance(fast, TrainableConvQ if which == "reference" else FastQNet)
    assert accelerate(dense_model()).__class__.__name__ == "Sequential"
    x = torch.randint(0, 12, (n, 1, 4, 4), device=cuda).double()
    a, b = net(x), fast(x)
    np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().cpu().numpy(), rtol=1e-12, atol=1e-13)
    a.square().sum().backward()
    g1 = [p.grad.clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    fast(x).square().sum().backward()
    g2 = [p.grad.clone() for p in net.parameters()]
    for u, v in zip(g1, g2):
        np.testing.assert_allclose(v.cpu().numpy(), u.cpu().numpy(), rtol=1e-10, atol=1e-10 * float(u.abs().max()))
    for p in net.parameters():
        p.grad = None
    fast(x).square().sum().backward()
    assert all(torch.equal(p.grad, v) for p, v in zip(net.parameters(), g2))       # fixed summation order
    with torch.no_grad():
        np.testing.assert_allclose(fast(x).cpu().numpy(), a.detach().cpu().numpy(), rtol=1e-12, atol=1e-13)
    xg = x.clone().requires_grad_(True)                   # input gradient requested: generic autograd path
    fast(xg).square().sum().backward()
    xr = x.clone().requires_grad_(True)
    net(xr).square().sum().backward()
    np.testing.assert_allclose(xg.grad.cpu().numpy(), xr.grad.cpu().numpy(), rtol=1e-10, atol=1e-10 * float(xr.grad.abs().max()))


@pytest.mark.parametrize("rows,c,k", [(45000, 64, 4), (5000, 4, 64), (1, 2, 3), (127, 16, 64), (128, 64, 16), (40001, 1, 1)])
def test_small_weight_gradient_kernel(cuda, rows, c, k):
    """layer_wgrad_small_f64: dW = g^T x and db = column sums of g against torch (1e-12 relative to the
    largest entry), bit-identical across runs, and argument checks."""
    from b2048 import _lib
    from b2048.env import _ptr, _stream
    _lib.init(torch.device(cuda).index or 0)
    L = _lib.lib()
    torch.manual_seed(rows)
    g = torch.randn(rows, c, dtype=torch.float64, device=cuda)
    x = torch.randn(rows, k, dtype=torch.float64, device=cuda)
    scratch = torch.empty(L.layer_wgrad_small_scratch_elems(rows, c, k), dtype=torch.float64, device=cuda)

    def run():
        dw = torch.full((c, k), float("nan"), dtype=torch.float64, device=cuda)
        db = torch.full((c,), float("nan"), dtype=torch.float64, device=cuda)
        with torch.cuda.device(cuda):
            assert L.layer_wgrad_small_f64(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), _ptr(scratch), rows, c, k, _stream(g)) == 0
        return dw, db

    dw, db = run()
    dw2, db2 = run()
    want_w, want_b = g.t() @ x, g.sum(0)
    assert float((dw - want_w).abs().max()) <= 1e-12 * max(1.0, float(want_w.abs().max()))
    assert float((db - want_b).abs().max()) <= 1e-12 * max(1.0, float(want_b.abs().max()))
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    with torch.cuda.device(cuda):
        assert L.layer_wgrad_small_f64(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), _ptr(scratch), rows, 65, k, _stream(g)) == -2
        assert L.layer_wgrad_small_f64(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), None, rows, c, k, _stream(g)) == -2


@pytest.mark.parametrize("rows,k", [(20000, 256), (5000, 256), (1, 32), (33, 64), (4737, 128), (63, 256)])
def test_dmma_weight_gradient_kernel(cuda, rows, k):
    """layer_wgrad64_f64 (64 x K weight gradient on the FP64 tensor cores) against torch, run to run
    bit-identical, ragged row counts (partial tiles, fewer row tiles than SMs)."""
    from b2048 import _lib
    from b2048.env import _ptr, _stream
    _lib.init(torch.device(cuda).index or 0)
    L = _lib.lib()
    torch.manual_seed(rows + k)
    g = torch.randn(rows, 64, dtype=torch.float64, device=cuda)
    x = torch.randn(rows, k, dtype=torch.float64, device=cuda)
    scratch = torch.empty(L.layer_wgrad64_scratch_elems(rows, k), dtype=torch.float64, device=cuda)

    def run():
        dw = torch.full((64, k), float("nan"), dtype=torch.float64, device=cuda)
        db = torch.full((64,), float("nan"), dtype=torch.float64, device=cuda)
        with torch.cuda.device(cuda):
            assert L.layer_wgrad64_f64(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), _ptr(scratch), rows, k, _stream(g)) == 0
        return dw, db

    dw, db = run()
    dw2, db2 = run()
    want_w, want_b = g.t() @ x, g.sum(0)
    assert float((dw - want_w).abs().max()) <= 1e-12 * max(1.0, float(want_w.abs().max()))
    assert float((db - want_b).abs().max()) <= 1e-12 * max(1.0, float(want_b.abs().max()))
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    with torch.cuda.device(cuda):
        assert L.layer_wgrad64_f64(_ptr(g), _ptr(x), _ptr(dw), _ptr(db), _ptr(scratch), rows, 48, _stream(g)) == -2


@pytest.mark.parametrize("n", [1, 5, 16, 17, 5000, 6001])
def test_fused_conv1_backward_kernel(cuda, n):
    """conv1_wgrad_fused_f64 = col2im + ReLU mask + (dW1, db1) in one pass, against the same thing spelled
    out with torch ops on the unfused tensors; run-to-run bit-identical."""
    from b2048 import _lib
    from b2048.env import _ptr, _stream
    _lib.init(torch.device(cuda).index or 0)
    L = _lib.lib()
    torch.manual_seed(n)
    x = torch.randint(0, 12, (n, 16), device=cuda).double()
    w1 = torch.randn(64, 4, dtype=torch.float64, device=cuda)
    b1 = torch.randn(64, dtype=torch.float64, device=cuda)
    # forward pieces: conv1 patches [9n,4] -> out1 [n,3,3,64] -> patches2 [4n,256] in (c, ky, kx) column order
    xi = x.view(n, 4, 4)
    p1 = torch.stack([xi[:, y + ky, xx + kx] for y in range(3) for xx in range(3) for ky in range(2) for kx in range(2)], 1).view(n * 9, 4)
    out1 = torch.relu(p1 @ w1.t() + b1).view(n, 3, 3, 64)
    p2 = torch.stack([out1[:, oy + ky, ox + kx, :] for oy in range(2) for ox in range(2) for ky in range(2) for kx in range(2)], 1)
    p2 = p2.view(n, 4, 4, 64).permute(0, 1, 3, 2).reshape(4 * n, 256).contiguous()        # [.., c*4 + tap]
    gp2 = torch.randn(4 * n, 256, dtype=torch.float64, device=cuda)
    # reference: col2im of gp2, mask by out1 > 0, then dW1 = g1^T p1, db1 = column sums
    g1 = torch.zeros(n, 3, 3, 64, dtype=torch.float64, device=cuda)
    gv = gp2.view(n, 2, 2, 64, 2, 2)
    for oy in range(2):
        for ox in range(2):
            for ky in range(2):
                for kx in range(2):
                    g1[:, oy + ky, ox + kx, :] += gv[:, oy, ox, :, ky, kx]
    g1 = (g1 * (out1 > 0)).view(9 * n, 64)
    want_w, want_b = g1.t() @ p1, g1.sum(0)
    scratch = torch.empty(L.conv1_wgrad_fused_scratch_elems(n), dtype=torch.float64, device=cuda)

    def run():
        dw = torch.full((64, 4), float("nan"), dtype=torch.float64, device=cuda)
        db = torch.full((64,), float("nan"), dtype=torch.float64, device=cuda)
        with torch.cuda.device(cuda):
            assert L.conv1_wgrad_fused_f64(_ptr(gp2), _ptr(p2), _ptr(x), _ptr(dw), _ptr(db), _ptr(scratch), n, _stream(x)) == 0
        return dw, db

    dw, db = run()
    dw2, db2 = run()
    assert float((dw - want_w).abs().max()) <= 1e-11 * max(1.0, float(want_w.abs().max()))
    assert float((db - want_b).abs().max()) <= 1e-11 * max(1.0, float(want_b.abs().max()))
    assert torch.equal(dw, dw2) and torch.equal(db, db2)


@pytest.mark.parametrize("n", [1, 2, 5, 17, 300, 5000, 6001, 11000])
def test_conv2_dgrad_conv1_backward_in_the_accumulators(cuda, n):
    """conv2_dgrad_conv1_wgrad_f64 (patch-matrix gradient consumed inside the tensor-core accumulators) against
    the unfused pair it replaces — dense_linear_dgrad_f64 (plain g2 W2) then conv1_wgrad_fused_f64 — and against
    torch; run-to-run bit-identical.  11000 boards = more than one round of chunks per CTA."""
    from b2048 import _lib
    from b2048.env import _ptr, _stream
    _lib.init(torch.device(cuda).index or 0)
    L = _lib.lib()
    torch.manual_seed(n)
    kw = dict(dtype=torch.float64, device=cuda)
    x = torch.randint(0, 12, (n, 16), device=cuda).double()
    w1, b1 = torch.randn(64, 4, **kw), torch.randn(64, **kw)
    w2 = torch.randn(64, 64, 2, 2, **kw)
    xi = x.view(n, 4, 4)
    p1 = torch.stack([xi[:, y + ky, xx + kx] for y in range(3) for xx in range(3) for ky in range(2) for kx in range(2)], 1).view(n * 9, 4)
    out1 = torch.relu(p1 @ w1.t() + b1).view(n, 3, 3, 64)
    p2 = torch.stack([out1[:, oy + ky, ox + kx, :] for oy in range(2) for ox in range(2) for ky in range(2) for kx in range(2)], 1)
    p2 = p2.view(n, 4, 4, 64).permute(0, 1, 3, 2).reshape(4 * n, 256).contiguous()
    g2 = torch.randn(4 * n, 64, **kw)
    gp2 = g2 @ w2.view(64, 256)
    g1 = torch.zeros(n, 3, 3, 64, **kw)
    gv = gp2.view(n, 2, 2, 64, 2, 2)
    for oy in range(2):
        for ox in range(2):
            for ky in range(2):
                for kx in range(2):
                    g1[:, oy + ky, ox + kx, :] += gv[:, oy, ox, :, ky, kx]
    g1 = (g1 * (out1 > 0)).view(9 * n, 64)
    want_w, want_b = g1.t() @ p1, g1.sum(0)
    scratch = torch.empty(L.conv2_dgrad_conv1_wgrad_scratch_elems(n), **kw)
    assert scratch.numel() > 0

    def run():
        dw, db = torch.full((64, 4), float("nan"), **kw), torch.full((64,), float("nan"), **kw)
        with torch.cuda.device(cuda):
            assert L.conv2_dgrad_conv1_wgrad_f64(_ptr(g2), _ptr(w2), _ptr(p2), _ptr(x), _ptr(dw), _ptr(db), _ptr(scratch), n,
                                                 _stream(x)) == 0
        return dw, db

    dw, db = run()
    dw2, db2 = run()
    tol_w = 1e-11 * max(1.0, float(want_w.abs().max()))
    tol_b = 1e-11 * max(1.0, float(want_b.abs().max()))
    assert float((dw - want_w).abs().max()) <= tol_w and float((db - want_b).abs().max()) <= tol_b
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    # the unfused kernels it replaces
    gp2k = torch.empty(4 * n, 256, **kw)
    s1 = torch.empty(L.conv1_wgrad_fused_scratch_elems(n), **kw)
    dwu, dbu = torch.empty(64, 4, **kw), torch.empty(64, **kw)
    with torch.cuda.device(cuda):
        assert L.dense_linear_dgrad_f64(_ptr(g2), _ptr(w2), None, _ptr(gp2k), 4 * n, 256, 64, _stream(x)) == 0
        assert L.conv1_wgrad_fused_f64(_ptr(gp2k), _ptr(p2), _ptr(x), _ptr(dwu), _ptr(dbu), _ptr(s1), n, _stream(x)) == 0
    assert float((dw - dwu).abs().max()) <= tol_w and float((db - dbu).abs().max()) <= tol_b
    with torch.cuda.device(cuda):
        assert L.conv2_dgrad_conv1_wgrad_f64(_ptr(g2), _ptr(w2), _ptr(p2), _ptr(x), _ptr(dw), _ptr(db), _ptr(scratch), 0,
                                             _stream(x)) == -2


@pytest.mark.parametrize("n,double", [(1, True), (7, True), (8, False), (333, True), (5000, True), (5000, False), (20011, True)])
def test_update_forwards_in_one_launch_are_bit_identical(cuda, n, double):
    """qnet_conv_forward_update_f64 — Q(s) with saved activations, Q_online(s') and Q_target(s') as ONE K6
    launch over two weight sets — gives exactly the tensors of the three separate launches
    (qnet_conv_forward_train_f64 + 2 x qnet_conv_forward_f64), for ragged sizes and without the online Q(s')."""
    import copy
    from b2048 import qfused
    from bench import conv_qnet
    torch.manual_seed(n)
    net = conv_qnet().to(cuda)
    tgt = copy.deepcopy(net)
    with torch.no_grad():
        for p in tgt.parameters():
            p.add_(0.05 * torch.randn_like(p))
    on, tg = qfused.TrainableConvQ(net), qfused.TrainableConvQ(tgt)
    x = torch.randint(0, 12, (n, 16), device=cuda).double()
    xn = torch.randint(0, 12, (n, 16), device=cuda).double()
    q, saved, qno, qnt = on.forward_update(x, xn, tg, use_double=double)
    q1, saved1 = on.forward_saving(x)
    assert torch.equal(q, q1)
    for a, b in zip(saved, saved1):
        assert torch.equal(a, b)
    assert torch.equal(qnt, qfused.FusedConvQ(tgt)(xn))
    if double:
        assert torch.equal(qno, qfused.FusedConvQ(net)(xn))
    else:
        assert qno is None
    want = tgt(xn.view(n, 1, 4, 4))
    assert float((qnt - want).abs().max()) <= 1e-12 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("rows", [1, 37, 5000])
def test_dgrad_regrouped_store_equals_dgrad_then_transpose(cuda, rows):
    """dense_linear_dgrad_regroup_f64(group=4) writes (g W) * (h > 0) of the layer behind nn.Flatten as rows
    (board, position) x channel: bit-identical to dense_linear_dgrad_f64 followed by the view/transpose/reshape
    the conv backward used to do with an ATen copy."""
    from b2048 import _lib
    from b2048.env import _ptr, _stream
    _lib.init(torch.device(cuda).index or 0)
    L = _lib.lib()
    torch.manual_seed(rows)
    kw = dict(dtype=torch.float64, device=cuda)
    g, w = torch.randn(rows, 64, **kw), torch.randn(64, 256, **kw)
    h = torch.relu(torch.randn(rows, 256, **kw))
    plain, re = torch.empty(rows, 256, **kw), torch.full((4 * rows, 64), float("nan"), **kw)
    with torch.cuda.device(cuda):
        assert L.dense_linear_dgrad_f64(_ptr(g), _ptr(w), _ptr(h), _ptr(plain), rows, 256, 64, _stream(g)) == 0
        assert L.dense_linear_dgrad_regroup_f64(_ptr(g), _ptr(w), _ptr(h), _ptr(re), rows, 256, 64, 4, _stream(g)) == 0
        assert L.dense_linear_dgrad_regroup_f64(_ptr(g), _ptr(w), _ptr(h), _ptr(re), rows, 256, 64, 3, _stream(g)) == -2
        assert L.dense_linear_dgrad_regroup_f64(_ptr(g), _ptr(w), None, _ptr(re), rows, 256, 64, 4, _stream(g)) == -2
    assert torch.equal(re, plain.view(rows, 64, 4).transpose(1, 2).reshape(4 * rows, 64))
    want = (g @ w) * (h > 0)
    assert float((plain - want).abs().max()) <= 1e-12 * max(1.0, float(want.abs().max()))


def test_batched_player_baselines(cuda):
    """The reference's player.py policies at scale: random-legal and up-left baselines."""
    from b2048.player import BatchedPlayer
    rnd = BatchedPlayer(4096, device=cuda, seed=1).random_baseline(6000)
    # random legal play: ~110-160 moves, max tile mostly 64/128/256 (reference notebook)
    assert rnd["games"] == 8192 and 65 < rnd["mean_moves"] < 130, rnd        # 2 complete games per board; reference: ~92 (BASELINE.md, 20 games)
    assert max(rnd["max_tile_hist"], key=rnd["max_tile_hist"].get) in (64, 128, 256)
    ul = BatchedPlayer(4096, device=cuda, seed=2).upleft_baseline(6000)
    assert ul["games"] == 8192 and 50 < ul["mean_moves"] < 400, ul
    assert ul["mean_merge_score"] > 300


def test_batched_player_model_policy(cuda):
    """Player.play_game(random_policy=False) at scale (src/player.py:40-64): argmax(mask * Q) on the
    board divided by its largest tile, checked per decision against numpy and run to the end."""
    from b2048.player import BatchedPlayer
    torch.manual_seed(7)
    net = conv_model().to(cuda)
    pl = BatchedPlayer(2048, device=cuda, seed=5)
    for _ in range(20):                                   # a few random moves so that boards differ
        pl.t += 1
        pl.boards = b2048.env.step(pl.boards, b2048.env.random_actions(pl.n, seed=pl.t, device=cuda), seed=5,
                                   step_index=pl.t)[0]
    got = pl._model_actions(b2048.qnet.accelerate(net), "normalized", True).cpu().numpy()
    fused = pl._model_actions(b2048.qfused.FusedConvQ(net), "normalized", True).cpu().numpy()
    tiles = b2048.env.unpack_tiles(pl.boards).cpu().numpy().astype(np.float64)
    x = torch.from_numpy(tiles / tiles.max(axis=1, keepdims=True)).view(-1, 1, 4, 4)
    with torch.no_grad():
        q = net.cpu()(x).numpy()
    legal = bo.legal_mask_packed(pl.boards.cpu().numpy().view(np.uint64))
    mask = ((legal[:, None] >> np.arange(4)) & 1).astype(np.float64)
    margin = np.sort(mask * q, axis=1)
    clear = (margin[:, -1] - margin[:, -2]) > 1e-9       # skip numerically tied decisions
    assert clear.mean() > 0.9
    assert np.array_equal(got[clear], np.argmax(mask * q, axis=1)[clear])
    assert np.array_equal(fused[clear], np.argmax(mask * q, axis=1)[clear])
    net.to(cuda)
    st = pl.model_policy(net, 2048)
    assert st["games"] == 2048 and st["mean_moves"] >= 1 and 0 <= st["stalled_games"] <= 2048, st
    with pytest.raises(ValueError):
        pl._observe("sqrt", True)


@pytest.mark.parametrize("use_graph", [False, True])
def test_checkpoint_resume_is_bit_identical(cuda, tmp_path, use_graph):
    """Save mid-run, keep going, reload, redo: same boards, same replay contents, same weights -- also
    in CUDA-graph mode, where load() drops the graph and the re-capture must not apply extra updates."""
    from b2048 import checkpoint
    torch.manual_seed(4)

    def make():
        ve = VectorEnv(2048, device=cuda, seed=21)
        ring = b2048.ReplayRing(15000, device=cuda)
        up = DDQNUpdater(dense_model().to(cuda), ring, batch_size=512, lr=1e-3, conv=False, use_graph=use_graph, seed=3)
        return ve, ring, up

    def run(ve, ring, up, steps):
        for _ in range(steps):
            ve.step(replay=ring)
            up.update()

    torch.manual_seed(4)
    ve, ring, up = make()
    run(ve, ring, up, 6)
    path = str(tmp_path / "ck.pt")
    checkpoint.save(path, updater=up, vector_env=ve, extra={"note": "mid-run"})
    run(ve, ring, up, 5)
    want = (ve.boards.clone(), ring.s.clone(), ring.head_size.clone(), [p.detach().clone() for p in up.model.parameters()])
    torch.manual_seed(99)                         # different init: everything must come from the file
    ve2, ring2, up2 = make()
    assert checkpoint.load(path, updater=up2, vector_env=ve2)["note"] == "mid-run"
    run(ve2, ring2, up2, 5)
    assert torch.equal(ve2.boards, want[0]) and torch.equal(ring2.s, want[1]) and torch.equal(ring2.head_size, want[2])
    for p, q in zip(up2.model.parameters(), want[3]):
        np.testing.assert_allclose(p.detach().cpu().numpy(), q.cpu().numpy(), rtol=1e-12, atol=1e-14)
    assert int(up2.opt.step_count.item()) == 11 == up2.updates


def test_batched_training_loop_follows_the_reference_schedule(cuda):
    """training_loop's per-episode rules on 512 concurrent games: one update per finished episode
    after the warm-up, target syncs at multiples of K, epsilon by episode index."""
    from b2048.train import TrainConfig, epsilon_for, train_batched
    cfg = TrainConfig(n_envs=512, replay_buffer_length=15000, batch_size=256, no_episodes=2500,
                      no_episodes_to_reach_epsilon=1000, no_episodes_before_training=1000,
                      no_episodes_before_updating_target=500, conv=False, seed=3, max_updates_per_step=64,
                      use_graph=False)
    assert epsilon_for(0, cfg) == 1.0 and epsilon_for(500, cfg) == 0.5 and epsilon_for(5000, cfg) == 0.01
    torch.manual_seed(0)
    model = dense_model().to(cuda)
    before = [p.detach().clone() for p in model.parameters()]
    logs = []
    out = train_batched(model, cfg, device=cuda, log_every=50, on_log=logs.append)
    assert out["games"] >= 2500
    owed = out["games"] - 1001                                   # episodes with index > 1000
    assert 0 < out["updates"] <= owed and out["updates"] >= owed - 64 * 2
    assert out["target_syncs"] >= 4                              # episode 0, 500, 1000, 1500, 2000 (several may share a step)
    assert not any(torch.equal(a, b) for a, b in zip(before, model.parameters()))
    assert logs and logs[-1]["epsilon"] <= logs[0]["epsilon"] and np.isfinite(out["final_loss"])
