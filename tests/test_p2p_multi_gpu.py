"""K5, the fused NVLink allreduce+Adam kernel.  2-GPU test (skipped on a single-GPU box): every rank ends
with bit-identical parameters that agree with the NCCL all_reduce + Adam path.  1-GPU tests: the same
kernel with world = 1 against `ddqn_adam_step` (bitwise), and the bounded wait for a peer that never
arrives (error flag raised, replica untouched)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import copy
    import torch.distributed as dist
    import b2048
    from b2048.rollout import VectorEnv
    from b2048.trainer import DDQNUpdater
    from test_shim_gpu import conv_model
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    ve = VectorEnv(4096, device=dev, seed=5, index_base=rank * 4096)          # different data per rank
    ring = b2048.ReplayRing(15000, device=dev)
    for _ in range(6):
        ve.step(replay=ring)
    torch.manual_seed(0)
    base = conv_model().to(dev)
    res = {}
    for mode in ("nccl", "p2p"):
        ring.head_size[2] = 0
        up = DDQNUpdater(copy.deepcopy(base), ring, batch_size=2000, lr=1e-3, use_graph=(mode == "p2p"),
                         seed=11, exchange=mode)
        for _ in range(5):               # graph capture + warm-up leave no trace (DDQNUpdater._capture)
            up.update()
        torch.cuda.synchronize()
        if mode == "p2p":
            assert not up.exchange.timed_out()
        res[mode] = up.params.flat.detach().cpu().numpy().copy()
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_p2p_allreduce_adam_matches_nccl_and_keeps_replicas_identical():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert np.array_equal(a["p2p"], b["p2p"])                      # replicas bit-identical
    np.testing.assert_allclose(a["p2p"], a["nccl"], rtol=1e-9, atol=1e-9)   # same update as NCCL sum + Adam
    assert not np.array_equal(a["p2p"], np.zeros_like(a["p2p"]))


def _k5_world1(cuda, n, steps, peer_flag_value=None):
    """Drive p2p_allreduce_adam_f64 directly through the C-ABI with world = 1 (the rank is its own peer)."""
    from b2048 import _lib
    dev = cuda.index or 0
    _lib.init(dev)
    g = torch.Generator(device="cpu").manual_seed(n)
    p0 = torch.randn(n, dtype=torch.float64, generator=g).to(cuda)
    grads = [torch.randn(n, dtype=torch.float64, generator=g).to(cuda) for _ in range(steps)]
    res = {}
    for mode in ("adam", "k5"):
        p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        step = torch.zeros(1, dtype=torch.int64, device=cuda)
        gbuf = torch.empty_like(p0)
        flags = torch.zeros(2, dtype=torch.int64, device=cuda)
        sync = torch.zeros(4, dtype=torch.int64, device=cuda)
        pg = torch.tensor([gbuf.data_ptr()], dtype=torch.int64, device=cuda)
        pf = torch.tensor([flags.data_ptr()], dtype=torch.int64, device=cuda)
        st = torch.cuda.current_stream(cuda).cuda_stream
        for gi in grads:
            gbuf.copy_(gi)
            with torch.cuda.device(cuda):
                if mode == "adam":
                    _lib.check(_lib.lib().ddqn_adam_step(p.data_ptr(), gbuf.data_ptr(), m.data_ptr(), v.data_ptr(),
                                                         step.data_ptr(), n, 1e-2, 0.9, 0.999, 1e-8, st))
                else:
                    _lib.check(_lib.lib().p2p_allreduce_adam_f64(pg.data_ptr(), pf.data_ptr(), sync.data_ptr(), 0, 1,
                                                                 p.data_ptr(), m.data_ptr(), v.data_ptr(), step.data_ptr(),
                                                                 n, 1e-2, 0.9, 0.999, 1e-8, st))
        torch.cuda.synchronize()
        res[mode] = (p.cpu(), m.cpu(), v.cpu(), int(step.item()), sync.cpu())
    return res


@pytest.mark.parametrize("n", [1, 255, 33476, 403716])        # incl. the conv and dense parameter counts
def test_k5_world_size_one_equals_fused_adam_bitwise(cuda, n):
    res = _k5_world1(cuda, n, steps=4)
    a, k = res["adam"], res["k5"]
    assert k[3] == a[3] == 4 and int(k[4][0]) == 4 and int(k[4][2]) == 0      # steps, epochs published, no error
    assert torch.equal(a[0], k[0]) and torch.equal(a[1], k[1]) and torch.equal(a[2], k[2])


def test_k5_gives_up_on_a_missing_peer_without_touching_the_replica(cuda, monkeypatch):
    """world = 2 with a 'peer' (a second flag block on the same GPU) that never signals: the bounded wait
    expires, sync_state[2] goes up, and params / moments / step are exactly as before.  The
    wait limit is 30 s of wall clock in the shipped build, so this test only runs when asked to."""
    if not os.environ.get("B2048_TEST_P2P_TIMEOUT"):
        pytest.skip("set B2048_TEST_P2P_TIMEOUT=1 to spend 30 s on the lost-peer path")
    from b2048 import _lib
    dev = cuda.index or 0
    _lib.init(dev)
    n = 1000
    p = torch.randn(n, dtype=torch.float64, device=cuda)
    p_before = p.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device=cuda)
    g0, g1 = torch.randn(n, dtype=torch.float64, device=cuda), torch.randn(n, dtype=torch.float64, device=cuda)
    f0, f1 = torch.zeros(4, dtype=torch.int64, device=cuda), torch.zeros(4, dtype=torch.int64, device=cuda)
    sync = torch.zeros(4, dtype=torch.int64, device=cuda)
    pg = torch.tensor([g0.data_ptr(), g1.data_ptr()], dtype=torch.int64, device=cuda)
    pf = torch.tensor([f0.data_ptr(), f1.data_ptr()], dtype=torch.int64, device=cuda)
    with torch.cuda.device(cuda):
        _lib.check(_lib.lib().p2p_allreduce_adam_f64(pg.data_ptr(), pf.data_ptr(), sync.data_ptr(), 0, 2, p.data_ptr(),
                                                     m.data_ptr(), v.data_ptr(), step.data_ptr(), n, 1e-2, 0.9, 0.999,
                                                     1e-8, torch.cuda.current_stream(cuda).cuda_stream))
    torch.cuda.synchronize()
    assert int(sync[2].item()) == 1 and int(step.item()) == 0 and int(sync[0].item()) == 0
    assert torch.equal(p, p_before) and not m.any() and not v.any()
