"""CPU (gloo, world_size 2): the multi-GPU host logic — index sharding, flat-gradient
sum-allreduce, weight broadcast, max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b2048 import dist as bdist


def test_shard_covers_everything_with_aligned_bases():
    for n, g in [(64 << 20, 8), (1000003, 3), (17, 2), (4096, 1), (5, 4)]:
        spans = [bdist.shard(n, r, g) for r in range(g)]
        assert spans[0][0] == 0 and sum(s[1] for s in spans) == n
        for r in range(g):
            assert spans[r][0] % 8 == 0
            if r:
                assert spans[r][0] == spans[r - 1][0] + spans[r - 1][1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = bdist.init_from_env("gloo")
    assert (r, w) == (rank, world) and bdist.world() == (rank, world)
    torch.manual_seed(100 + rank)                      # different initial weights per rank
    model = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.ReLU(), torch.nn.Linear(8, 4)).double()
    bdist.broadcast_module(model)
    w0 = torch.cat([p.detach().flatten() for p in model.parameters()])
    fg = bdist.FlatGrads(model)
    x = torch.full((5, 16), float(rank + 1), dtype=torch.float64)
    fg.zero_()
    model(x).sum().backward()                          # local gradient, written into the flat buffer
    local = fg.flat.clone()
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in model.parameters())
    fg.allreduce_()
    t = bdist.max_over_ranks(float(rank + 1))
    out[rank] = (w0, local, fg.flat.clone(), t)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_and_broadcast_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    (w0a, la, ra, ta), (w0b, lb, rb, tb) = out[0], out[1]
    assert torch.equal(w0a, w0b)                        # broadcast made the weights identical
    assert not torch.equal(la, lb)                      # local gradients differ (different data)
    assert torch.allclose(ra, la + lb, rtol=0, atol=0) and torch.equal(ra, rb)   # SUM, identical on both ranks
    assert ta == tb == 2.0
