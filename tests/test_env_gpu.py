"""GPU parity tests for K1 (env step): CUDA path through the C-ABI vs the reference goldens
(bit-exact) and vs the C oracle on seeded inputs, plus size-independent properties at full size."""
import os

import numpy as np
import pytest
import torch

import b2048
from b2048 import env
from oracle import board_oracle as bo

pytestmark = pytest.mark.gpu


def u64(t):
    return t.cpu().numpy().view(np.uint64)


def to_dev(a_u64, cuda):
    return torch.from_numpy(np.ascontiguousarray(a_u64).view(np.int64)).to(cuda)


@pytest.fixture(scope="module")
def boards_g(golden_dir):
    return np.load(os.path.join(golden_dir, "boards.npz"))


@pytest.fixture(scope="module")
def games_g(golden_dir):
    return np.load(os.path.join(golden_dir, "games.npz"))


def spawn_override_from(cells, vals):
    o = np.full(cells.shape, 0xFF, dtype=np.uint8)
    m = cells >= 0
    o[m] = (cells[m].astype(np.int64) | (np.log2(vals[m]).astype(np.int64) << 4)).astype(np.uint8)
    return o


def test_step_matches_reference_boards_all_actions(cuda, boards_g):
    """Slide/merge, reward, legal mask, done, changed and the replayed spawn are bit-exact against
    src/board.py on 3003 boards x 4 actions (dense, sparse, dead, and 16384/32768-tile boards)."""
    g = boards_g
    st = g["state"]
    ok_in = st.max(axis=1) <= 32768
    packed = to_dev(bo.pack(st), cuda)
    for a in range(4):
        ovr = spawn_override_from(g["spawn_cell"][:, a], g["spawn_val"][:, a])
        actions = torch.full((len(st),), a, dtype=torch.uint8, device=cuda)
        nxt, rew, flg = env.step(packed, actions, spawn_override=torch.from_numpy(ovr).to(cuda))
        flg = flg.cpu().numpy()
        overflow = (g["slide"][:, a].max(axis=1) > 32768)
        ok = ok_in & ~overflow
        assert np.array_equal((flg & 0x40) != 0, overflow)
        assert np.array_equal(u64(nxt)[ok], bo.pack(g["next"][:, a])[ok])
        assert np.array_equal(rew.cpu().numpy()[ok], g["reward"][:, a][ok].astype(np.int32))
        assert np.array_equal(flg & 0x0F, g["legal"])
        assert np.array_equal((flg & 0x10) != 0, g["legal"] == 0)
        assert np.array_equal((flg & 0x20) != 0, g["spawn_cell"][:, a] >= 0)
        assert not (flg & 0x80).any()


def test_step_all4_matches_reference_boards(cuda, boards_g):
    g = boards_g
    st = g["state"]
    packed = to_dev(bo.pack(st), cuda)
    ovr = np.stack([spawn_override_from(g["spawn_cell"][:, a], g["spawn_val"][:, a]) for a in range(4)], axis=1)
    nxt4, rew4, flg = env.step_all4(packed, spawn_override4=torch.from_numpy(np.ascontiguousarray(ovr)).to(cuda))
    flg = flg.cpu().numpy()
    assert np.array_equal(flg & 0x0F, g["legal"])
    assert np.array_equal((flg & 0x10) != 0, g["legal"] == 0)
    for a in range(4):
        ok = g["slide"][:, a].max(axis=1) <= 32768
        assert np.array_equal(u64(nxt4[:, a].contiguous())[ok], bo.pack(g["next"][:, a])[ok])
        assert np.array_equal(rew4[:, a].cpu().numpy()[ok], g["reward"][:, a][ok].astype(np.int32))


def test_bench_stream_head_matches_reference(cuda, golden_dir):
    """SURVEY 8(d) parity subset: the device generators produce the bench stream the goldens were made from,
    and the streaming kernel (1 Mi boards, so the persistent shared-memory-table path runs) agrees with the
    REFERENCE on its first 65536 boards: slide-only successor (spawn off through the override hook), reward,
    legal mask, done."""
    d = np.load(os.path.join(golden_dir, "bench_stream.npz"))
    m, n = len(d["boards"]), 1 << 20
    boards = env.random_boards(n, seed=2048, device=cuda)
    actions = env.random_actions(n, seed=2050, device=cuda)
    assert np.array_equal(u64(boards[:m]), d["boards"]) and np.array_equal(actions[:m].cpu().numpy(), d["actions"])
    assert np.array_equal(u64(boards[:4096]), bo.stream_boards(4096, seed=2048))
    base = 1 << 35                                          # generators honour the global index base
    assert np.array_equal(u64(env.random_boards(4099, seed=9, index_base=base + 3, p_empty=0.5, max_exp=15, device=cuda)),
                          bo.stream_boards(4099, seed=9, index_base=base + 3, p_empty=0.5, max_exp=15))
    assert np.array_equal(env.random_actions(4099, seed=9, step_index=5, index_base=base + 3, device=cuda).cpu().numpy(),
                          bo.stream_actions(4099, seed=9, step=5, index_base=base + 3))
    skip = torch.full((n,), 0xFE, dtype=torch.uint8, device=cuda)
    nxt, rew, flg = env.step(boards, actions, spawn_override=skip)
    f = flg[:m].cpu().numpy()
    assert np.array_equal(u64(nxt[:m]), d["slide"]) and np.array_equal(rew[:m].cpu().numpy(), d["reward"])
    assert np.array_equal(f & 0x0F, d["legal"]) and np.array_equal((f & 0x10) != 0, d["legal"] == 0)
    # with spawns: identical to the slide-only board except for one new 2 / 4 in an empty cell, iff changed
    nxt2, rew2, flg2 = env.step(boards, actions, seed=7, step_index=3)
    assert torch.equal(rew2, rew) and torch.equal(flg2, flg)
    x = (u64(nxt2[:m]) ^ d["slide"])
    changed = (f & 0x20) != 0
    assert np.array_equal(x != 0, changed)
    nib = np.array([bin(int(v)).count("1") for v in x[changed][:2000]])
    assert (nib == 1).all()


def test_step_replays_reference_games(cuda, games_g):
    t = games_g
    S, A, R, S2, D = t["state"], t["action"], t["reward"], t["next"], t["done"]
    # spawn = the single cell where the reference's next differs from the slide-only board
    slide_tiles = np.array([bo.slide(S[i], int(A[i]))[0].reshape(16) for i in range(len(S))])
    diff = slide_tiles != S2
    cells = np.where(diff.any(axis=1), diff.argmax(axis=1), -1)
    vals = np.where(cells >= 0, S2[np.arange(len(S)), np.maximum(cells, 0)], 0)
    ovr = spawn_override_from(cells, vals)
    nxt, rew, flg = env.step(to_dev(bo.pack(S), cuda), torch.from_numpy(A).to(cuda),
                             spawn_override=torch.from_numpy(ovr).to(cuda))
    assert np.array_equal(u64(nxt), bo.pack(S2))
    assert np.array_equal(rew.cpu().numpy(), R.astype(np.int32))
    assert np.array_equal((flg.cpu().numpy() & 0x10) != 0, D.astype(bool))


@pytest.mark.parametrize("n,index_base,offset", [
    (1, 0, 0), (2, 0, 0), (1000, 0, 0), (1001, 7, 0), (4097, 0, 1),        # small / unaligned kernel
    (1 << 20, 0, 0), ((1 << 20) + 1, 0, 0), (1 << 20, 12345, 0),            # streaming kernel (+tail, odd base)
    ((1 << 20) + 7, 8, 0), ((1 << 19) + 8, 1 << 40, 0),                     # 7-board tail, one spare octet, huge base
    (1 << 20, 1, 0), (1 << 20, 2, 0), (1 << 20, 3, 0), (1 << 20, 4, 0),     # every misalignment of the index base
    (1 << 20, 5, 0), (1 << 20, 6, 0), (1 << 20, 7, 0),                      #   against the 8-board Philox groups
    (1 << 20, 1 << 33, 1), (1 << 20, 0, 2),                                 # misaligned views -> small kernel
    (1 << 20, 0, 4),                                                        # 32-byte aligned view -> streaming kernel
])
def test_step_matches_oracle_with_philox_spawns(cuda, n, index_base, offset):
    """No override: the library's Philox spawn stream is reproduced by the oracle bit for bit, for
    both kernels, odd sizes, odd index bases and unaligned views."""
    boards_all = env.random_boards(n + offset, seed=2048, index_base=index_base, device=cuda)
    actions_all = env.random_actions(n + offset, seed=2050, step_index=9, index_base=index_base, device=cuda)
    boards, actions = boards_all[offset:], actions_all[offset:]
    nxt, rew, flg = env.step(boards, actions, seed=0xDEADBEEFCAFE, step_index=41, index_base=index_base)
    o_nxt, o_rew, o_flg = bo.step_packed(u64(boards), actions.cpu().numpy(), seed=0xDEADBEEFCAFE, step=41,
                                         index_base=index_base, threads=bo.num_threads())
    assert np.array_equal(u64(nxt), o_nxt)
    assert np.array_equal(rew.cpu().numpy(), o_rew)
    assert np.array_equal(flg.cpu().numpy(), o_flg)


@pytest.mark.parametrize("index_base,max_exp", [(0, 15), (12345, 15), (3, 14)])
def test_step_fifty_percent_fours_and_high_tiles(cuda, index_base, max_exp):
    """p4 = 0.5 (the reference's spawn rule) and boards with 16384/32768 tiles (rows
    that miss the shared-memory table: the streaming kernel redoes those quads from the global
    one), also with an index base that is not a multiple of four."""
    n = 1 << 20
    boards = env.random_boards(n, seed=5, p_empty=0.25, max_exp=max_exp, device=cuda)
    actions = env.random_actions(n, seed=6, device=cuda)
    nxt, rew, flg = env.step(boards, actions, seed=1, step_index=2, p4=env.P4_FIFTY_PERCENT, index_base=index_base)
    o_nxt, o_rew, o_flg = bo.step_packed(u64(boards), actions.cpu().numpy(), seed=1, step=2, index_base=index_base,
                                         p4_threshold=env.P4_FIFTY_PERCENT, threads=bo.num_threads())
    f = flg.cpu().numpy()
    assert np.array_equal(f, o_flg)
    ok = (f & 0x40) == 0
    assert (max_exp < 15) == bool(ok.all())     # 32768+32768 merges exist (max_exp 15) and are flagged
    assert np.array_equal(u64(nxt)[ok], o_nxt[ok])
    assert np.array_equal(rew.cpu().numpy()[ok], o_rew[ok])


def test_spawn_draws_are_uniform_and_value_is_independent_of_the_cell(cuda):
    """Spawn stream v2 gives each board ONE 16-bit Philox lane d: rank = floor(d*n/65536), "4" iff the
    fraction of d*n falls below p4.  Check on 16 Mi boards that, for every number of empty cells n, all
    ranks are equally likely and the share of fours does not depend on the rank (5-sigma bands)."""
    n = 1 << 24
    boards = env.random_boards(n, seed=99, p_empty=0.5, max_exp=6, device=cuda)
    actions = env.random_actions(n, seed=98, device=cuda)
    skip = torch.full((n,), 0xFE, dtype=torch.uint8, device=cuda)
    slid, _, flg = env.step(boards, actions, spawn_override=skip)
    nxt, _, _ = env.step(boards, actions, seed=1234, step_index=7)
    changed = (flg & 0x20) != 0
    sh = 4 * torch.arange(16, device=cuda, dtype=torch.int64)
    es = ((slid[:, None] >> sh) & 0xF)[changed]
    en = ((nxt[:, None] >> sh) & 0xF)[changed]
    diff = es != en
    assert bool((diff.sum(dim=1) == 1).all())                       # exactly one new tile
    cell = diff.to(torch.int64).argmax(dim=1)
    val = en.gather(1, cell[:, None])[:, 0]
    assert bool(((val == 1) | (val == 2)).all()) and bool((es.gather(1, cell[:, None]) == 0).all())
    empty = es == 0
    n_empty = empty.sum(dim=1)
    rank = (empty.cumsum(dim=1) - 1).gather(1, cell[:, None])[:, 0]
    for ne in range(1, 15):
        m = n_empty == ne
        tot = int(m.sum())
        if tot < 20000:
            continue
        cnt = torch.bincount(rank[m], minlength=ne).double()
        p = 1.0 / ne
        sigma = (tot * p * (1 - p)) ** 0.5
        assert float((cnt - tot * p).abs().max()) <= 5 * sigma + 1, (ne, cnt.tolist())
        fours = torch.bincount(rank[m], weights=(val[m] == 2).double(), minlength=ne)
        for r in range(ne):
            c = float(cnt[r])
            assert abs(float(fours[r]) - 0.1 * c) <= 5 * (c * 0.09) ** 0.5 + 1, (ne, r, float(fours[r]), c)


def test_stream_kernel_split_launches(cuda):
    """launch_step cuts a batch into launches of at most STREAM_MAX_OCTS octets (2^33 boards in the
    shipped library: unreachable in a test).  build.py also builds a copy of the library with the limit
    lowered to 40 000 octets; the 1 Mi-board call below is then 4 streaming launches with different
    pointer / index-base offsets, and must still agree with the oracle bit for bit."""
    import subprocess, sys, textwrap
    lib = os.path.join(os.path.dirname(b2048._lib.LIB_PATH), "libb2048_splittest.so")
    if not os.path.exists(lib):
        pytest.fail(f"{lib} missing: run __graft_entry__.build()")
    code = textwrap.dedent("""
        import sys, numpy as np, torch
        sys.path[:0] = [%r, %r]
        from b2048 import env
        from oracle import board_oracle as bo
        dev = torch.device("cuda:0")
        n = (1 << 20) + 13
        for base in (0, 5):
            b = env.random_boards(n, seed=31, device=dev); a = env.random_actions(n, seed=32, device=dev)
            nxt, rew, flg = env.step(b, a, seed=77, step_index=3, index_base=base)
            o = bo.step_packed(b.cpu().numpy().view(np.uint64), a.cpu().numpy(), seed=77, step=3, index_base=base,
                               threads=bo.num_threads())
            assert np.array_equal(nxt.cpu().numpy().view(np.uint64), o[0]) and np.array_equal(rew.cpu().numpy(), o[1])
            assert np.array_equal(flg.cpu().numpy(), o[2])
        print("SPLIT-OK")
    """) % (os.path.dirname(os.path.dirname(b2048._lib.LIB_PATH)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B2048_LIB=lib), capture_output=True, text=True,
                       timeout=600)
    assert "SPLIT-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_sharding_is_invisible(cuda):
    """Global-index Philox counters: stepping [0,n) at once == stepping G shards with index_base
    (the multi-GPU partition of SURVEY §8e), bit for bit."""
    n = 3 * (1 << 19)
    boards = env.random_boards(n, seed=2048, device=cuda)
    actions = env.random_actions(n, seed=2050, device=cuda)
    whole = env.step(boards, actions, seed=3, step_index=4)
    for G in (2, 3, 8):
        per = n // G
        for r in range(G):
            sl = slice(r * per, (r + 1) * per if r < G - 1 else n)
            part = env.step(boards[sl].clone(), actions[sl].clone(), seed=3, step_index=4, index_base=sl.start)
            for w, p in zip(whole, part):
                assert torch.equal(w[sl], p)


@pytest.mark.parametrize("n,max_exp,index_base", [(1 << 16, 11, 0), ((1 << 20) + 5, 11, 0), (1 << 20, 15, 8 << 30),
                                                  (1 << 20, 11, 3)])
def test_all4_consistent_with_step(cuda, n, max_exp, index_base):
    """b2048_step_all4 (BASELINE config 2) == four b2048_step calls, for the one-board-per-thread kernel
    (64 Ki boards; also index bases that are not multiples of 8) and for the persistent shared-memory-table
    kernel (>= 512 Ki boards, incl. a ragged tail and boards with 16384 / 32768 tiles that take its cold path)."""
    boards = env.random_boards(n, seed=11, max_exp=max_exp, device=cuda)
    nxt4, rew4, flg4 = env.step_all4(boards, seed=8, step_index=1, index_base=index_base)
    for a in range(4):
        nxt, rew, flg = env.step(boards, torch.full((n,), a, dtype=torch.uint8, device=cuda), seed=8, step_index=1,
                                 index_base=index_base)
        ok = (flg & 0x40) == 0                           # next / reward are unspecified for overflowing moves
        assert torch.equal(nxt4[:, a][ok], nxt[ok]) and torch.equal(rew4[:, a][ok], rew[ok])
        assert torch.equal(flg4 & 0x1F, flg & 0x1F)
        assert torch.equal(((flg4 >> a) & 1).bool(), (flg & 0x20) != 0)   # legal == changed
        assert bool(((flg4 & 0x40) >= (flg & 0x40)).all())                # an overflow in any direction is flagged
    if max_exp < 15:
        assert not bool((flg4 & 0x40).any())
    # against the oracle too (all four successors of the first boards)
    m = min(n, 4096)
    for a in range(4):
        o = bo.step_packed(u64(boards[:m]), np.full(m, a, dtype=np.uint8), seed=8, step=1, index_base=index_base)
        ok = (o[2] & 0x40) == 0
        assert np.array_equal(u64(nxt4[:m, a].contiguous())[ok], o[0][ok]) and np.array_equal(rew4[:m, a].cpu().numpy()[ok], o[1][ok])


def test_legal_mask_reset_pack_unpack(cuda, boards_g):
    st = boards_g["state"]
    ok = st.max(axis=1) <= 32768
    tiles = torch.from_numpy(st[ok]).to(cuda)
    packed = env.pack(tiles)
    assert np.array_equal(u64(packed), bo.pack(st[ok]))
    assert torch.equal(env.unpack_tiles(packed), tiles)
    f = env.unpack_f64(packed)
    assert np.array_equal(f.cpu().numpy(), bo.exponents(bo.pack(st[ok])))
    assert env.unpack_f64(packed, conv=True).shape == (ok.sum(), 1, 4, 4)
    assert np.array_equal(env.legal_mask(packed).cpu().numpy() & 0xF, boards_g["legal"][ok])
    with pytest.raises(ValueError):
        env.pack(torch.tensor([[3] + [0] * 15], dtype=torch.int64, device=cuda))
    # reset: zeros + two spawns, reproduces the oracle's reset stream; masked reset touches only DONE boards
    fresh = env.new_boards(100000, device=cuda, seed=4, step_index=6, index_base=10)
    assert np.array_equal(u64(fresh), bo.reset_packed(100000, seed=4, step=6, index_base=10))
    t = env.unpack_tiles(fresh)
    assert ((t != 0).sum(dim=1) == 2).all() and set(torch.unique(t).tolist()) <= {0, 2, 4}
    flags = torch.zeros(100000, dtype=torch.uint8, device=cuda)
    flags[::3] = 0x10
    b2 = env.random_boards(100000, seed=1, device=cuda)
    keep = b2.clone()
    env.reset(b2, seed=4, step_index=6, index_base=10, where_flags=flags)
    assert torch.equal(b2[::3], fresh[::3]) and torch.equal(b2[1::3], keep[1::3]) and torch.equal(b2[2::3], keep[2::3])


def test_empty_and_bad_inputs(cuda):
    e = torch.empty(0, dtype=torch.int64, device=cuda)
    nxt, rew, flg = env.step(e, torch.empty(0, dtype=torch.uint8, device=cuda))
    assert nxt.numel() == 0
    with pytest.raises(ValueError):
        env.step(torch.zeros(4, dtype=torch.int64, device=cuda), torch.zeros(3, dtype=torch.uint8, device=cuda))
    with pytest.raises(Exception):
        env.step(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.uint8))   # CPU tensors: no fallback
    # override naming an occupied cell is reported, not applied
    b = env.pack(torch.tensor([[2, 0, 0, 2] + [0] * 12], dtype=torch.int64, device=cuda))
    nxt, _, flg = env.step(b, torch.tensor([2], dtype=torch.uint8, device=cuda),
                           spawn_override=torch.tensor([0x10], dtype=torch.uint8, device=cuda))
    assert int(flg[0]) & 0x80 and env.unpack_tiles(nxt)[0].tolist() == [4] + [0] * 15


def test_step_host_buffers_match_device_path(cuda):
    n = (9 << 20) + 3          # > 2 chunks of the host pipeline, odd
    boards = env.random_boards(n, seed=2048, device=cuda)
    actions = env.random_actions(n, seed=2050, device=cuda)
    want = env.step(boards, actions, seed=5, step_index=6)
    hb, ha = boards.cpu().pin_memory(), actions.cpu().pin_memory()
    hn = torch.empty(n, dtype=torch.int64).pin_memory()
    hr = torch.empty(n, dtype=torch.int32).pin_memory()
    hf = torch.empty(n, dtype=torch.uint8).pin_memory()
    env.step_host(hb, ha, hn, hr, hf, seed=5, step_index=6)
    assert torch.equal(hn, want[0].cpu()) and torch.equal(hr, want[1].cpu()) and torch.equal(hf, want[2].cpu())


def test_full_size_properties(cuda):
    """BASELINE full size (64M boards, one step): size-independent invariants of a 2048 move —
    total face value grows by exactly the spawned tile iff the board changed; reward is a
    non-negative multiple of 4; CHANGED <=> the action's legal bit; DONE <=> no legal bit."""
    n = 1 << 26
    boards = env.random_boards(n, seed=2048, device=cuda)
    actions = env.random_actions(n, seed=2050, device=cuda)
    nxt, rew, flg = env.step(boards, actions, seed=1, step_index=0)
    sh = (4 * torch.arange(16, device=cuda, dtype=torch.int64))
    chunk = 1 << 22
    n4 = 0
    nchanged = 0
    for s in range(0, n, chunk):
        sl = slice(s, s + chunk)
        eb = (boards[sl, None] >> sh) & 0xF
        en = (nxt[sl, None] >> sh) & 0xF
        face_b = torch.where(eb > 0, 1 << eb, 0).sum(dim=1)
        face_n = torch.where(en > 0, 1 << en, 0).sum(dim=1)
        d = face_n - face_b
        f = flg[sl].to(torch.int64)
        changed = (f & 0x20) != 0
        assert ((d == 0) | (d == 2) | (d == 4)).all()
        assert torch.equal(d != 0, changed)
        assert torch.equal(changed, ((f >> actions[sl].to(torch.int64)) & 1) == 1)
        assert torch.equal((f & 0x10) != 0, (f & 0xF) == 0)
        r = rew[sl]
        assert (r >= 0).all() and (r % 4 == 0).all()
        # tiles never decrease in count by more than the merges implied by the reward
        n4 += int((d == 4).sum())
        nchanged += int(changed.sum())
    assert 0.095 < n4 / nchanged < 0.105        # 10 % fours
