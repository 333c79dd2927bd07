"""CPU: the kernel's SWAR pipeline (Python bit-model in tests/swar_model.py, fed with the row table
the C-ABI library builds on the host) against the reference goldens — slide, reward, legal mask,
done, changed, overflow for 3003 boards x 4 actions, plus the spawn selection vs the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from b2048 import env
from oracle import board_oracle as bo
import swar_model as sm


def test_model_matches_reference_boards(golden_dir):
    g = np.load(os.path.join(golden_dir, "boards.npz"))
    lut = env.row_lut_host()
    st = g["state"]
    packed = bo.pack(st)
    rng = np.random.default_rng(0)
    n_spawn = 0
    for i in range(0, len(st), 2):
        b = int(packed[i])
        lo, hi = b & sm.M32, b >> 32
        for a in range(4):
            nl, nh, rew, flags = sm.slide_board(lut, lo, hi, a)
            assert flags & 0x0F == int(g["legal"][i]), (i, a)
            assert bool(flags & 0x10) == (int(g["legal"][i]) == 0)
            assert bool(flags & 0x20) == bool(g["legal"][i] >> a & 1)
            sl = g["slide"][i, a]
            if sl.max() > 32768:
                assert flags & 0x40
                continue
            assert not flags & 0x40
            want = int(bo.pack(sl[None])[0])
            assert ((nh << 32) | nl) == want and rew == int(g["reward"][i, a]), (i, a)
            if flags & 0x20 and i % 4 == 0:
                w = int(rng.integers(0, 2 ** 32))
                sl2, sh2, cnt = sm.spawn_kth_empty(nl, nh, w, 1)
                ne = int((sl == 0).sum())
                assert cnt == ne
                t = sl.copy()
                t[np.nonzero(sl == 0)[0][(w * ne) >> 32]] = 2
                assert ((sh2 << 32) | sl2) == int(bo.pack(t[None])[0])
                n_spawn += 1
    assert n_spawn > 500


def test_row_0xEEEE_reward_goes_through_the_global_path():
    lut = env.row_lut_host()
    # [16384]*4 in one row: two merges of 32768 = 65536, does not fit the 14-bit field
    nl, nh, rew, flags = sm.slide_board(lut, 0xEEEE, 0, 2)
    assert rew == 65536 and nl == 0x00FF and not flags & 0x40


def test_lut_bank_swizzle_is_a_bijection_that_spreads_the_banks():
    """csrc/b2048_common.cuh::lut_swizzle: i -> i ^ ((i >> 6) & 31) permutes the staged part of the row table
    (rows < 57344) onto itself, and on the benchmark's row distribution (30 % empty cells) it cuts the
    expected shared-memory wavefronts per warp lookup from ~6.3 to ~3.7 (uniformly random banks: ~3.5)."""
    import numpy as np
    i = np.arange(57344)
    j = i ^ ((i >> 6) & 31)
    assert j.max() < 57344 and np.array_equal(np.sort(j), i)
    rng = np.random.default_rng(0)
    cells = np.where(rng.random((4000 * 32, 4)) < 0.3, 0, rng.integers(1, 12, (4000 * 32, 4)))
    rows = (cells[:, 0] | cells[:, 1] << 4 | cells[:, 2] << 8 | cells[:, 3] << 12).reshape(4000, 32)

    def wavefronts(idx):      # per warp: the most distinct addresses that fall into one of the 32 banks
        return np.mean([np.bincount(np.unique(w) & 31, minlength=32).max() for w in idx])

    plain, swz = wavefronts(rows), wavefronts(rows ^ ((rows >> 6) & 31))
    assert plain > 5.5 and swz < 4.0, (plain, swz)


def test_spawn_cell_choice_every_occupancy_pattern():
    """spawn_draw16 picks the empty cell of row-major rank k = floor(d * n / 65536): the device form (prefix and
    suffix counts from one wide multiply per half, zero-nibble test, optional low-half draw) against the round-1
    form and against a plain loop, for every occupancy pattern with at least one tile and draws on both sides of
    every rank boundary."""
    rng = np.random.default_rng(5)
    for pat in range(1, 1 << 16):                 # bit i = cell i occupied
        lo = hi = 0
        vals = rng.integers(1, 16, size=16)
        for i in range(8):
            if pat >> i & 1:
                lo |= int(vals[i]) << (4 * i)
            if pat >> (8 + i) & 1:
                hi |= int(vals[8 + i]) << (4 * i)
        empties = [i for i in range(16) if not pat >> i & 1]
        n = len(empties)
        ds = {0, 65535, int(rng.integers(0, 65536))}
        for j in range(1, n):                     # smallest d of rank j and the one before it
            d = (j * 65536 + n - 1) // n
            ds.update((d - 1, min(d, 65535)))
        for d in ds:
            want = sm.spawn_draw16_prefix_form(lo, hi, d << 16)
            assert sm.spawn_draw16(lo, hi, d << 16) == want, (hex(pat), d)
            assert sm.spawn_draw16(lo, hi, d, dlow=True) == want, (hex(pat), d)
            if n:
                cell = empties[(d * n) >> 16]
                h = (want[0] | (want[1] << 32)) >> 3
                assert h == 1 << (4 * cell), (hex(pat), d)
            else:
                assert want[0] == 0 and want[1] == 0


def test_pair_step_equals_the_single_board_legality():
    """The streaming kernel tests the vertical row pairs of TWO boards in three words (third pairs packed together):
    same answers as the one-board form and as a plain column scan, on random boards (sparse to full, tiles up to
    32768) and on structured ones."""
    rng = np.random.default_rng(11)

    def scan(zl, zh):
        rows = [zl & 0xFFFF, zl >> 16, zh & 0xFFFF, zh >> 16]
        cell = lambda r, c: (rows[r] >> (4 * c)) & 15
        up = any(cell(r + 1, c) and (cell(r, c) == 0 or cell(r, c) == cell(r + 1, c)) for r in range(3) for c in range(4))
        dn = any(cell(r, c) and (cell(r + 1, c) == 0 or cell(r, c) == cell(r + 1, c)) for r in range(3) for c in range(4))
        return up, dn

    boards = [(0, 0), (0xFFFFFFFF, 0xFFFFFFFF), (0x12341234, 0x12341234), (0x00010000, 0), (0, 0x00010000),
              (0x43211234, 0x43211234), (0x21212121, 0x12121212), (0x0000FFFF, 0), (0, 0xFFFF0000)]
    for p_empty in (0.0, 0.3, 0.7, 0.95):
        for _ in range(600):
            cells = np.where(rng.random(16) < p_empty, 0, rng.integers(1, 16, size=16))
            v = sum(int(c) << (4 * i) for i, c in enumerate(cells))
            boards.append((v & 0xFFFFFFFF, v >> 32))
    for i in range(0, len(boards) - 1):
        a, b = boards[i], boards[(i * 7 + 3) % len(boards)]
        got = sm.perp_legal_pair(a, b)
        assert got[0] == sm.perp_legal_single(*a) == scan(*a), (hex(a[0]), hex(a[1]))
        assert got[1] == sm.perp_legal_single(*b) == scan(*b), (hex(b[0]), hex(b[1]))
