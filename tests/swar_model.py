"""Bit-level Python model of the CUDA step kernel's SWAR pipeline (b2048_common.cuh):
byte-permute + masked delta-swap transforms, row-table lookups, transformed-frame legality and the
prefix-multiply spawn.  It lets the kernel's constants and tables be checked against the reference
goldens on a CPU-only box (tests/test_swar_model.py); the GPU tests check the real kernel."""
import numpy as np

M32 = 0xFFFFFFFF


def byte_perm(x, y, s):
    b = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    r = 0
    for i in range(4):
        r |= b[(s >> (4 * i)) & 7] << (8 * i)
    return r


def umulhi(a, b):
    return ((a * b) >> 32) & M32


def nz3(v):
    return ((((v & 0x77777777) + 0x77777777) | v) & 0x88888888) & M32


def ne3_dirty(a, b):
    return ((((a ^ b) & 0x77777777) + 0x77777777) | (a ^ b)) & M32


# act_xform(a): sel_fwd, sel_inv, mul_l, mul_r, mask
XF = [(0x6240 | (0x7351 << 16), 0x6240 | (0x7351 << 16), 1 << 12, 1 << 20, 0x0000F0F0),
      (0x0426 | (0x1537 << 16), 0x5173 | (0x4062 << 16), 1 << 12, 1 << 20, 0x0000F0F0),
      (0x3210 | (0x7654 << 16), 0x3210 | (0x7654 << 16), 1 << 4, 1 << 28, 0),
      (0x2301 | (0x6745 << 16), 0x2301 | (0x6745 << 16), 1 << 4, 1 << 28, 0x0F0F0F0F)]


def zframe_to_legal(a, m):
    b0, b1, b2, b3 = m & 1, (m >> 1) & 1, (m >> 2) & 1, (m >> 3) & 1
    legal = [b0 | b1 << 1 | b2 << 2 | b3 << 3, b1 | b0 << 1 | b2 << 2 | b3 << 3,
             b2 | b3 << 1 | b0 << 2 | b1 << 3, b2 | b3 << 1 | b1 << 2 | b0 << 3][a]
    return legal | (0 if legal else 0x10)


def delta_swap(v, x):
    t = (v ^ umulhi(v, x[3])) & x[4]
    return (v ^ t ^ ((t * x[2]) & M32)) & M32


def slide_board(lut, lo, hi, a):
    """-> (olo, ohi, reward, flags) exactly like the device function."""
    x = XF[a]
    zl = byte_perm(lo, hi, x[0] & 0xFFFF)
    zh = byte_perm(lo, hi, x[0] >> 16)
    zl, zh = delta_swap(zl, x), delta_swap(zh, x)
    idx = [zl & 0xFFFF, zl >> 16, zh & 0xFFFF, zh >> 16]
    e = [int(lut[i]) for i in idx]
    extra = sum(65536 for i in idx if i == 0xEEEE)
    wl = ((e[0] & 0xFFFF) + ((e[1] << 16) & M32)) & M32
    wh = ((e[2] & 0xFFFF) + ((e[3] << 16) & M32)) & M32
    h01 = byte_perm(e[0], e[1], 0x7632)
    h23 = byte_perm(e[2], e[3], 0x7632)
    fl = h01 | h23
    s = ((h01 & 0x3FFF3FFF) + (h23 & 0x3FFF3FFF)) & M32
    reward = (4 * (s & 0xFFFF) + 4 * (s >> 16) + extra) & M32
    changed = (wl ^ zl) | (wh ^ zh)
    n_l, n_h = nz3(zl), nz3(zh)
    v_l, v_h = byte_perm(zl, zh, 0x5432), zh >> 16
    ne_l, ne_h = ne3_dirty(zl, v_l), ne3_dirty(zh, v_h)
    nv_l, nv_h = byte_perm(n_l, n_h, 0x5432), n_h >> 16
    up = ((nv_l & ~(n_l & ne_l)) | (nv_h & ~(n_h & ne_h))) & M32
    dn_l, dn_h = (n_l & ~(nv_l & ne_l)) & M32, (n_h & ~(nv_h & ne_h)) & M32
    m = (1 if changed else 0) | (2 if fl & 0x40004000 else 0) | (4 if up else 0) | (8 if (dn_l | (dn_h & 0xFFFF)) else 0)
    flags = zframe_to_legal(a, m) | (0x20 if changed else 0) | (0x40 if fl & 0x80008000 else 0)
    wl, wh = delta_swap(wl, x), delta_swap(wh, x)
    return byte_perm(wl, wh, x[1] & 0xFFFF), byte_perm(wl, wh, x[1] >> 16), reward, flags


def spawn_kth_empty(lo, hi, w_pos, e):
    e3_lo = (~(((lo & 0x77777777) + 0x77777777) | lo)) & 0x88888888
    e3_hi = (~(((hi & 0x77777777) + 0x77777777) | hi)) & 0x88888888
    e_lo, e_hi = e3_lo >> 3, e3_hi >> 3
    p_lo = (e_lo * 0x11111111) & M32
    c_lo = p_lo >> 28
    p_hi = (e_hi * 0x11111111 + c_lo * 0x11111111) & M32
    cnt = p_hi >> 28
    tgt = (umulhi(w_pos, cnt) * 0x11111111 + 0x11111111) & M32
    h_lo = (~ne3_dirty(p_lo, tgt)) & e3_lo
    h_hi = (~ne3_dirty(p_hi, tgt)) & e3_hi
    return (lo + (h_lo >> 3) * e) & M32, (hi + (h_hi >> 3) * e) & M32, cnt


def spawn_draw16_prefix_form(lo, hi, D):
    """Round-1 form of spawn_draw16's cell choice: shifted masks, inclusive prefix counts, equality test.
    -> (bit-3 mask of the chosen nibble in lo, in hi, frac)"""
    e3_lo = (~(((lo & 0x77777777) + 0x77777777) | lo)) & 0x88888888
    e3_hi = (~(((hi & 0x77777777) + 0x77777777) | hi)) & 0x88888888
    e_lo, e_hi = e3_lo >> 3, e3_hi >> 3
    p_lo = (e_lo * 0x11111111) & M32
    c_lo = p_lo >> 28
    p_hi = (e_hi * 0x11111111 + c_lo * 0x11111111) & M32
    cnt = p_hi >> 28
    prod = D * cnt
    k, frac = prod >> 32, prod & M32
    tgt = (k * 0x11111111 + 0x11111111) & M32
    return (~ne3_dirty(p_lo, tgt)) & e3_lo, (~ne3_dirty(p_hi, tgt)) & e3_hi, frac


def spawn_draw16(lo, hi, D, dlow=False):
    """The device function spawn_draw16 (b2048_common.cuh) bit for bit: exclusive prefix / inclusive suffix counts
    from one wide multiply per half, x = k - prefix, zero-nibble test.  dlow: D = d instead of d << 16."""
    e3_lo = (~(((lo & 0x77777777) + 0x77777777) | lo)) & 0x88888888
    e3_hi = (~(((hi & 0x77777777) + 0x77777777) | hi)) & 0x88888888
    q = e3_lo * 0x22222222
    r = e3_hi * 0x22222222
    q_lo, q_hi, r_lo, r_hi = q & M32, q >> 32, r & M32, r >> 32
    ep_hi = (r_lo + q_lo + q_hi) & M32
    cnt = ((ep_hi + r_hi) & M32) & (0x000F0000 if dlow else 15)
    prod = D * cnt
    assert prod < 1 << 64
    k, frac = prod >> 32, prod & M32
    km = (k * 0x11111111) & M32
    x_lo, x_hi = (km - q_lo) & M32, (km - ep_hi) & M32

    def zero_and_empty(x, e3):
        t = ((x & 0x77777777) + 0x77777777) & M32
        return (~(t | x)) & e3 & M32
    return zero_and_empty(x_lo, e3_lo), zero_and_empty(x_hi, e3_hi), frac


def perp_legal_single(zl, zh):
    """Perpendicular legality of ONE transformed board, the one-board form of slide_board: (toward row 0, toward row 3)."""
    n_l, n_h = nz3(zl), nz3(zh)
    v_l, v_h = byte_perm(zl, zh, 0x5432), zh >> 16
    ne_l, ne_h = ne3_dirty(zl, v_l), ne3_dirty(zh, v_h)
    nv_l, nv_h = byte_perm(n_l, n_h, 0x5432), n_h >> 16
    up = ((nv_l & ~(n_l & ne_l)) | (nv_h & ~(n_h & ne_h))) & M32
    dn_l, dn_h = (n_l & ~(nv_l & ne_l)) & M32, (n_h & ~(nv_h & ne_h)) & M32
    return bool(up), bool(dn_l | (dn_h & 0xFFFF))


def perp_legal_pair(a, b):
    """The streaming kernel's pair step (env_kernels.cu: pair_legal) for two transformed boards a = (zl, zh),
    b = (zl, zh): the third vertical pairs of both boards share one word.  -> ((upA, dnA), (upB, dnB))."""
    P, Q = byte_perm(a[1], b[1], 0x5410), byte_perm(a[1], b[1], 0x7632)      # rows 2 / rows 3 of A and B
    nP, nQ, neP = nz3(P), nz3(Q), ne3_dirty(P, Q)
    up3, dn3 = (nQ & ~(nP & neP)) & M32, (nP & ~(nQ & neP)) & M32
    out = []
    for (zl, zh), sel, half in ((a, 0x5432, 0x0000FFFF), (b, 0x7632, 0xFFFF0000)):
        n, v = nz3(zl), byte_perm(zl, zh, 0x5432)
        ne = ne3_dirty(zl, v)
        nv = byte_perm(n, nP, sel)
        out.append((bool(((nv & ~(n & ne)) | (up3 & half)) & M32), bool(((n & ~(nv & ne)) | (dn3 & half)) & M32)))
    return tuple(out)
