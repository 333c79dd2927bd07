"""CPU oracle for the Board2048 environment step — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product (reinforcement-learning-2048_b200/) never does: it has no CPU
fallback.

Two independent restatements of the reference's algorithm live here:

* ``py_*`` — pure Python/numpy, written to follow ``src/board.py`` line by line on the reference's
  own data type (``np.int64[4,4]`` tile values).  Small cases only.
* ``liboracle.so`` (``board_oracle.c``) — the same algorithm in plain C for batches, plus the
  packed-u64 mirror of ``include/b2048.h`` and the library's documented Philox spawn rule.

Both are pinned against ``tests/golden/*.npz`` (outputs of the reference itself, produced by
``oracle/gen_golden.py``) in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ACTIONS = ("up", "down", "left", "right")  # src/board.py:129,191

F_DONE, F_CHANGED, F_OVERFLOW, F_BADSPAWN = 0x10, 0x20, 0x40, 0x80


def build(force: bool = False) -> str:
    """Compile board_oracle.c with gcc (no network, no extra deps)."""
    src = os.path.join(_HERE, "board_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        i64p = ctypes.POINTER(ctypes.c_int64)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i32p = ctypes.POINTER(ctypes.c_int32)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        L.oracle_slide.argtypes = [i64p, ctypes.c_int, i64p, i64p]
        L.oracle_slide.restype = ctypes.c_int
        L.oracle_legal_mask.argtypes = [i64p]
        L.oracle_legal_mask.restype = ctypes.c_int
        L.oracle_row_left.argtypes = [i64p, i64p, i64p]
        L.oracle_pack.argtypes = [i64p]
        L.oracle_pack.restype = ctypes.c_uint64
        L.oracle_unpack.argtypes = [ctypes.c_uint64, i64p]
        L.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
        step_args = [u64p, u8p, u64p, i32p, u8p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64,
                     ctypes.c_uint64, ctypes.c_uint32, u8p]
        L.oracle_step_packed.argtypes = step_args
        L.oracle_step_packed_mt.argtypes = step_args + [ctypes.c_int]
        L.oracle_reset_packed.argtypes = [u64p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64,
                                          ctypes.c_uint64, ctypes.c_uint32]
        L.oracle_legal_mask_packed.argtypes = [u64p, u8p, ctypes.c_int64]
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


# ----------------------------------------------------------------------------------------------
# pure-Python restatement (small cases)
# ----------------------------------------------------------------------------------------------

def py_row_left(vector) -> tuple[np.ndarray, int]:
    """src/board.py:92-126 ``_apply_action_to_vector`` — same cursor walk, returns (row, merge score)."""
    v = np.array(vector, dtype=np.int64).copy()
    score = 0
    current = 0
    while current < len(v) - 1:
        nz = np.where(v != 0)[0]
        if len(nz) == 0 or nz[-1] <= current:
            return v, score
        nz = nz[current < nz]
        if len(nz) == 0:
            return v, score
        j = nz[0]
        if v[current] == 0:
            v[current] += v[j]
            v[j] = 0
        elif v[current] == v[j]:
            v[current] += v[j]
            score += int(v[current])
            v[j] = 0
            current += 1
        elif current + 1 == j:
            current += 1
        else:
            v[current + 1] = v[j]
            v[j] = 0
            current += 1
    return v, score


def py_slide(state: np.ndarray, action: int) -> tuple[np.ndarray, int, bool]:
    """src/board.py:147-183 without the spawn.  Returns (new state, reward, changed)."""
    s = np.asarray(state, dtype=np.int64).reshape(4, 4)
    lines = s.T if action < 2 else s           # up/down work on rows of state.T
    flip = action in (1, 3)                    # down/right reverse each vector first
    out = np.zeros_like(lines)
    reward = 0
    for i in range(4):
        vec = lines[i][::-1] if flip else lines[i]
        res, sc = py_row_left(vec)
        reward += sc
        out[i] = res[::-1] if flip else res
    out = out.T if action < 2 else out
    return out.copy(), reward, bool((out != s).any())


def py_legal_mask(state: np.ndarray) -> int:
    """src/board.py:128-135: a move is legal iff it changes the board."""
    return sum(1 << a for a in range(4) if py_slide(state, a)[2])


def py_pack(state: np.ndarray) -> int:
    """Tile values [4,4] -> packed u64 (include/b2048.h): nibble 4r+c = log2(tile), 0 = empty."""
    b = 0
    for i, t in enumerate(np.asarray(state, dtype=np.int64).reshape(16)):
        b |= (int(t).bit_length() - 1 if t else 0) << (4 * i)
    return b


def py_unpack(b: int) -> np.ndarray:
    return np.array([(1 << ((b >> (4 * i)) & 0xF)) if (b >> (4 * i)) & 0xF else 0 for i in range(16)],
                    dtype=np.int64).reshape(4, 4)


# ----------------------------------------------------------------------------------------------
# C oracle wrappers (batches)
# ----------------------------------------------------------------------------------------------

def slide(state: np.ndarray, action: int) -> tuple[np.ndarray, int, bool]:
    s = np.ascontiguousarray(state, dtype=np.int64).reshape(16)
    out = np.zeros(16, dtype=np.int64)
    r = ctypes.c_int64(0)
    ch = lib().oracle_slide(_p(s, ctypes.c_int64), int(action), _p(out, ctypes.c_int64), ctypes.byref(r))
    return out.reshape(4, 4), int(r.value), bool(ch)


def legal_mask(state: np.ndarray) -> int:
    s = np.ascontiguousarray(state, dtype=np.int64).reshape(16)
    return int(lib().oracle_legal_mask(_p(s, ctypes.c_int64)))


def row_left(vec) -> tuple[np.ndarray, int]:
    v = np.ascontiguousarray(vec, dtype=np.int64)
    out = np.zeros(4, dtype=np.int64)
    r = ctypes.c_int64(0)
    lib().oracle_row_left(_p(v, ctypes.c_int64), _p(out, ctypes.c_int64), ctypes.byref(r))
    return out, int(r.value)


def pack(states: np.ndarray) -> np.ndarray:
    """[n,16] or [n,4,4] int64 tile values -> [n] uint64."""
    s = np.ascontiguousarray(states, dtype=np.int64).reshape(-1, 16)
    e = np.zeros_like(s)
    nzm = s != 0
    e[nzm] = np.log2(s[nzm]).astype(np.int64)
    shifts = (4 * np.arange(16, dtype=np.uint64))
    return (e.astype(np.uint64) << shifts).sum(axis=1, dtype=np.uint64)


def unpack(boards: np.ndarray) -> np.ndarray:
    """[n] uint64 -> [n,16] int64 tile values."""
    b = np.ascontiguousarray(boards, dtype=np.uint64).reshape(-1, 1)
    e = (b >> (4 * np.arange(16, dtype=np.uint64))) & np.uint64(0xF)
    return np.where(e > 0, np.left_shift(np.int64(1), e.astype(np.int64)), 0).astype(np.int64)


def exponents(boards: np.ndarray) -> np.ndarray:
    """[n] uint64 -> [n,16] float64 exponents (the network input, src/board.py:224-237)."""
    b = np.ascontiguousarray(boards, dtype=np.uint64).reshape(-1, 1)
    return ((b >> (4 * np.arange(16, dtype=np.uint64))) & np.uint64(0xF)).astype(np.float64)


def step_packed(boards, actions, seed=0, step=0, index_base=0, p4_threshold=0x1999999A,
                spawn_override=None, threads=1):
    """Batched oracle step on packed boards -> (next u64[n], reward i32[n], flags u8[n])."""
    b = np.ascontiguousarray(boards, dtype=np.uint64)
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    n = b.shape[0]
    nxt = np.zeros(n, dtype=np.uint64)
    rew = np.zeros(n, dtype=np.int32)
    flg = np.zeros(n, dtype=np.uint8)
    ov = None
    if spawn_override is not None:
        ov = np.ascontiguousarray(spawn_override, dtype=np.uint8)
    lib().oracle_step_packed_mt(_p(b, ctypes.c_uint64), _p(a, ctypes.c_uint8), _p(nxt, ctypes.c_uint64),
                                _p(rew, ctypes.c_int32), _p(flg, ctypes.c_uint8), n, int(seed), int(step),
                                int(index_base), int(p4_threshold),
                                _p(ov, ctypes.c_uint8) if ov is not None else None, int(threads))
    return nxt, rew, flg


def reset_packed(n, seed=0, step=0, index_base=0, p4_threshold=0x1999999A):
    b = np.zeros(n, dtype=np.uint64)
    lib().oracle_reset_packed(_p(b, ctypes.c_uint64), n, int(seed), int(step), int(index_base), int(p4_threshold))
    return b


def legal_mask_packed(boards):
    b = np.ascontiguousarray(boards, dtype=np.uint64)
    f = np.zeros(b.shape[0], dtype=np.uint8)
    lib().oracle_legal_mask_packed(_p(b, ctypes.c_uint64), _p(f, ctypes.c_uint8), b.shape[0])
    return f


def philox4x32_10(ctr, key):
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c, ctypes.c_uint32), _p(k, ctypes.c_uint32), _p(o, ctypes.c_uint32))
    return o


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def stream_boards(n: int, seed: int = 2048, index_base: int = 0, p_empty: float = 0.3, max_exp: int = 11) -> np.ndarray:
    """The library's own synthetic board stream (b2048_random_boards), bit for bit, on the CPU."""
    L = lib()
    L.oracle_stream_boards.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64,
                                       ctypes.c_uint32, ctypes.c_uint32]
    L.oracle_stream_boards.restype = None
    out = np.zeros(n, dtype=np.uint64)
    thr = max(0, min(0xFFFFFFFF, int(round(p_empty * 4294967296.0))))     # = b2048.env.p4_threshold
    L.oracle_stream_boards(_p(out, ctypes.c_uint64), n, seed & (2**64 - 1), index_base & (2**64 - 1), thr, max_exp)
    return out


def stream_actions(n: int, seed: int = 2050, step: int = 0, index_base: int = 0) -> np.ndarray:
    """The library's own synthetic action stream (b2048_random_actions), bit for bit, on the CPU."""
    L = lib()
    L.oracle_stream_actions.argtypes = [ctypes.POINTER(ctypes.c_uint8), ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64,
                                        ctypes.c_uint64]
    L.oracle_stream_actions.restype = None
    out = np.zeros(n, dtype=np.uint8)
    L.oracle_stream_actions(_p(out, ctypes.c_uint8), n, seed & (2**64 - 1), step & (2**64 - 1), index_base & (2**64 - 1))
    return out


def random_boards(n: int, seed: int, p_empty: float = 0.3, max_exp: int = 11) -> np.ndarray:
    """numpy-side synthetic boards with the distribution of SURVEY.md §8(d) (not bit-equal to the
    library's Philox generator; used where only the distribution matters)."""
    rng = np.random.default_rng(seed)
    e = rng.integers(1, max_exp + 1, size=(n, 16), dtype=np.uint64)
    e[rng.random((n, 16)) < p_empty] = 0
    return (e << (4 * np.arange(16, dtype=np.uint64))).sum(axis=1, dtype=np.uint64)
