"""CPU: the C-ABI library loads, exports every symbol include/b2048.h declares, its host-built
row table equals the reference's row move on all 65536 rows, and compute calls fail loudly
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import b2048
from b2048 import _lib, env

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "b2048.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"^\s*(?:const\s+)?(?:int64_t|int|char\s*\*|const char\*)\s+\**(\w+)\s*\(", text, flags=re.M)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_functions()
    assert len(names) >= 19, names
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b2048.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in b2048/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert _lib.lib().b2048_abi_version() == 2


def test_row_table_equals_reference_rows(golden_dir):
    g = np.load(os.path.join(golden_dir, "rows.npz"))
    lut = env.row_lut_host()
    res = lut & 0xFFFF
    rew4 = (lut >> 16) & 0x3FFF
    can_right = (lut >> 30) & 1
    ovf = lut >> 31
    want_exp = np.where(g["result"] > 0, np.log2(np.maximum(g["result"], 1)).astype(np.int64), 0)
    is_ovf = (g["result"] > 32768).any(axis=1)
    want = (want_exp[:, 0] | (want_exp[:, 1] << 4) | (want_exp[:, 2] << 8) | (want_exp[:, 3] << 12))
    ok = ~is_ovf
    assert np.array_equal(res[ok], want[ok])
    ok_r = ok.copy()
    ok_r[0xEEEE] = False                     # reward 65536 does not fit 14 bits: added on the global path
    assert np.array_equal(rew4[ok_r] * 4, g["reward"][ok_r])
    assert g["reward"][0xEEEE] == 65536 and rew4[0xEEEE] == 0
    # RIGHT bit: the reversed row changes under the reference's left move
    rows = np.arange(65536)
    rev = ((rows & 0xF) << 12) | ((rows & 0xF0) << 4) | ((rows & 0xF00) >> 4) | ((rows & 0xF000) >> 12)
    tiles_rev = np.stack([np.where((rev >> (4 * c)) & 0xF, 1 << ((rev >> (4 * c)) & 0xF), 0) for c in range(4)], axis=1)
    assert np.array_equal(can_right.astype(bool), (g["result"][rev] != tiles_rev).any(axis=1))
    assert np.array_equal(ovf.astype(bool), is_ovf)
    assert is_ovf.sum() > 0


def test_error_strings():
    assert "init" in _lib.error_string(-1)
    assert "fallback" in _lib.error_string(-3)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_compute_fails_loudly_without_gpu():
    with pytest.raises(b2048.B2048Error):
        _lib.init(0)
    with pytest.raises(b2048.B2048Error):
        env.step(torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.uint8))
    # raw ABI: uninitialised / no device -> negative error code, never a silent success
    z = np.zeros(4, dtype=np.uint64)
    p = ctypes.c_void_p(z.ctypes.data)
    rc = _lib.lib().b2048_legal_mask(p, p, 4, None)
    assert rc in (-1, -3)


def test_docs_state_the_current_entry_point_count():
    """README / DESIGN / INTEGRATION quote the number of C-ABI entry points; keep them honest."""
    n = len(declared_functions())
    for name in ("README.md", "DESIGN.md", "INTEGRATION.md"):
        text = open(os.path.join(ROOT, name)).read()
        found = [int(m) for m in re.findall(r"(\d+) (?:`extern \"C\"` )?entry points", text)]
        assert found and all(f == n for f in found), (name, found, n)
