import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "reinforcement-learning-2048_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Test-infrastructure convenience only: build the extension and the oracle if a fresh checkout has
# no binaries yet (the product itself never builds or falls back: b2048._lib.lib() raises).
if not os.path.exists(os.path.join(PKG, "b2048", "libb2048.so")) or \
        not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
    import __graft_entry__
    __graft_entry__.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test started without a CUDA device")
    import b2048
    b2048._lib.init(0)
    return torch.device("cuda:0")
