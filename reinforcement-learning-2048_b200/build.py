#!/usr/bin/env python3
"""Build libb2048.so in-tree with nvcc for sm_100a only (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "b2048", "libb2048.so")
SOURCES = ["env_kernels.cu", "replay_kernels.cu", "ddqn_kernels.cu", "p2p_kernels.cu", "qnet_kernels.cu", "wgrad_kernels.cu", "dense_kernels.cu", "host_api.cu"]
HEADERS = ["b2048_common.cuh", os.path.join("..", "..", "include", "b2048.h")]


# Test-only copy of the library whose streaming env-step kernel splits a batch into launches of at most
# 40 000 octets (shipped: 2^30), so that tests/test_env_gpu.py can run the split path of launch_step.
OUT_SPLITTEST = os.path.join(HERE, "b2048", "libb2048_splittest.so")
SPLITTEST_DEFINES = ("B2048_STREAM_MAX_OCTS=40000",)


def needs_build() -> bool:
    if not os.path.exists(OUT) or not os.path.exists(OUT_SPLITTEST):
        return True
    t = min(os.path.getmtime(OUT), os.path.getmtime(OUT_SPLITTEST))
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    if not force and not defines and not needs_build():
        return OUT
    if not defines and out == OUT:          # the default build also makes the split-test copy, in parallel
        import threading
        err = []

        def _split():
            try:
                _compile(SPLITTEST_DEFINES, OUT_SPLITTEST, False)
            except Exception as e:          # noqa: BLE001 - re-raised below
                err.append(e)
        th = threading.Thread(target=_split)
        th.start()
        try:
            _compile((), OUT, verbose)
        finally:
            th.join()
        if err:
            raise err[0]
        return OUT
    return _compile(defines, out, verbose)


def _compile(defines, out: str, verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-o", out] + [f"-D{d}" for d in defines] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    extra = os.environ.get("B2048_NVCC_EXTRA")        # kernel-variant experiments only
    if extra:
        cmd += extra.split()
    subprocess.check_call(cmd, cwd=CSRC)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[5:] for a in sys.argv[1:] if a.startswith("-out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else OUT))
