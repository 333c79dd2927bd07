#!/usr/bin/env python3
"""Materialise the UNMODIFIED reference under oracle/_ref/ (test / baseline infrastructure).

The reference (ribal-aladeeb/reinforcement-learning-2048) is pure Python: there is nothing to compile,
"building" it means making its source tree available where the GPU box can see it.  /root/reference
exists only in the build container; oracle/_ref/ is git-ignored (the reference's sources never enter
this repository's history) but NOT gpurun-ignored, so it travels to the GPU box with the snapshot, like
the built .so files.  Used by
  * tests/test_reference_drivers_gpu.py: the reference's own driver scripts (src/player.py,
    src/double_dqn_conv.py, src/double_dqn_dense.py) executed unchanged against this repo's drop-in
    `board` / `dqn_lib` modules;
  * bench.py's reference-CPU figures (oracle/ref_bench.py): the reference's own numpy/torch path timed
    on the bench host's cores.
Nothing under reinforcement-learning-2048_b200/ imports or reads oracle/_ref/.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("B2048_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
WHAT = ("src", "tests", "LICENSE", "README.md", "requirements.txt")


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "src", "board.py"))


def make(force: bool = False) -> str | None:
    """Copy the reference tree; returns the destination, or None when the reference is not present
    (e.g. on the GPU box, where only the already materialised copy is used)."""
    if not os.path.isdir(os.path.join(REF_SRC, "src")):
        return DST if available() else None
    if available() and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for name in WHAT:
        src = os.path.join(REF_SRC, name)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(DST, name), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(src):
            shutil.copy2(src, os.path.join(DST, name))
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write(f"verbatim copy of {REF_SRC} made by oracle/make_ref.py; not part of this repository's history\n")
    return DST


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
