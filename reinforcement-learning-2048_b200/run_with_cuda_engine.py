#!/usr/bin/env python3
"""Run one of the reference's unchanged driver scripts on the CUDA engine.

    python run_with_cuda_engine.py /path/to/reference/src/double_dqn_conv.py   # or player.py, ...

The reference's scripts do `from board import Board2048` / `from dqn_lib import ...` and resolve them
through sys.path[0] (their own directory).  This launcher puts this package's directory first, so
`board` and `dqn_lib` resolve to the drop-in modules here while `configs`, `device`, `experiments`
and everything else still come from the reference tree.  The scripts ask for a job name on stdin
and write under <git root>/experiments/, exactly as they do on the reference implementation.
"""
import os
import runpy
import sys


def main() -> None:
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(script)]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
