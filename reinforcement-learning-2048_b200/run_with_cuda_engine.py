#!/usr/bin/env python3
"""Run one of the reference's unchanged driver scripts on the CUDA engine.

    python run_with_cuda_engine.py [options] /path/to/reference/src/double_dqn_conv.py   # or player.py, ...

The reference's scripts do `from board import Board2048` / `from dqn_lib import ...` and resolve them
through sys.path[0] (their own directory).  This launcher puts this package's directory first, so
`board` and `dqn_lib` resolve to the drop-in modules here while `configs`, `device`, `experiments`
and everything else still come from the reference tree.  The scripts ask for a job name on stdin
and write under <git root>/experiments/, exactly as they do on the reference implementation.

Options (for short smoke runs of scripts whose run lengths are hard-coded; the script files themselves
stay byte-for-byte the reference's):
    --set MODULE.ATTR=VALUE   import MODULE from the reference tree first and overwrite one attribute,
                              e.g. --set configs.double_dqn_conv.no_episodes=12  (VALUE is a Python literal)
    --limit-loops N           every `tqdm(...)` loop of the script stops after N items
                              (player.py plays 1000 + 1000 games inside tqdm loops)
"""
import ast
import importlib
import itertools
import os
import runpy
import sys


def main() -> None:
    args = sys.argv[1:]
    sets, limit = [], None
    while args and args[0].startswith("--"):
        if args[0] == "--set" and len(args) >= 2:
            sets.append(args[1])
            args = args[2:]
        elif args[0] == "--limit-loops" and len(args) >= 2:
            limit = int(args[1])
            args = args[2:]
        else:
            raise SystemExit(__doc__)
    if not args:
        raise SystemExit(__doc__)
    script = os.path.abspath(args[0])
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(script)]
    sys.argv = args
    for item in sets:
        target, _, value = item.partition("=")
        module, _, attr = target.rpartition(".")
        setattr(importlib.import_module(module), attr, ast.literal_eval(value))
    if limit is not None:
        import tqdm as _tqdm
        real = _tqdm.tqdm

        def short_tqdm(iterable=None, *a, **kw):
            return real(itertools.islice(iterable, limit), *a, **kw)
        _tqdm.tqdm = short_tqdm
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
