"""b2048 — host-side Python over the C-ABI CUDA library (include/b2048.h).

Importing this package never touches the GPU; the first compute call initialises the library for
the tensor's device and raises if the CUDA extension or a GPU is missing (no CPU fallback).
"""
from . import _lib
from ._lib import B2048Error, LIB_PATH
from . import checkpoint, env, replay, ddqn, dist, p2p, player, qfused, qnet, rollout, train, trainer
from .env import (FLAG_BADSPAWN, FLAG_CHANGED, FLAG_DONE, FLAG_LEGAL, FLAG_OVERFLOW, P4_FIFTY_PERCENT,
                  P4_TEN_PERCENT, SPAWN_NONE, p4_threshold)
from .replay import ReplayDeque, ReplayRing

__all__ = ["_lib", "env", "replay", "ddqn", "dist", "rollout", "trainer", "ReplayRing", "ReplayDeque", "B2048Error", "LIB_PATH", "p4_threshold",
           "P4_TEN_PERCENT", "P4_FIFTY_PERCENT", "FLAG_LEGAL", "FLAG_DONE", "FLAG_CHANGED", "FLAG_OVERFLOW",
           "FLAG_BADSPAWN", "SPAWN_NONE"]
