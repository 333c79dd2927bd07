"""Checkpoint / resume of the batched trainer state.

The reference only pickles the whole model (`Experiment.save`, src/experiments.py:128-148); the
optimizer, target network, replay buffer and RNG position are lost.  Here everything needed for a
bit-identical continuation is saved: online and target weights, Adam state, the replay ring incl.
its device counters, the boards of every running game with their per-game accumulators, and the
Philox step counter (the spawn stream is counter-based, so no generator state exists).
"""
from __future__ import annotations

import torch


def _ring_state(ring):
    return {"capacity": ring.capacity, "s": ring.s.cpu(), "s2": ring.s2.cpu(), "r": ring.r.cpu(), "a": ring.a.cpu(),
            "d": ring.d.cpu(), "head_size": ring.head_size.cpu()}


def _load_ring(ring, st):
    if st["capacity"] != ring.capacity:
        raise ValueError(f"replay capacity mismatch: checkpoint {st['capacity']} vs ring {ring.capacity}")
    for k in ("s", "s2", "r", "a", "d", "head_size"):
        getattr(ring, k).copy_(st[k])


_ENV_TENSORS = ("boards", "ep_score", "ep_moves", "ep_qsum", "totals", "qmean_sum", "max_tile_hist")


def save(path: str, updater=None, vector_env=None, ring=None, extra: dict | None = None) -> None:
    """`extra`: primitives, lists, dicts and tensors only (load() uses torch.load(weights_only=True))."""
    ck = {"format": "b2048-checkpoint-1", "extra": extra or {}}
    if updater is not None:
        ck["model"] = updater.model.state_dict()
        ck["target"] = updater.target.state_dict()
        ck["optimizer"] = updater.opt.state_dict()
        ck["updates"] = updater.updates
        ring = ring if ring is not None else updater.ring
    if ring is not None:
        ck["ring"] = _ring_state(ring)
    if vector_env is not None:
        ck["env"] = {k: getattr(vector_env, k).cpu() for k in _ENV_TENSORS}
        ck["env"].update(t=vector_env.t, seed=vector_env.seed, index_base=vector_env.index_base, n=vector_env.n)
    torch.save(ck, path)


def load(path: str, updater=None, vector_env=None, ring=None) -> dict:
    # weights_only: a checkpoint holds tensors, ints, floats, strings, lists and dicts only, so nothing
    # from an untrusted file is ever unpickled into code (`extra` must stay primitives / tensors)
    ck = torch.load(path, map_location="cpu", weights_only=True)
    if ck.get("format") != "b2048-checkpoint-1":
        raise ValueError("not a b2048 checkpoint")
    if updater is not None:
        with torch.no_grad():                     # in place: the parameters are views of one flat buffer
            for k, v in updater.model.state_dict().items():
                v.copy_(ck["model"][k])
        updater.target.load_state_dict(ck["target"])
        updater.opt.load_state_dict(ck["optimizer"])
        updater.updates = ck["updates"]
        updater.graph = None                      # re-capture with the restored optimizer state
        ring = ring if ring is not None else updater.ring
    if ring is not None and "ring" in ck:
        _load_ring(ring, ck["ring"])
    if vector_env is not None:
        e = ck["env"]
        if e["n"] != vector_env.n:
            raise ValueError("number of environments differs from the checkpoint")
        for k in _ENV_TENSORS:
            getattr(vector_env, k).copy_(e[k])
        vector_env.t, vector_env.seed, vector_env.index_base = e["t"], e["seed"], e["index_base"]
    return ck["extra"]
