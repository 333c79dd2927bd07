"""ctypes binding of libb2048.so (the C-ABI declared in include/b2048.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, every
compute entry point raises.  `lib()` alone (loading + symbol resolution) works on a CPU-only box so
that the ABI can be checked without a GPU.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2048_LIB") or os.path.join(_HERE, "libb2048.so")   # B2048_LIB: kernel-variant experiments only

c_void_p, c_int, c_int64, c_uint64, c_uint32 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                ctypes.c_uint64, ctypes.c_uint32)


class Ring(ctypes.Structure):
    """struct b2048_ring (include/b2048.h)."""
    _fields_ = [("s", c_void_p), ("s2", c_void_p), ("r", c_void_p), ("a", c_void_p), ("d", c_void_p),
                ("head_size", c_void_p), ("capacity", c_int64)]


class BoardResult(ctypes.Structure):
    """struct b2048_board_result (include/b2048.h)."""
    _fields_ = [("next", (ctypes.c_int64 * 16) * 4), ("reward", ctypes.c_int32 * 4), ("flags", ctypes.c_uint32),
                ("bad", ctypes.c_uint32)]


# symbol -> (restype, argtypes); must list every function include/b2048.h declares
SIGNATURES = {
    "b2048_init": (c_int, [c_int]),
    "b2048_shutdown": (c_int, [c_int]),
    "b2048_abi_version": (c_int, []),
    "b2048_error_string": (ctypes.c_char_p, [c_int]),
    "b2048_copy_row_lut_host": (c_int, [c_void_p]),
    "b2048_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_uint64, c_uint64,
                           c_uint64, c_uint32, c_void_p, c_void_p]),
    "b2048_step_all4": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_uint64, c_uint64,
                                c_uint64, c_uint32, c_void_p, c_void_p]),
    "b2048_legal_mask": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "b2048_reset": (c_int, [c_void_p, c_int64, c_uint64, c_uint64, c_uint64, c_uint32, c_void_p, c_void_p]),
    "b2048_spawn": (c_int, [c_void_p, c_int64, c_uint64, c_uint64, c_uint64, c_uint32, c_void_p, c_void_p]),
    "b2048_episode_end": (c_int, [c_void_p] * 11 + [c_int64, c_uint64, c_uint64, c_uint64, c_uint32, c_void_p]),
    "b2048_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2048_unpack_tiles": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "b2048_unpack_f64": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "b2048_random_boards": (c_int, [c_void_p, c_int64, c_uint64, c_uint64, c_uint32, c_uint32, c_void_p]),
    "b2048_random_actions": (c_int, [c_void_p, c_int64, c_uint64, c_uint64, c_uint64, c_void_p]),
    "b2048_step_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_uint64,
                                c_uint64, c_uint64, c_uint32, c_void_p, c_int]),
    "b2048_board_host": (c_int, [c_int, c_void_p, c_int, c_int, c_uint64, c_uint64, c_uint32, c_void_p, c_int]),
    "b2048_host_alloc": (c_int, [ctypes.POINTER(c_void_p), ctypes.c_size_t, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "b2048_host_free": (c_int, [c_void_p]),
    "b2048_bind_thread_near": (c_int, [c_int]),
    "replay_append": (c_int, [ctypes.POINTER(Ring), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int64, c_void_p]),
    "replay_sample": (c_int, [ctypes.POINTER(Ring), c_int64, c_uint64, c_uint64, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ddqn_target_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 ctypes.c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                 c_void_p]),
    "conv_patches_f64": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "conv_patches_grad_f64": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ddqn_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_double,
                               ctypes.c_double, ctypes.c_double, ctypes.c_double, c_void_p]),
    "p2p_get_ipc_handle": (c_int, [c_void_p, ctypes.c_char_p]),
    "p2p_open_ipc_handle": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_void_p)]),
    "p2p_allreduce_adam_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, c_void_p]),
    "dense_linear_forward_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "dense_linear_dgrad_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "dense_linear_dgrad_regroup_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "dense_linear_wgrad_scratch_elems": (c_int64, [c_int64, c_int, c_int]),
    "dense_linear_wgrad_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "egreedy_select": (c_int, [c_void_p, c_void_p, ctypes.c_double, c_uint64, c_uint64, c_uint64,
                               c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "layer_wgrad_small_scratch_elems": (c_int64, [c_int64, c_int, c_int]),
    "layer_wgrad_small_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "layer_wgrad64_scratch_elems": (c_int64, [c_int64, c_int]),
    "layer_wgrad64_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "qnet_conv_forward_train_f64": (c_int, [c_void_p] * 13 + [c_int64, c_void_p]),
    "qnet_conv_forward_update_f64": (c_int, [c_void_p] * 10 + [c_int64, c_void_p]),
    "conv1_wgrad_fused_scratch_elems": (c_int64, [c_int64]),
    "conv1_wgrad_fused_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "conv2_dgrad_conv1_wgrad_scratch_elems": (c_int64, [c_int64]),
    "conv2_dgrad_conv1_wgrad_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "qnet_conv_forward_f64": (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 8 + [c_void_p, c_int64, c_void_p]),
}

_lib = None
_lock = threading.Lock()
_inited: set[int] = set()


class B2048Error(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load libb2048.so and bind every symbol; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise B2048Error(
                        f"{LIB_PATH} is missing: build the CUDA extension first "
                        "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def error_string(code: int) -> str:
    return lib().b2048_error_string(int(code)).decode()


def check(code: int, what: str = "") -> None:
    if code != 0:
        raise B2048Error(f"{what or 'b2048 call'} failed: [{code}] {error_string(code)}")


def init(device: int = 0) -> None:
    """b2048_init(device) once per process and device; raises without a usable GPU."""
    if device in _inited:
        return
    check(lib().b2048_init(int(device)), f"b2048_init({device})")
    _inited.add(device)
