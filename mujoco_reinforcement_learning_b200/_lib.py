"""ctypes binding of `libb200ppo.so` (the C ABI declared in `include/b200ppo.h`).

The product path fails loudly when the CUDA library is missing or no CUDA device is present: there is no
CPU fallback and no alternative backend.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import build as _build

_LOCK = threading.Lock()
_LIB = None

c_f32p = C.c_void_p
c_i64 = C.c_int64
c_i32 = C.c_int32
c_dbl = C.c_double
c_ptr = C.c_void_p

MAX_LAYERS = 8
ACT_TANH, ACT_RELU = 0, 1
PREC_FP32, PREC_BF16 = 0, 1
PROF_CLASSES = ("gather", "gemm_fwd", "loss", "gemm_dgrad", "gemm_wgrad", "adam", "allreduce", "other")


class MlpDesc(C.Structure):
    _fields_ = [
        ("n_layers", c_i32),
        ("in_dim", c_i32),
        ("dims", c_i32 * MAX_LAYERS),
        ("activation", c_i32),
        ("final_tanh", c_i32),
        ("out_scale", C.c_float),
    ]


class HParams(C.Structure):
    _fields_ = [
        ("learning_rate_actor", c_dbl),
        ("learning_rate_critic", c_dbl),
        ("beta1", c_dbl),
        ("beta2", c_dbl),
        ("adam_eps", c_dbl),
        ("clip_epsilon", c_dbl),
        ("entropy_eps", c_dbl),
    ]


# name -> (restype, argtypes); mirrors include/b200ppo.h one to one
SIGNATURES = {
    "b200ppo_version": (c_i32, []),
    "b200ppo_last_error": (C.c_char_p, []),
    "b200ppo_gae": (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_dbl, c_dbl, c_i32, c_i32, c_dbl,
                            c_ptr, c_ptr, c_ptr]),
    "b200ppo_gather_minibatch": (c_i32, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr,
                                         c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "b200ppo_gather_rows": (c_i32, [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "b200ppo_adam_step": (c_i32, [c_ptr, c_ptr, c_i32, c_i64, c_ptr, c_ptr, c_i64, c_dbl, c_dbl, c_dbl, c_dbl, c_i64,
                                  c_ptr]),
    "b200ppo_create": (c_i32, [C.POINTER(MlpDesc), C.POINTER(MlpDesc), c_i64, c_i32, C.POINTER(c_ptr)]),
    "b200ppo_destroy": (None, [c_ptr]),
    "b200ppo_param_count": (c_i64, [c_ptr]),
    "b200ppo_actor_param_count": (c_i64, [c_ptr]),
    "b200ppo_param_offset": (c_i64, [c_ptr, c_i32, c_i32, c_i32]),
    "b200ppo_saved_size": (c_i64, [c_ptr, c_i32, c_i64]),
    "b200ppo_mlp_forward": (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "b200ppo_mlp_backward": (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "b200ppo_policy_infer": (c_i32, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "b200ppo_evaluate": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "b200ppo_minibatch_grads": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, C.POINTER(HParams),
                                        c_ptr, c_ptr, c_ptr]),
    "b200ppo_train": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, C.POINTER(c_i64), c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                              c_ptr, c_i32, c_i64, c_i64, C.POINTER(HParams), c_ptr, c_ptr]),
    "b200ppo_update_host": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, C.POINTER(c_i64), c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                    c_ptr, c_ptr, c_i64, c_i64, c_dbl, c_dbl, c_i32, c_i32, c_dbl, c_ptr, c_i32, c_i64,
                                    c_i64, C.POINTER(HParams), c_ptr, c_ptr]),
    "b200ppo_update_host_begin": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_i32, c_i32]),
    "b200ppo_update_host_end": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, C.POINTER(c_i64), c_dbl, c_dbl, c_i32, c_i32, c_dbl, c_i64, c_i64,
                                        C.POINTER(HParams), c_ptr, c_i32, c_ptr]),
    "b200ppo_normalize_obs": (c_i32, [c_ptr, c_i32, c_i64, c_i32, c_i32, c_ptr, c_i32, c_i32, c_ptr, c_ptr]),
    "b200ppo_poll_error": (c_i32, [c_ptr, c_ptr]),
    "b200ppo_launch_count": (c_i64, []),
    "b200ppo_profile_begin": (c_i32, [c_ptr]),
    "b200ppo_profile_end": (c_i32, [c_ptr, C.POINTER(c_dbl), C.POINTER(c_i64)]),
    "b200ppo_debug_tc_gemm": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "b200ppo_set_fp32_terms": (c_i32, [c_ptr, c_i32]),
    "b200ppo_rollout_step": (c_i32, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "b200ppo_debug_gemm_split": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "b200ppo_debug_activations": (c_i32, [c_ptr, c_i32, c_i32, c_i32, c_i64, c_ptr, c_ptr]),
    "b200ppo_polyak_update": (c_i32, [c_ptr, c_ptr, c_i64, c_dbl, c_ptr]),
    "b200ppo_comm_unique_id": (c_i32, [c_ptr]),
    "b200ppo_comm_init": (c_i32, [c_ptr, c_ptr, c_i32, c_i32]),
    "b200ppo_comm_world": (c_i32, [c_ptr, C.POINTER(c_i32), C.POINTER(c_i32)]),
    "b200ppo_p2p_export": (c_i32, [c_ptr, c_ptr]),
    "b200ppo_p2p_import": (c_i32, [c_ptr, c_ptr, c_i32]),
    "b200ppo_p2p_enable": (c_i32, [c_ptr, c_i32]),
    "b200ppo_set_perm_layout": (c_i32, [c_ptr, c_i32]),
    "b200ppo_table_export": (c_i32, [c_ptr, c_i64, c_ptr]),
    "b200ppo_table_import": (c_i32, [c_ptr, c_ptr, c_i32, c_i64]),
    "b200ppo_table_fill": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr]),
    "b200ppo_table_replicate": (c_i32, [c_ptr, c_ptr]),
}


def library_path() -> str:
    return _build.LIB_PATH


def load() -> C.CDLL:
    """Load (building first if the in-tree .so is absent and nvcc is available).  Raises if that fails."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = library_path()
        if not os.path.exists(path):
            try:
                _build.build_library()
            except Exception as exc:  # no silent fallback
                raise RuntimeError(
                    f"libb200ppo.so is missing at {path} and could not be built ({exc}). "
                    "Run `python -m mujoco_reinforcement_learning_b200.build`; this package has no CPU fallback.") from exc
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


E_INDEX = -6


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b200ppo_last_error().decode("utf-8", "replace")
        if rc == E_INDEX:  # what the reference's `memory[idx]` raises (ppo.py:104)
            raise IndexError(f"libb200ppo {what}: {msg}")
        raise RuntimeError(f"libb200ppo {what} failed (code {rc}): {msg}")


def require_cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    """Inputs must be CUDA tensors of the right dtype — an error, never a fallback."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback (got device {t.device})")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t


def ptr(t) -> C.c_void_p:
    if t is None:
        return C.c_void_p(None)
    assert t.is_contiguous(), "internal: non-contiguous tensor passed to the C ABI"
    return C.c_void_p(t.data_ptr())


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
