"""ctypes binding of the C-ABI library `libprl_b200.so` (include/prl_b200.h).

There is NO CPU fallback: if the library is missing, or a call fails, this raises.
Build it with `python -c "import __graft_entry__ as g; g.build()"` or `make -C prime_rl_cse_b200/csrc`.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PRL_LIB_PATH") or os.path.join(_HERE, "libprl_b200.so")

PRL_ADV_KINDS = {"drgrpo": 0, "drgrpo-negclipped": 1, "rloo": 2, "opo": 3,
                 "grpo": 4}  # "grpo" (std-normalised) is an extension: not in the reference's registry (advantage.py:63-68)
PRL_RATIO = {"token": 0, "sequence": 1}
PRL_DT_BF16, PRL_DT_F32 = 0, 1

_lib = None

# name -> (restype, argtypes); mirrors include/prl_b200.h one to one
_SIGNATURES = {
    "prl_version": (c_int32, []),
    "prl_strerror": (ctypes.c_char_p, [c_int32]),
    "prl_last_error_detail": (ctypes.c_char_p, []),
    "prl_launch_count": (c_int64, [c_int32]),
    "prl_advantages_f32": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "prl_segments_from_position_ids": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p]),
    "prl_loss_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "prl_loss_packed_fwd": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int32, c_float, c_double,
         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "prl_scale_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "prl_pack_rollouts": (c_int32, [c_void_p] * 8 + [c_int64] + [c_void_p] * 6),
    "prl_masked_moments_workspace_bytes": (c_size_t, []),
    "prl_masked_moments_f32": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "prl_masked_keys_u32": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "prl_radix_select_hist": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "prl_radix_select_pick": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "prl_logits_stats_f32": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_int32, c_void_p, c_void_p, c_void_p]),
    "prl_logits_logsoftmax_bwd_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_int32, c_void_p, c_void_p]),
    "prl_lmhead_fwd_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int64]),
    "prl_lmhead_fwd": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64, c_int64, c_float, c_int32, c_int32,
         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "prl_lmhead_merge_partials": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "prl_lmhead_bwd_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int64, c_int64]),
    "prl_lmhead_bwd": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64, c_float, c_int32, c_int32,
         c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int64, c_int32, c_void_p, c_size_t, c_void_p],
    ),
    "prl_lmhead_bwd_scatter": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64, c_float, c_int32, c_int32,
         c_void_p, c_int32, c_int64, c_int32, c_void_p, c_int32, c_int32, c_int64, c_int32, c_void_p, c_size_t, c_void_p],
    ),
    "prl_lmhead_bwd_dw_scatter": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_float, c_int32, c_int32, c_void_p, c_int32,
         c_void_p, c_int32, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_size_t, c_void_p],
    ),
    "prl_ipc_alloc": (c_int32, [c_size_t, c_void_p, c_void_p]),
    "prl_ipc_free": (c_int32, [c_void_p]),
    "prl_ipc_open": (c_int32, [c_void_p, c_void_p]),
    "prl_ipc_close": (c_int32, [c_void_p]),
    "prl_lmhead_saved_logits_ld": (c_int64, [c_int64]),
    "prl_lmhead_fwd_save": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64, c_int64, c_float, c_int32,
         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p],
    ),
    "prl_lmhead_bwd_saved": (
        c_int32,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64, c_float, c_int32,
         c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p],
    ),
    "prl_lmhead_bwd_saved_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int64]),
    "prl_partition_positions": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "prl_gather_rows_bf16": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "prl_gather_tokens_f32": (c_int32, [c_void_p, c_int64] + [c_void_p] * 6 + [c_void_p]),
    "prl_scatter_tokens_f32": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_float,
                                         c_void_p, c_void_p, c_float, c_void_p]),
    "prl_scatter_rows_bf16": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "prl_set_knob": (c_int32, [ctypes.c_char_p, c_int64]),
    "prl_get_knob": (c_int32, [ctypes.c_char_p, c_void_p]),
    "prl_debug_schedule": (c_int32, [c_int32, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "prl_debug_schedule_scatter": (c_int32, [c_int64, c_int64, c_int64, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "prl_gemm_bf16_f32": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p]),
}


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the B200 CUDA library has not been built and there is no CPU fallback. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc)."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.prl_strerror(rc).decode()
    detail = lib.prl_last_error_detail().decode()
    text = f"{what}: {msg} ({detail})" if detail else f"{what}: {msg}"
    if rc in (-1, -2):
        raise ValueError(text)
    raise RuntimeError(text)


def ptr(t: torch.Tensor | None) -> c_void_p:
    return c_void_p(0) if t is None else c_void_p(t.data_ptr())


def stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors: torch.Tensor) -> None:
    """Every operand on the CURRENT CUDA device: the kernels are launched on torch's current stream of the current device
    (one process per GPU, as in the reference), so a tensor that lives on another GPU is refused instead of being addressed
    from the wrong device."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("prime_rl_cse_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: call torch.cuda.set_device "
                               "(one process per GPU) before using prime_rl_cse_b200")


def set_knob(name: str, value: int) -> None:
    """Tuning knob of the library (include/prl_b200.h "Tuning knobs"); the environment variable of the same name is only the
    initial value, read once."""
    check(load().prl_set_knob(name.encode(), int(value)), f"prl_set_knob({name})")


def get_knob(name: str) -> int:
    v = c_int64()
    check(load().prl_get_knob(name.encode(), ctypes.byref(v)), f"prl_get_knob({name})")
    return int(v.value)


def launch_count(reset: bool = False) -> int:
    return int(load().prl_launch_count(1 if reset else 0))
