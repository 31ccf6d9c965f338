"""ctypes binding of libvipant_b200.so (include/vipant_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails the caller
gets an exception.  The library is built in-tree by ``python -m vipant_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p

from . import build as _build

F32, BF16, F16, U8 = 0, 1, 2, 3
PREC_BF16_TC, PREC_FP32_SIMT = 0, 1

# name -> (restype, argtypes); mirrors include/vipant_b200.h one to one
_SIGNATURES = {
    "vpa_version": (c_int, []),
    "vpa_last_error_string": (c_char_p, []),
    "vpa_plan_query": (c_int, [c_int64, c_int64, c_int, c_int, c_int, POINTER(c_int)]),
    "vpa_profile_enable": (c_int, [c_int]),
    "vpa_profile_hold": (c_int, [c_int]),
    "vpa_launch_count": (ctypes.c_ulonglong, []),
    "vpa_profile_read": (c_int, [c_int, POINTER(c_float), POINTER(c_int)]),
    "vpa_normalize_cast": (c_int, [c_void_p, c_int, c_int64, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vpa_normalize_pair": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int64, c_int64, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vpa_infonce_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "vpa_infonce_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_int64,
                                c_void_p, c_float, c_void_p, c_void_p, c_size_t,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vpa_infonce_colsum_floats": (c_size_t, [c_int64]),
    "vpa_infonce_fwd_sweep": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_int64,
                                      c_void_p, c_float, c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "vpa_infonce_fwd_finish": (c_int, [c_int, c_int64, c_int64, c_int, c_int64, c_void_p, c_float, c_void_p, c_void_p, c_size_t,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vpa_infonce_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "vpa_infonce_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_int64,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64,
                                c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vpa_infonce_multi_state_bytes": (c_size_t, [c_int64, c_int, c_int, c_int, c_int]),
    "vpa_infonce_multi_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "vpa_infonce_multi_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                      c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "vpa_comm_load": (c_int, [c_char_p]),
    "vpa_comm_unique_id": (c_int, [c_void_p]),
    "vpa_comm_init": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "vpa_comm_destroy": (c_int, [c_void_p]),
    "vpa_sharded_state_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "vpa_infonce_fwd_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int64, c_int64, c_int,
                                        c_void_p, c_float, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "vpa_infonce_bwd_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int64, c_int64, c_int,
                                        c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vpa_p2p_create": (c_int, [c_int64, c_int, c_int, c_int, c_int, POINTER(c_void_p), c_void_p]),
    "vpa_p2p_connect": (c_int, [c_void_p, c_void_p]),
    "vpa_p2p_destroy": (c_int, [c_void_p]),
    "vpa_debug_relay_item": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int64, POINTER(c_int)]),
    "vpa_infonce_fwd_p2p": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int64, c_int64, c_int,
                                    c_void_p, c_float, c_int, c_void_p, POINTER(c_uint32), c_void_p]),
    "vpa_infonce_bwd_p2p": (c_int, [c_void_p, c_uint32, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int64, c_int64,
                                    c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vpa_sim_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "vpa_sim_rank_topk": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vpa_sim_fused_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "vpa_sim_rank_fused": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p, c_int, c_void_p, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vpa_encoder_tail": (c_int, [c_void_p, c_int, c_int64, c_int, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vpa_multilabel_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "vpa_multilabel_scores": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vpa_infonce_host_scratch_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "vpa_infonce_step_host": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_float, c_int,
                                      c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_LIB = None


class VipantB200Error(RuntimeError):
    pass


def library_path() -> str:
    """In-tree library; VIPANT_B200_LIB selects another build of the same C-ABI (A/B measurements only)."""
    return os.environ.get("VIPANT_B200_LIB") or _build.LIB_PATH


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise VipantB200Error(
                f"{path} not found: build it with `python -m vipant_b200.build` (needs nvcc). "
                "vipant_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().vpa_last_error_string()
        raise VipantB200Error(f"{what} failed with code {code}: {msg.decode() if msg else ''}")
