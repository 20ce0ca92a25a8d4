"""ctypes binding of libhiecoattn_b200.so -- the C ABI declared in include/hiecoattn_b200.h.

There is no CPU implementation and no fallback: if the shared library is missing or a call fails, the
caller gets a RuntimeError.  The library is built in-tree by build.py (``python
visual-question-answering_b200/build.py`` or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhiecoattn_b200.so")

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/hiecoattn_b200.h one to one
SIGNATURES = {
    "hca_abi_version": (_i, []),
    "hca_last_error": (C.c_char_p, []),
    "hca_launch_count": (_i64, []),
    "hca_set_option": (_i, [C.c_char_p, C.c_char_p]),
    "hca_get_option": (C.c_char_p, [C.c_char_p]),
    "hca_pinned_alloc": (_i, [_sz, _i, C.POINTER(_p)]),
    "hca_pinned_free": (_i, [_p]),
    "hca_embedding_fwd": (_i, [_p, _p, _p, _i64, _i, _i64, _p]),
    "hca_embedding_bwd": (_i, [_p, _p, _p, _i64, _i, _i64, _p]),
    "hca_phrase_conv_pool_workspace": (_sz, [_i, _i, _i]),
    "hca_phrase_conv_pool_saved_bytes": (_sz, [_i, _i, _i]),
    "hca_phrase_conv_pool_fwd": (_i, [_p] * 11 + [_sz] + [_i, _i, _i, _p, _sz, _p]),
    "hca_phrase_conv_pool_bwd": (_i, [_p] * 9 + [_sz] + [_p] * 7 + [_i, _i, _i, _p, _sz, _p]),
    "hca_lstm_supported": (_i, [_i, _i, _i, _i]),
    "hca_lstm_saved_bytes": (_sz, [_i, _i, _i, _i]),
    "hca_lstm_workspace": (_sz, [_i, _i, _i, _i]),
    "hca_lstm_fwd": (_i, [_p] * 8 + [_sz] + [_i, _i, _i, _i, _p, _sz, _p]),
    "hca_lstm_bwd": (_i, [_p] * 4 + [_sz] + [_p] * 6 + [_i, _i, _i, _i, _p, _sz, _p]),
    "hca_coattn_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "hca_coattn_saved_bytes": (_sz, [_i, _i, _i, _i]),
    "hca_coattn_saved_attention": (_i, [_i, _i, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "hca_coattn_fwd": (_i, [_p, _i64, _i64, _i64] + [_p] * 14 + [_sz] + [_i, _i, _i, _i, _p, _sz, _p]),
    "hca_coattn_bwd": (_i, [_p] * 5 + [_sz] + [_p] * 12 + [_i, _i, _i, _i, _p, _sz, _p]),
    "hca_mlp_workspace": (_sz, [_i, _i, _i, _i]),
    "hca_mlp_saved_bytes": (_sz, [_i, _i, _i, _i]),
    "hca_mlp_fwd": (_i, [_p] * 12 + [_sz] + [_i, _i, _i, _i, _p, _sz, _p]),
    "hca_mlp_bwd": (_i, [_p] * 2 + [_sz] + [_p] * 9 + [_i, _i, _i, _i, _p, _sz, _p]),
    "hca_adam_step": (_i, [_p, _p, _p, _p, _i64, _p, _p, C.c_float, C.c_float, C.c_float, C.c_float, _p]),
    "hca_adam_prep": (_i, [_p, _p, C.c_float, C.c_float, C.c_float, _p]),
    "hca_dp_flags_bytes": (_sz, []),
    "hca_dp_reduce_adam": (_i, [_p, C.c_uint64, _sz, _sz, _sz, _i64, _i64, _i, _i, _i, _i, _p, _p, _p, C.c_float, C.c_float, C.c_float, _i, _p]),
    "hca_ce_loss_workspace": (_sz, [_i]),
    "hca_ce_loss": (_i, [_p, _i64, _p, _i, _i, C.c_float, _p, _p, _i64, _p, _sz, _p]),
    "hca_gemm_workspace": (_sz, [_i, _i, _i]),
    "hca_split_planes": (_i, [_p, _i64, _i, _p, _p]),
    "hca_proj_planes": (_i, [_p, _i64, _i, _p, _i, _p, _p, _p]),
    "hca_debug_gemm_timeline": (_i, [_p, _i]),
    "hca_debug_gemm_timeline_select": (_i, [_p, _i, _i]),
    "hca_debug_lstm_timeline": (_i, [_p]),
    "hca_debug_lstm_events": (_i, [_p, _p, _i]),
    "hca_wgrad_planes": (_i, [_p, _p, _i, _i, _i64, _p, _p]),
    "hca_gemm": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
}

ABI_VERSION = 6
_lib = None


def lib():
    """The loaded shared library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                "`python visual-question-answering_b200/build.py` (needs nvcc). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.hca_abi_version() != ABI_VERSION:
            raise RuntimeError("libhiecoattn_b200.so: ABI version mismatch, rebuild it")
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().hca_last_error().decode(errors="replace")
        raise RuntimeError(f"hiecoattn_b200 {what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().hca_launch_count())


def set_option(name: str, value: str):
    check(lib().hca_set_option(name.encode(), value.encode()), "set_option")


def get_option(name: str) -> str:
    return lib().hca_get_option(name.encode()).decode()
