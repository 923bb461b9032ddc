"""ctypes binding of the C-ABI library (include/egm_b200.h).

There is deliberately no fallback: if `libegm_b200.so` cannot be loaded every operator of this
package raises. The library is looked up in-tree (`lib/libegm_b200.so`); when it is absent and
`nvcc` is on PATH it is built once (see build.py).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import threading
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libegm_b200.so")

PREC_FP32_SIMT, PREC_BF16X3, PREC_BF16 = 0, 1, 2
MHD_SYMMETRIC_GRAPH = 1      # EGM_MHD_SYMMETRIC_GRAPH
_PREC_NAMES = {"fp32_simt": PREC_FP32_SIMT, "fp32": PREC_BF16X3, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}

_lock = threading.Lock()
_lib = None


class EgmError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


# name -> (restype, argtypes); mirrors include/egm_b200.h one to one
_P, _I, _F, _Z, _LL = c_void_p, c_int, c_float, c_size_t, c_longlong
SIGNATURES = {
    "egm_version": (_I, []),
    "egm_last_error": (c_char_p, []),
    "egm_launch_count": (ctypes.c_ulonglong, []),
    "egm_prof_enable": (None, [_I]),
    "egm_prof_reset": (None, []),
    "egm_prof_count": (_I, []),
    "egm_prof_read": (_I, [_I, _P, _P, _P]),
    "egm_gpf_fused_ok": (_I, [_I, _I, _I, _I, _I]),
    "egm_gpf_raw_planes_ok": (_I, [_I, _I, _I, _I, _I]),
    "egm_gpf_ldr": (_LL, [_I]),
    "egm_gpf_fwd_workspace": (_Z, [_I, _I, _I, _I]),
    "egm_gpf_state_bytes": (_Z, [_I, _I, _I, _I]),
    "egm_align_fwd": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "egm_align_bwd": (_I, [_P, _P, _I, _I, _P, _P]),
    "egm_gpf_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "egm_gpf_bwd_workspace": (_Z, [_I, _I, _I, _I, _I, _I]),
    "egm_gpf_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P, _P, _I, _P, _Z, _P]),
    "egm_pool_state_bytes": (_Z, [_I, _I, _I, _I]),
    "egm_pool_fwd_workspace": (_Z, [_I, _I, _I, _I]),
    "egm_pool_fwd": (_I, [_P, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "egm_pool_bwd_workspace": (_Z, [_I, _I, _I, _I]),
    "egm_pool_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _I, _P, _Z, _P]),
    "egm_ns_state_bytes": (_Z, [_I, _I, _I, _I]),
    "egm_ns_fwd_workspace": (_Z, [_I, _I, _I, _I]),
    "egm_ns_fwd": (_I, [_P, _I, _I, _I, _F, _I, _P, _P, _P, _I, _P, _Z, _P]),
    "egm_ns_bwd_workspace": (_Z, [_I, _I, _I, _I]),
    "egm_ns_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P, _I, _P, _Z, _P]),
    "egm_mlr_state_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "egm_mlr_fwd_workspace": (_Z, [_I, _I, _I, _I, _I]),
    "egm_mlr_fwd": (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "egm_mlr_bwd_workspace": (_Z, [_I, _I, _I, _I, _I]),
    "egm_mlr_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _I, _P, _Z, _P]),
    "egm_mhd_state_bytes": (_Z, [_I, _I, _I, _I, _I]),
    "egm_mhd_fwd_workspace": (_Z, [_I, _I, _I, _I, _I]),
    "egm_mhd_fwd": (_I, [_P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "egm_mhd_bwd_workspace": (_Z, [_I, _I, _I, _I, _I]),
    "egm_mhd_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _I, _P, _Z, _P]),
    "egm_rowdot_bias": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "egm_linear_state_bytes": (_Z, [_I, _I, _I, _I]),
    "egm_linear_fwd_workspace": (_Z, [_I, _I, _I, _I]),
    "egm_linear_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _Z, _P]),
    "egm_linear_bwd_workspace": (_Z, [_I, _I, _I, _I]),
    "egm_linear_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _I, _P, _Z, _P]),
    "egm_triu_pack": (_I, [_P, _I, _I, _P, _P]),
    "egm_triu_unpack": (_I, [_P, _I, _I, _P, _P]),
    "egm_sketch_fwd": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "egm_sketch_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "egm_gram_workspace": (_Z, [_I, _I, _I, _I]),
    "egm_gram_fwd": (_I, [_P, _I, _I, _I, _I, _F, _P, _P, _I, _P, _Z, _P]),
    "egm_gram_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _I, _P, _Z, _P]),
    "egm_normalize_graph": (_I, [_P, _I, _I, _I, _F, _P, _P, _P]),
    "egm_batch_trace": (_I, [_P, _I, _I, _P, _P]),
    "egm_feature_tail_fwd": (_I, [_P, _I, _I, _P, _P, _P, _P, _I, _F, _F, _F, ctypes.c_ulonglong, _P, _P, _P, _P]),
    "egm_feature_tail_bwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, ctypes.c_ulonglong, _P, _P, _P, _P]),
    "egm_bmm_workspace": (_Z, [_I, _I, _I, _I, _I]),
    "egm_bmm": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _F, _P, _I, _P, _Z, _P]),
}


def lib_path() -> str:
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building first if needed and possible) the shared library; raise loudly otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if shutil.which(os.environ.get("NVCC", "nvcc")) is None:
                raise EgmError(
                    f"{LIB_PATH} is missing and nvcc is not available to build it; "
                    "this package has no CPU or PyTorch fallback (run __graft_entry__.build())")
            from . import build as _build
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export the header's symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().egm_last_error()
        raise EgmError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")


def precision_id(name) -> int:
    if isinstance(name, int):
        if name not in (0, 1, 2):
            raise ValueError(f"unknown precision id {name}")
        return name
    try:
        return _PREC_NAMES[str(name).lower()]
    except KeyError:
        raise ValueError(f"unknown precision mode {name!r}; expected one of {sorted(_PREC_NAMES)}") from None
