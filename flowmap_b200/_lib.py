"""ctypes binding of csrc/libflowmap_b200.so (the C ABI declared in include/flowmap_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets an exception.  The library is built in-tree by ``flowmap_b200.build.build()``
(``nvcc -gencode arch=compute_100a,code=sm_100a``), see ``__graft_entry__.build``.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_float, c_int, c_size_t, c_void_p
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
SO_PATH = CSRC / "libflowmap_b200.so"

_lib = None

# name -> (restype, argtypes); must list every symbol of include/flowmap_b200.h
_P = c_void_p
SIGNATURES = {
    "fm_version": (c_int, []),
    "fm_last_error": (ctypes.c_char_p, []),
    "fm_launch_count": (ctypes.c_ulonglong, []),
    "fm_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "fm_workspace_reset": (c_int, [_P, c_int, c_int, c_int, c_int, _P]),
    "fm_unproject": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P]),
    "fm_unproject_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "fm_reproject": (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    "fm_procrustes_fwd": (c_int, [_P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "fm_procrustes_bwd": (c_int, [_P, _P, _P, _P, _P, c_int, _P, c_int, _P, _P, _P, _P, _P,
                                  c_int, c_int, c_int, c_int, _P]),
    "fm_mask_sum": (c_int, [_P, _P, _P, c_size_t, _P]),
    "fm_flow_loss_fwd_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_float, c_float,
                                     _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "fm_pose_chain": (c_int, [_P, _P, c_int, c_int, _P]),
    "fm_pose_chain_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    "fm_track_workspace_bytes": (c_size_t, [c_int, ctypes.c_longlong]),
    "fm_track_loss_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, ctypes.c_longlong,
                                  c_int, c_float, c_float, _P, _P, c_int, c_int, c_int, _P]),
    "fm_track_loss_bwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P, ctypes.c_longlong,
                                  c_int, c_float, c_float, _P, _P, _P, _P, _P, c_int, c_int, c_int,
                                  _P]),
    "fm_adam_step": (c_int, [_P, _P, _P, _P, c_size_t, c_double, c_double, c_double, c_double,
                             c_int, _P]),
}


class FlowmapLibraryError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not SO_PATH.exists():
            raise FlowmapLibraryError(
                f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  flowmap_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(str(SO_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().fm_last_error().decode("utf-8", "replace")
        raise FlowmapLibraryError(f"{what or 'flowmap_b200'} failed ({rc}): {msg}")
