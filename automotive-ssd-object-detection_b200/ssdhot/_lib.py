"""ctypes binding of libssdhot.so (include/ssdhot.h).  The library is the only compute path of
this package: if it is missing or was not built for the GPU at hand every op raises -- there is
no CPU or eager-PyTorch fallback."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSDHOT_LIB_PATH") or os.path.join(_HERE, "lib", "libssdhot.so")     # (the override serves A/B builds of the kernels)
ABI_VERSION = 2

_lock = threading.Lock()
_lib = None

vp, i32, i64, f32, f64, u64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_ulonglong

# name -> (restype, argtypes); mirrors include/ssdhot.h one to one
PROTOTYPES = {
    "ssdhot_abi_version": (i32, []),
    "ssdhot_status_string": (C.c_char_p, [i32]),
    "ssdhot_launch_count": (u64, []),
    "ssdhot_debug_timeline": (i32, [vp]),
    "ssdhot_debug_stream_probe": (i32, [vp, i32, i64, i32, vp, vp]),
    "ssdhot_prior_tables": (i32, [vp, i32, vp, vp, vp]),
    "ssdhot_prior_aux": (i32, [vp, i32, vp, vp]),
    "ssdhot_ssd300_layout_host": (i32, [vp, i32]),
    "ssdhot_match_encode": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, i32, i32, f32, f32, f32, f32, f32,
                                  vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "ssdhot_match_workspace_bytes": (u64, [i32, i32]),
    "ssdhot_compact_rows": (i32, [vp, vp, vp, i32, i32, vp, vp]),
    "ssdhot_loss_workspace_bytes": (u64, [i32, i32, i32]),
    "ssdhot_multibox_loss_fwd": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, i32, i32, f32, f32, vp, vp, i32,
                                       f32, f32, f32, f64, vp, vp, vp, vp, vp, vp, vp, vp]),
    "ssdhot_share_bytes": (u64, [i32, i32]),
    "ssdhot_share_reset": (i32, [vp, i32, vp]),
    "ssdhot_mined_ce_fwd": (i32, [vp, vp, vp, i32, i32, i32, f64, vp, vp, vp, vp]),
    "ssdhot_multibox_loss_bwd": (i32, [vp, i32, vp, vp, i32, f32, f32, vp, vp, i32, f32, f32, vp, vp, vp, vp, vp, vp]),
    "ssdhot_decode": (i32, [vp, vp, i32, f32, f32, vp, vp]),
    "ssdhot_pack_heads": (i32, [vp, i32, i32, vp, vp]),
    "ssdhot_nms_workspace_bytes": (u64, [i64]),
    "ssdhot_nms": (i32, [vp, vp, vp, i32, i64, i32, f32, i32, i32, vp, vp, vp, vp]),
    "ssdhot_predict_workspace_bytes": (u64, [i32, i32, i32]),
    "ssdhot_predict": (i32, [vp, i32, vp, vp, i32, i32, f32, f32, i32, i32, i32, f32, f32, f32, f32,
                             vp, vp, vp, vp, vp, vp, vp]),
    "ssdhot_predict_stages": (i32, [vp, i32, vp, vp, i32, i32, f32, f32, i32, i32, i32, f32, f32, f32, f32,
                                    vp, vp, vp, vp, vp, vp, i32, vp, vp]),
    "ssdhot_predict_heads": (i32, [vp, vp, vp, i32, i32, i32, f32, f32, i32, i32, i32, f32, f32, f32, f32,
                                   vp, vp, vp, vp, vp, vp, i32, vp, vp]),
    "ssdhot_multibox_loss_heads_fwd": (i32, [vp, vp, vp, i32, vp, vp, vp, i32, i32, f32, f32, vp, vp, i32, i32,
                                             f32, f32, f32, f64, vp, vp, vp, vp, vp, vp, vp, vp]),
    "ssdhot_multibox_loss_heads_bwd": (i32, [vp, vp, vp, i32, f32, f32, vp, vp, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp]),
    "ssdhot_peer_mailbox_bytes": (u64, []),
    "ssdhot_peer_alloc": (i32, [vp]),
    "ssdhot_peer_free": (i32, [vp]),
    "ssdhot_peer_export": (i32, [vp, vp]),
    "ssdhot_peer_open": (i32, [vp, vp]),
    "ssdhot_peer_close": (i32, [vp]),
    "ssdhot_allreduce_sums_peer": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "ssdhot_allreduce_partials_peer": (i32, [vp, i32, vp, vp, vp, i32, i32, i32, vp, vp]),
}


class SsdhotError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise SsdhotError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C automotive-ssd-object-detection_b200/csrc`). ssdhot has no CPU fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(handle, name)          # AttributeError here = header / library mismatch
                fn.restype, fn.argtypes = res, args
            if handle.ssdhot_abi_version() != ABI_VERSION:
                raise SsdhotError("libssdhot.so ABI version mismatch")
            _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().ssdhot_status_string(status).decode()
        raise SsdhotError(f"{what} failed: {msg} (status {status})")


def launch_count() -> int:
    return int(lib().ssdhot_launch_count())
