"""ctypes binding of libp24_b200.so — the C ABI declared in ``include/p24.h``.

There is no CPU fallback: if the library is missing or does not load, importing the product
path raises.  ``P24_AUTOBUILD=1`` (default) rebuilds the library with nvcc when sources changed
and nvcc is available; on a box without nvcc the prebuilt in-tree ``.so`` is used as is.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None
ABI_VERSION = 4

c_f32p = C.c_void_p  # device pointers travel as integers
c_ptr = C.c_void_p

_PROTOS = {
    "p24_abi_version": (C.c_int, []),
    "p24_error_string": (C.c_char_p, [C.c_int]),
    "p24_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "p24_simota_loss_batch": (C.c_int, [c_ptr, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                        c_ptr, C.c_int64, C.c_int64, C.c_int,
                                        c_ptr, c_ptr, c_ptr,
                                        C.POINTER(C.c_int32), C.c_int,
                                        c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                        c_ptr, c_ptr, c_ptr,
                                        c_ptr, C.c_size_t, C.c_uint32,
                                        C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint32, c_ptr]),
    "p24_simota_loss_batch_raw": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int,
                                            c_ptr, C.c_int64, C.c_int64, C.c_int,
                                            c_ptr, c_ptr, c_ptr,
                                            C.POINTER(C.c_int32), C.c_int,
                                            c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                            c_ptr, c_ptr, c_ptr,
                                            c_ptr, C.c_size_t, C.c_uint32,
                                            C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint32, c_ptr]),
    "p24_comm_mailbox_bytes": (C.c_size_t, []),
    "p24_comm_alloc": (C.c_int, [C.POINTER(C.c_void_p)]),
    "p24_comm_free": (C.c_int, [c_ptr]),
    "p24_comm_export": (C.c_int, [c_ptr, C.c_char_p]),
    "p24_comm_import": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "p24_comm_close": (C.c_int, [c_ptr]),
    "p24_comm_finish": (C.c_int, [c_ptr, C.c_int, C.c_uint32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int,
                                  c_ptr]),
    "p24_workspace_init": (C.c_int, [c_ptr, C.c_size_t, c_ptr]),
    "p24_loss_finalize": (C.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "p24_circle_inter_fwd": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int64, c_ptr, c_ptr, c_ptr, C.c_int64, C.c_int,
                                       c_ptr, c_ptr, c_ptr]),
    "p24_iou_loss_fwd": (C.c_int, [c_ptr, C.c_int64, c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr]),
    "p24_iou_loss_bwd": (C.c_int, [c_ptr, C.c_int64, c_ptr, C.c_int64, c_ptr, C.c_int, c_ptr, c_ptr]),
    "p24_pair_iou": (C.c_int, [c_ptr, C.c_int64, C.c_int, c_ptr, C.c_int64, C.c_int, c_ptr, c_ptr]),
    "p24_loss_bwd": (C.c_int, [c_ptr, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                               c_ptr, C.c_int64, C.c_int64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "p24_loss_bwd_raw": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                   C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, C.c_int64, C.c_int64, c_ptr, c_ptr, c_ptr,
                                   c_ptr, c_ptr, c_ptr]),
    "p24_pack_labels": (C.c_int, [c_ptr, c_ptr, c_ptr, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr]),
    "p24_dynamic_k_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "p24_dynamic_k_matching": (C.c_int, [c_ptr, c_ptr, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                         c_ptr, C.c_size_t, c_ptr]),
    "p24_postprocess_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "p24_postprocess": (C.c_int, [c_ptr, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int,
                                  c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "p24_postprocess_raw": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_int,
                                      C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_int,
                                      c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.c_size_t, c_ptr]),
    "p24_read_status": (C.c_int, [c_ptr, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), c_ptr]),
    "p24_profile_enable": (C.c_int, [C.c_int]),
    "p24_profile_read": (C.c_int, [C.POINTER(C.c_float)]),
}


class P24Error(RuntimeError):
    pass


def load(path: str | None = None):
    """Load (building first if needed) the shared library and attach prototypes."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    if path is None:
        path = _build.lib_path()
        if os.environ.get("P24_AUTOBUILD", "1") == "1":
            try:
                if not _build.is_fresh():
                    _build.build()
            except Exception as exc:
                if not os.path.exists(path):
                    raise
                import warnings
                warnings.warn(f"p24: rebuilding libp24_b200.so failed ({exc}); loading the existing {path}, which does "
                              "not match the current sources", RuntimeWarning)
    if not os.path.exists(path):
        raise P24Error(f"{path} not found: build it with `python -m p24.build` (no CPU fallback exists)")
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.p24_abi_version() != ABI_VERSION:
        raise P24Error("libp24_b200 ABI version mismatch")
    _LIB = lib
    return lib


def check(code: int, what: str = ""):
    if code != 0:
        msg = load().p24_error_string(code).decode()
        raise P24Error(f"{what}: {msg} (code {code})")


def exported_names():
    return list(_PROTOS)
