"""ctypes binding of the C ABI in include/bsr.h (libbsr_b200.so).

There is no CPU fallback: if the shared library is missing, or a compute entry point fails
(for instance because no CUDA device is visible), a BsrError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSR_LIB_VARIANT=tl selects the debug build with pipeline time stamps (csrc/Makefile, target `timeline`)
LIB_PATH = os.path.join(_HERE, "libbsr_b200_tl.so" if os.environ.get("BSR_LIB_VARIANT") == "tl" else "libbsr_b200.so")

BSR_OK = 0
BSR_ERR_INVALID = -1
BSR_ERR_CUDA = -2
BSR_ERR_OOM = -3
BSR_ERR_RANGE = -4
BSR_ERR_UNSUPPORTED = -5

BSR_F32 = 0
BSR_BF16 = 1
BSR_COSINE = 0
BSR_DOT = 1
BSR_MAX_TOP_K = 1024
BSR_COMM_ID_BYTES = 128


class BsrError(RuntimeError):
    """A C-ABI call returned a negative bsr_status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"bsr status {status}: {message}")
        self.status = status
        self.message = message


_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)
_store = C.c_void_p

# name -> (restype, argtypes); exactly the entry points include/bsr.h declares
SIGNATURES = {
    "bsr_version": (C.c_int, []),
    "bsr_last_error": (C.c_char_p, []),
    "bsr_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "bsr_store_create": (C.c_int, [C.c_int, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64,
                                   C.POINTER(_store)]),
    "bsr_store_destroy": (C.c_int, [_store]),
    "bsr_store_reset": (C.c_int, [_store]),
    "bsr_store_append": (C.c_int, [_store, C.c_void_p, C.c_uint64]),
    "bsr_store_append_device": (C.c_int, [_store, C.c_void_p, C.c_uint64]),
    "bsr_store_append_synthetic": (C.c_int, [_store, C.c_uint64, C.c_uint64, C.c_uint64]),
    "bsr_store_set_metric": (C.c_int, [_store, C.c_int]),
    "bsr_store_metric": (C.c_int, [_store, C.POINTER(C.c_int)]),
    "bsr_store_count": (C.c_int, [_store, _u64p]),
    "bsr_store_dim": (C.c_int, [_store, _u32p]),
    "bsr_store_row_base": (C.c_int, [_store, _u64p]),
    "bsr_store_device": (C.c_int, [_store, C.POINTER(C.c_int)]),
    "bsr_store_scan_bytes": (C.c_int, [_store, _u64p]),
    "bsr_store_get_many": (C.c_int, [_store, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bsr_search": (C.c_int, [_store, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                             C.c_void_p]),
    "bsr_candidate_slots": (C.c_uint32, [C.c_uint32]),
    "bsr_search_device": (C.c_int, [_store, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.c_void_p]),
    "bsr_merge_device": (C.c_int, [_store, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                   C.c_void_p, C.c_void_p]),
    "bsr_unpack_device": (C.c_int, [_store, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "bsr_merge_host": (C.c_int, [_store, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p,
                                 C.c_void_p, _u32p]),
    "bsr_comm_unique_id": (C.c_int, [C.c_void_p]),
    "bsr_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "bsr_comm_destroy": (C.c_int, [C.c_void_p]),
    "bsr_search_parallel": (C.c_int, [_store, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "bsr_search_parallel_device": (C.c_int, [_store, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                             C.c_void_p]),
    "bsr_comm_uses_mailbox": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "bsr_store_set_tensor_min_batch": (C.c_int, [_store, C.c_uint32]),
    "bsr_store_set_option": (C.c_int, [_store, C.c_char_p, C.c_int64]),
    "bsr_store_launch_count": (C.c_int, [_store, _u64p]),
    "bsr_store_last_scan_ms": (C.c_int, [_store, _f32p]),
    "bsr_store_set_profiling": (C.c_int, [_store, C.c_int]),
    "bsr_store_scan_profile": (C.c_int, [_store, C.POINTER(C.c_double), _u64p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libbsr_b200.so (built in-tree by `__graft_entry__.build()` / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BsrError(
                BSR_ERR_CUDA,
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `make -C better_search_rag_rust_b200/csrc`). There is no CPU "
                "fallback for the search path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != BSR_OK:
        msg = lib().bsr_last_error()
        raise BsrError(status, msg.decode("utf-8", "replace") if msg else "")
