"""ctypes binding of libqldpc_b200.so (C ABI: include/qldpc_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded this module raises, and so does
every decoder entry point of the package.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QLDPC_B200_LIB") or os.path.join(HERE, "libqldpc_b200.so")     # override: experiments only

NG, BF, MS, BP = 0, 1, 2, 3
DEC_TYPES = {"NG": NG, "BF": BF, "MS": MS, "BP": BP}
NUM_COUNTERS = 10
CNT_FAIL_X, CNT_FAIL_Z, CNT_EXACT, CNT_DEGEN, CNT_ITERS_X, CNT_ITERS_Z, CNT_SHOTS = range(7)
CNT_TRUE_DEGEN, CNT_LOGICAL, CNT_FAIL_ANY = 7, 8, 9      # README.md:15-22 classes (extension, see include/qldpc_b200.h)
ABI_VERSION = 2

# every symbol include/qldpc_b200.h declares
EXPORTS = ["qldpc_abi_version", "qldpc_last_error", "qldpc_words", "qldpc_plan_create", "qldpc_plan_destroy",
           "qldpc_plan_info", "qldpc_decode", "qldpc_decode_host", "qldpc_osd", "qldpc_classify", "qldpc_sample",
           "qldpc_launch_count", "qldpc_plan_set_logicals", "qldpc_plan_work", "qldpc_simulate_host"]


class Graph(ctypes.Structure):
    _fields_ = [("m", ctypes.c_int32), ("n", ctypes.c_int32), ("nnz", ctypes.c_int32),
                ("row_ptr", ctypes.c_void_p), ("col_idx", ctypes.c_void_p),
                ("n_layers", ctypes.c_int32), ("layer_ptr", ctypes.c_void_p), ("layer_chk", ctypes.c_void_p)]


class Opts(ctypes.Structure):
    _fields_ = [("dec_type", ctypes.c_int32), ("max_iter", ctypes.c_int32), ("prior_llr", ctypes.c_double),
                ("beta", ctypes.c_double), ("eps", ctypes.c_double), ("osd_order", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class QldpcError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QldpcError(f"{LIB_PATH} not built: run `python -m qldpcsim_b200.build` (needs nvcc); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64, f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
    L.qldpc_abi_version.restype = ctypes.c_int
    L.qldpc_last_error.restype = ctypes.c_char_p
    L.qldpc_words.argtypes = [i32]
    L.qldpc_plan_create.argtypes = [ctypes.POINTER(Graph), ctypes.POINTER(Opts), ctypes.c_int, ctypes.POINTER(vp)]
    L.qldpc_plan_destroy.argtypes = [vp]
    L.qldpc_plan_info.argtypes = [vp, ctypes.c_int]
    L.qldpc_plan_info.restype = i64
    L.qldpc_decode.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    L.qldpc_decode_host.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.qldpc_osd.argtypes = [vp, vp, vp, vp, vp, i64, i32, vp]
    L.qldpc_classify.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp]
    L.qldpc_sample.argtypes = [vp, vp, f64, u64, i64, i64, vp, vp, vp, vp, vp]
    L.qldpc_launch_count.restype = i64
    L.qldpc_plan_set_logicals.argtypes = [vp, vp, i32]
    L.qldpc_simulate_host.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp]
    L.qldpc_plan_work.argtypes = [vp, ctypes.c_int]
    L.qldpc_plan_work.restype = i64
    if L.qldpc_abi_version() != ABI_VERSION:
        raise QldpcError("libqldpc_b200.so ABI version mismatch; rebuild")
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        msg = lib().qldpc_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(msg)
        raise QldpcError(f"libqldpc_b200 error {rc}: {msg}")
