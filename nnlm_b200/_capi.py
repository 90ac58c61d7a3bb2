"""ctypes binding of libnnlm_b200.so (the C ABI declared in include/nnlm_b200.h).

This is the same boundary the R `.Call` shim binds; Python stands in for R in this image because no R
toolchain is present. There is NO CPU fallback: if the shared library is missing, or no CUDA device is visible, every
compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnnlm_b200.so")

OK, E_ARG, E_NO_DEVICE, E_CUDA, E_NCCL, E_INTERRUPT, E_NOMEM = 0, -1, -2, -3, -4, -5, -6
PREC_AUTO, PREC_EXACT, PREC_FAST = 0, 1, 2
COMM_ID_BYTES = 128

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
INTERRUPT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class Options(C.Structure):
    _fields_ = [("precision", C.c_int32), ("device", C.c_int32), ("verbose_timing", C.c_int32),
                ("n_gpus", C.c_int32), ("comm", C.c_void_p), ("print", C.c_void_p), ("print_user", C.c_void_p),
                ("mkl_trace", C.c_int32), ("reserved1", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("upload_ms", C.c_double), ("loop_ms", C.c_double), ("download_ms", C.c_double),
                ("cross_ms", C.c_double), ("solve_ms", C.c_double), ("error_ms", C.c_double),
                ("launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("precision_used", C.c_int32), ("reserved0", C.c_int32),
                ("gram_ms", C.c_double), ("cross_launches", C.c_uint64), ("solve_launches", C.c_uint64),
                ("comm_ms", C.c_double), ("comm_bytes", C.c_uint64),
                ("host_setup_ms", C.c_double), ("host_loop_ms", C.c_double), ("host_finish_ms", C.c_double),
                ("host_total_ms", C.c_double), ("n_gpus_used", C.c_int32), ("mse_from_identity", C.c_int32),
                ("host_alloc_ms", C.c_double), ("host_teardown_ms", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_ if not f.startswith("reserved")}


class NnlmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nnlm_b200 error {code}: {msg}")
        self.code = code


class Interrupted(NnlmError):
    pass


_lib = None

# every symbol include/nnlm_b200.h declares (tests/test_abi.py checks the library exports each of them)
SYMBOLS = [
    "nnlm_abi_version", "nnlm_sizeof", "nnlm_device_count", "nnlm_nnmf", "nnlm_nnlm", "nnlm_update", "nnlm_na_mask", "nnlm_cross", "nnlm_na_corrections",
    "nnlm_session_create", "nnlm_session_create_synthetic", "nnlm_session_create_sharded", "nnlm_synth_block", "nnlm_session_set_factors", "nnlm_session_get_factors",
    "nnlm_session_run", "nnlm_session_error", "nnlm_session_mse", "nnlm_session_stats", "nnlm_session_reset_stats", "nnlm_session_destroy",
    "nnlm_synth_matrix",
    "nnlm_comm_unique_id", "nnlm_comm_init", "nnlm_comm_destroy",
]


def lib():
    """Load the CUDA library; fail loudly when it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `make -C nnlm_b200/csrc` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.nnlm_abi_version.restype = C.c_int
        L.nnlm_device_count.restype = C.c_int
        L.nnlm_device_count.argtypes = [C.c_char_p, C.c_size_t]
        for name in SYMBOLS:
            getattr(L, name)
        L.nnlm_session_destroy.restype = None
        L.nnlm_session_destroy.argtypes = [C.c_void_p]
        L.nnlm_comm_destroy.restype = None
        L.nnlm_comm_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def device_count():
    buf = C.create_string_buffer(256)
    n = lib().nnlm_device_count(buf, 256)
    return n, buf.value.decode()


def check(rc, err):
    if rc == OK:
        return
    msg = err.value.decode(errors="replace")
    if rc == E_INTERRUPT:
        raise Interrupted(rc, msg)
    raise NnlmError(rc, msg)


def d(a):
    return a.ctypes.data_as(_dp)


def i32(a):
    return None if a is None else a.ctypes.data_as(_ip)


def f64(x, copy=True):
    """Column-major float64 array (what R hands to .Call)."""
    if copy:
        return np.array(x, dtype=np.float64, order="F", copy=True)
    return np.asfortranarray(x, dtype=np.float64)


def lgl(x):
    """R logical matrix = int32, column-major."""
    return None if x is None else np.asfortranarray(np.asarray(x).astype(np.int32))


def vec3(x):
    return np.array(list(np.atleast_1d(np.asarray(x, dtype=np.float64))) + [0.0] * 3, dtype=np.float64)[:3].copy()
