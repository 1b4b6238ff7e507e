"""ctypes binding of the C-ABI declared in include/tadbit_b200.h.

The shared library is built in-tree (``python -m tadbit_b200.build``) and loaded from
the package directory.  There is no CPU fallback: if the library is missing, or no
CUDA device is present when a compute entry point is called, the call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TADBIT_B200_LIB: explicit path of the library (a build variant under test)
LIB_PATH = os.environ.get("TADBIT_B200_LIB") or os.path.join(_HERE, "libtadbit_b200.so")

TB_OK, TB_ERR_ARG, TB_ERR_CUDA, TB_ERR_ALL_BAD, TB_ERR_NCCL, TB_ERR_STATE, TB_ERR_ZERO_DIV = range(7)
TB_MEM_HOST, TB_MEM_DEVICE = 0, 1
TB_CELLS_UPPER, TB_CELLS_FULL_ROWS = 0, 1


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("cis_scale", C.c_double), ("cis_alpha", C.c_double),
                ("trans_mean", C.c_double), ("vis_sigma", C.c_double), ("low_frac", C.c_double),
                ("low_scale", C.c_double), ("empty_frac", C.c_double)]


class B200Error(RuntimeError):
    """CUDA / NCCL / state failure inside libtadbit_b200."""


_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)
_store = C.c_void_p

# name -> (argtypes, restype); every symbol include/tadbit_b200.h declares
PROTOTYPES = {
    "tb_version": ([], C.c_int),
    "tb_last_error": ([], C.c_char_p),
    "tb_device_count": ([C.POINTER(C.c_int)], C.c_int),
    "tb_trim_memory": ([C.c_int], C.c_int),
    "tb_store_create_coo": ([C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_int32, _i64p, C.c_int64, C.c_int64, C.c_int,
                             C.POINTER(_store)], C.c_int),
    "tb_store_create_csr": ([C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_int32, _i64p, C.c_int64, C.c_int64, C.c_int,
                             C.POINTER(_store)], C.c_int),
    "tb_store_export_csr": ([_store, C.c_int, _i64p, _i64p, _i32p, _i32p], C.c_int),
    "tb_store_create_synthetic": ([C.c_int64, C.c_int32, _i64p, C.POINTER(SynthParams), C.c_int64,
                                   C.c_int64, C.c_int, C.POINTER(_store)], C.c_int),
    "tb_synthetic_row_counts": ([C.c_int64, C.c_int32, _i64p, C.POINTER(SynthParams), C.c_int64,
                                 C.c_int64, C.c_int, _i64p], C.c_int),
    "tb_store_destroy": ([_store], C.c_int),
    "tb_store_info": ([_store, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p], C.c_int),
    "tb_store_set_option": ([_store, C.c_char_p, C.c_char_p], C.c_int),
    "tb_store_layout": ([_store, _i64p], C.c_int),
    "tb_store_export_upper": ([_store, _i64p, _i32p, _i32p, _i32p], C.c_int),
    "tb_comm_unique_id": ([_u8p], C.c_int),
    "tb_comm_init": ([_store, C.c_int32, C.c_int32, _u8p], C.c_int),
    "tb_comm_create": ([C.c_int32, C.c_int32, _u8p, C.c_int, C.POINTER(C.c_void_p)], C.c_int),
    "tb_comm_destroy": ([C.c_void_p], C.c_int),
    "tb_store_attach": ([_store, C.c_void_p], C.c_int),
    "tb_row_stats": ([_store, _i64p, _i64p], C.c_int),
    "tb_col_sums_masked": ([_store, _u8p, _i64p], C.c_int),
    "tb_ice": ([_store, _u8p, C.c_int32, C.c_double, _f64p, _f64p, C.POINTER(C.c_int32)], C.c_int),
    "tb_store_create_batch": ([C.c_int32, C.POINTER(_store), C.c_int, C.POINTER(_store)], C.c_int),
    "tb_ice_batch": ([_store, _u8p, C.c_int32, C.c_double, _f64p, _f64p, C.POINTER(C.c_int32)],
                     C.c_int),
    "tb_probe_rowsum": ([_store, C.c_int32, _f64p], C.c_int),
    "tb_probe_k1": ([_store, C.c_int32, _f64p], C.c_int),
    "tb_row_work": ([_store, _i64p, _i64p], C.c_int),
    "tb_ice_last_timing": ([_store, _f64p, _f64p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
                           C.c_int),
    "tb_last_timing": ([_store, _f64p], C.c_int),
    "tb_norm_sum": ([_store, _f64p, _u8p, _f64p], C.c_int),
    "tb_diag_sums": ([_store, _u8p, C.c_int64, _i64p], C.c_int),
    "tb_pair_stats": ([_store, _f64p, _u8p, C.c_double, _f64p], C.c_int),
    "tb_pair_values": ([_store, _f64p, _u8p, C.c_int32, C.c_double, C.c_double, _i64p, _i32p, _i32p,
                        _f64p], C.c_int),
    "tb_window": ([_store, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _f64p, _f64p], C.c_int),
    "tb_region_cells": ([_store, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _u8p, _f64p, _f64p, C.c_int32,
                         _i64p, _i32p, _i32p, _i32p, _f64p, _f64p], C.c_int),
    "tb_write_abc_lines": ([C.c_char_p, C.c_int64, _i32p, _i32p, _i32p, _f64p], C.c_int),
    "tb_write_pair_lines": ([C.c_char_p, C.c_int64, _i64p, _i64p, _i64p, _f64p, C.c_char_p, C.c_char_p], C.c_int),
    "tb_pool_expected": ([C.c_int64, C.c_int64, _i64p, _i64p, C.c_double, _f64p, _i64p], C.c_int),
    "tb_pool_decay": ([C.c_int64, _f64p, _i64p, _i64p, C.c_double, _f64p, _u8p, _i32p], C.c_int),
    "tb_decay_sums": ([_store, _f64p, _u8p, _f64p, _i64p, _f64p], C.c_int),
    "tb_corr_build": ([_store, C.c_int64, C.c_int64, _f64p, _u8p, _f64p, C.c_int32, _i64p], C.c_int),
    "tb_corr_get": ([_store, _f64p], C.c_int),
    "tb_corr_set": ([_store, C.c_int64, _f64p], C.c_int),
    "tb_corr_eig": ([_store, C.c_int32, _f64p, _f64p, _i32p], C.c_int),
    "tb_corr_timing": ([_store, _f64p], C.c_int),
    "tb_tridiag_eig": ([C.c_int32, _f64p, _f64p, _f64p], C.c_int),
}

_lib = None


def lib():
    """The loaded library (loaded on first use)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "tadbit_b200: %s not found -- build it with `python -m tadbit_b200.build` "
                "(needs nvcc; there is no CPU fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (argtypes, restype) in PROTOTYPES.items():
            fn = getattr(handle, name)      # AttributeError if a declared symbol is missing
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def check(status):
    if status == TB_OK:
        return
    msg = lib().tb_last_error().decode("utf-8", "replace")
    if status == TB_ERR_ALL_BAD:
        raise ZeroDivisionError(msg)         # normalize_hic.py:207-208
    if status == TB_ERR_ZERO_DIV:
        raise ZeroDivisionError(msg)         # normalize_hic.py:148 / hic_data.py:371-374
    if status == TB_ERR_ARG:
        raise ValueError(msg)
    raise B200Error("tadbit_b200 status %d: %s" % (status, msg))


def device_count():
    n = C.c_int(0)
    st = lib().tb_device_count(C.byref(n))
    return n.value if st == TB_OK else 0


def trim_memory(device=0):
    """Return the device memory cached by the library's allocation pool to the driver."""
    check(lib().tb_trim_memory(device))


def ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def as_i32(a, what="array"):
    """Contiguous int32 view/copy; refuses values that int32 cannot hold (no silent wrap)."""
    a = np.asarray(a)
    if a.dtype != np.int32 and a.size:
        if a.dtype.kind == 'f':
            if not np.all(a == np.rint(a)):
                raise ValueError("%s must hold integers" % what)
        if a.min() < -2147483648 or a.max() > 2147483647:
            raise OverflowError("%s does not fit int32" % what)
    return np.ascontiguousarray(a, dtype=np.int32)


def bad_mask(size, bads):
    """bads dict (or None) -> uint8[N] mask."""
    mask = np.zeros(size, dtype=np.uint8)
    if bads:
        try:                                               # integer keys: no Python-level loop
            idx = np.fromiter(bads, dtype=np.int64, count=len(bads))
        except (TypeError, ValueError):
            idx = np.fromiter((int(k) for k in bads), dtype=np.int64, count=len(bads))
        idx = idx[(idx >= 0) & (idx < size)]
        mask[idx] = 1
    return mask
