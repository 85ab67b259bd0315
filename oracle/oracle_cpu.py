"""ctypes view of oracle/liboracle_cpu.so (the C restatement of Base Julia's CPU path).
TEST / BASELINE INFRASTRUCTURE ONLY — see oracle_cpu.c's header."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle_cpu.so")
        if not os.path.exists(path):
            raise ImportError("oracle/liboracle_cpu.so missing: run `make -C oracle` (or __graft_entry__.build())")
        L = ctypes.CDLL(path)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.ref_sum_f32.restype = ctypes.c_float
        L.ref_sum_f32.argtypes = [vp, i64]
        L.ref_maximum_f32.restype = ctypes.c_float
        L.ref_maximum_f32.argtypes = [vp, i64]
        L.ref_maximum_bf16.restype = ctypes.c_float
        L.ref_maximum_bf16.argtypes = [vp, i64]
        L.ref_count_bool.restype = i64
        L.ref_count_bool.argtypes = [vp, i64]
        for name in ("ref_sum_dims1_f32", "ref_sum_dims2_f32", "ref_sum_dims1_bf16", "ref_sum_dims2_bf16"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, i64, i64, vp]
        for name in ("ref_cumsum_f32", "ref_cumsum_i64", "ref_sortperm_f32", "ref_sortperm_u32"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, vp, i64]
        for name in ("ref_sort_u32", "ref_sort_f32", "ref_sort_f32_cmp"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [vp, i64]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def sum_f32(a):
    a = np.ascontiguousarray(a.reshape(-1, order="F"), dtype=np.float32)
    return np.float32(lib().ref_sum_f32(_p(a), a.size))


def maximum_f32(a):
    a = np.ascontiguousarray(a.reshape(-1, order="F"), dtype=np.float32)
    return np.float32(lib().ref_maximum_f32(_p(a), a.size))


def maximum_bf16(a_u16):
    a = np.ascontiguousarray(a_u16.reshape(-1, order="F"))
    return np.float32(lib().ref_maximum_bf16(_p(a), a.size))


def count_bool(a):
    a = np.ascontiguousarray(a.reshape(-1, order="F")).view(np.uint8)
    return int(lib().ref_count_bool(_p(a), a.size))


def sum_dims(a, dims):
    """a: Fortran-ordered 2-D float32 or uint16 (bfloat16 bits)."""
    assert a.ndim == 2 and a.flags.f_contiguous
    m, n = a.shape
    suffix = "f32" if a.dtype == np.float32 else "bf16"
    out = np.empty(n if dims == 1 else m, dtype=a.dtype)
    getattr(lib(), f"ref_sum_dims{dims}_{suffix}")(_p(a), m, n, _p(out))
    return out.reshape((1, n) if dims == 1 else (m, 1))


def cumsum(v):
    v = np.ascontiguousarray(v)
    out = np.empty_like(v)
    (lib().ref_cumsum_f32 if v.dtype == np.float32 else lib().ref_cumsum_i64)(_p(v), _p(out), v.size)
    return out


def sort(v):
    v = np.ascontiguousarray(v).copy()
    (lib().ref_sort_f32 if v.dtype == np.float32 else lib().ref_sort_u32)(_p(v), v.size)
    return v


def sort_cmp(v):
    """Comparison-sort (qsort) variant of sort for Float32 — timing baseline only."""
    v = np.ascontiguousarray(v, dtype=np.float32).copy()
    lib().ref_sort_f32_cmp(_p(v), v.size)
    return v


def sortperm(v):
    v = np.ascontiguousarray(v)
    perm = np.empty(v.size, dtype=np.int64)
    (lib().ref_sortperm_f32 if v.dtype == np.float32 else lib().ref_sortperm_u32)(_p(v), _p(perm), v.size)
    return perm
