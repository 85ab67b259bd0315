"""ctypes view of bench/baselines/libref_algos.so — the reference's ALGORITHMS restated in CUDA C++
("R" baseline of BASELINE.md §2).  Baseline only; never imported by the product package."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libref_algos.so")
        if not os.path.exists(path):
            raise ImportError("bench/baselines/libref_algos.so missing: run `make -C bench/baselines`")
        L = ctypes.CDLL(path)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.ref_sum_f32.argtypes = [vp, vp, vp, i64, i64, i64, vp]
        L.ref_cumsum_f32.argtypes = [vp, vp, vp, i64, vp]
        L.ref_cumsum_i64.argtypes = [vp, vp, vp, i64, vp]
        L.ref_bitonic_sort_u32.argtypes = [vp, i64, vp]
        for f in (L.ref_sum_f32, L.ref_cumsum_f32, L.ref_cumsum_i64, L.ref_bitonic_sort_u32):
            f.restype = ctypes.c_int
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(f"reference-algorithm baseline failed: cudaError/rc {rc}")


def sum_f32(A_ptr, R_ptr, partial_ptr, pre, red, post, stream):
    _chk(lib().ref_sum_f32(A_ptr, R_ptr, partial_ptr, pre, red, post, stream))


def cumsum(in_ptr, out_ptr, scratch_ptr, n, stream, i64=False):
    _chk((lib().ref_cumsum_i64 if i64 else lib().ref_cumsum_f32)(in_ptr, out_ptr, scratch_ptr, n, stream))


def bitonic_sort_u32(ptr, n, stream):
    _chk(lib().ref_bitonic_sort_u32(ptr, n, stream))
