"""ctypes view of bench/baselines/libcub_baseline.so — CUB's device-wide sort / scan / reduce, the library
comparator ("L") of SURVEY §8(d).  Baseline only; never imported by the product package."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libcub_baseline.so")
        if not os.path.exists(path):
            raise ImportError("bench/baselines/libcub_baseline.so missing: run `make -C bench/baselines`")
        _lib = ctypes.CDLL(path)
    return _lib


def call(name, torch, dev, *ptr_args, n, stream):
    """Runs the two-phase CUB call `name` once and returns a closure that repeats it with the same temp storage."""
    fn = getattr(lib(), name)
    fn.restype = ctypes.c_int
    nbytes = ctypes.c_size_t(0)
    args = [ctypes.c_void_p(p) for p in ptr_args]
    rc = fn(None, ctypes.byref(nbytes), *args, ctypes.c_int64(n), ctypes.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"{name}: cudaError {rc} (size query)")
    temp = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)

    def run():
        rc = fn(ctypes.c_void_p(temp.data_ptr()), ctypes.byref(nbytes), *args, ctypes.c_int64(n), ctypes.c_void_p(stream))
        if rc != 0:
            raise RuntimeError(f"{name}: cudaError {rc}")
    return run
