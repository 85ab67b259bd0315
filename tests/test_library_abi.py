"""The C-ABI shared library loads and exports every symbol include/b200arrays.h declares
(no compute calls: this runs without a GPU), and the host mirror guards behave."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200arrays.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    import b200arrays as b
    lib = b._lib.load()
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200arrays.h but not exported"
        assert s in b._lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert set(b._lib.SIGNATURES) == set(syms)


def test_artifact_is_shippable_and_env_free():
    """Round-1 verdict: a 202 MB library with getenv knobs in product code is not a JLL artifact.  The product build
    compresses its fatbins and reads no environment variable (tuning knobs exist only in -DB200_TUNING builds)."""
    import b200arrays as b
    path = b._lib.LIB_PATH
    if os.environ.get("B200ARRAYS_LIB"):
        pytest.skip("a variant library was selected through B200ARRAYS_LIB")
    assert os.path.getsize(path) < 40 * 1024 * 1024, f"{path} is {os.path.getsize(path) >> 20} MB"
    blob = open(path, "rb").read()
    for knob in (b"B200_SORT_HYBRID", b"B200_MSD_MIN_LOG2N", b"B200_SCAN_CHAIN", b"B200_SELECT_FAST"):
        assert knob not in blob, f"{knob.decode()} is read by the product library"
    # the size queries of the workspaces the Julia layer asks for answer without a GPU
    n = ctypes.c_size_t(0)
    lib = b._lib.load()
    assert lib.b200_sort_workspace_bytes(6, -1, 1 << 30, 1, ctypes.byref(n)) == 0 and (4 << 30) < n.value < (5 << 30)
    assert lib.b200_select_kth_workspace_bytes(2, 1 << 30, ctypes.byref(n)) == 0 and n.value < (1 << 30)
    assert lib.b200_select_kth_workspace_bytes(2, 1 << 33, ctypes.byref(n)) == 2        # 32-bit bins: UNSUPPORTED


def test_host_only_entry_points():
    import b200arrays as b
    lib = b._lib.load()
    assert lib.b200_version() >= 100
    assert b"B200_OK" in lib.b200_status_string(0)
    assert b"UNSUPPORTED" in lib.b200_status_string(2)
    # argument validation happens before any CUDA call
    n = ctypes.c_size_t(123)
    assert lib.b200_reduce_workspace_bytes(2, 2, 0, 0, -1, 1, 1, ctypes.byref(n)) == 1      # INVALID_VALUE
    assert lib.b200_reduce_workspace_bytes(2, 4, 0, 0, 1, 8, 1, ctypes.byref(n)) == 2       # F32 -> I32 unsupported
    assert lib.b200_reduce_workspace_bytes(2, 2, 0, 0, 1, 0, 1, ctypes.byref(n)) == 0 and n.value == 0   # empty
    assert lib.b200_scan_workspace_bytes(2, 2, 6, 1, 8, 1, ctypes.byref(n)) == 2            # EXTREMA is reduce-only
    assert lib.b200_sort_workspace_bytes(99, -1, 10, 1, ctypes.byref(n)) == 2
    assert lib.b200_sort_workspace_bytes(2, -1, 0, 1, ctypes.byref(n)) == 0 and n.value == 0
    assert lib.b200_sort_keys(None, 0, None, 2, 1, 1, 0, None) == 0                          # n <= 1: nothing to do
    assert lib.b200_sortperm(None, 0, None, None, 2, 2, 5, 1, 0, 0, None) == 2               # Float32 index eltype


def test_mirror_guards_without_gpu():
    import numpy as np
    import b200arrays as b
    assert b.neutral_element(b.add_sum, b.dtypes.I64) == 0
    assert b.neutral_element(b.max_, b.dtypes.F32) == -np.inf
    assert b.neutral_element(b.min_, b.dtypes.I32) == np.iinfo(np.int32).max
    assert b.dtypes.widen_add(b.dtypes.I32) == b.dtypes.I64 and b.dtypes.widen_add(b.dtypes.BOOL) == b.dtypes.I64
    from b200arrays.mapreduce import _factor
    assert _factor((16384, 16384), (1, 16384)) == (1, 16384, 16384)
    assert _factor((16384, 16384), (16384, 1)) == (16384, 16384, 1)
    assert _factor((3, 4, 5), (3, 1, 5)) == (3, 4, 5)
    assert _factor((3, 4, 5), (1, 1, 5)) == (1, 12, 5)
    assert _factor((3, 4, 5), (3,)) == (3, 20, 1)              # R padded with singleton dims
    with pytest.raises(b.NotGraftedError):
        _factor((3, 4, 5), (1, 4, 1))                           # non-contiguous reduced dims
    with pytest.raises(b.DimensionMismatch):
        _factor((3, 4), (2, 4))
