"""ctypes binding of libb200arrays.so — the same exported symbols the Julia package
`lib/B200Arrays/src/libb200arrays.jl` ccalls (include/b200arrays.h).

Fails loudly when the shared library is missing: there is no fallback path.
"""
import ctypes
import os

from .errors import B200ArraysError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ARRAYS_LIB", os.path.join(_HERE, "libb200arrays.so"))

_c = ctypes
_i32, _i64, _sz, _vp = _c.c_int32, _c.c_int64, _c.c_size_t, _c.c_void_p

# name -> (restype, argtypes); every symbol include/b200arrays.h declares
SIGNATURES = {
    "b200_version": (_i32, []),
    "b200_status_string": (_c.c_char_p, [_i32]),
    "b200_last_cuda_error": (_i32, []),
    "b200_check_device": (_i32, []),
    "b200_launch_count": (_i64, []),
    "b200_reset_launch_count": (None, []),
    "b200_reduce_workspace_bytes": (_i32, [_i32, _i32, _i32, _i32, _i64, _i64, _i64, _c.POINTER(_sz)]),
    "b200_reduce": (_i32, [_vp, _sz, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _i32, _vp]),
    "b200_findminmax_workspace_bytes": (_i32, [_i32, _i64, _i64, _i64, _c.POINTER(_sz)]),
    "b200_findminmax": (_i32, [_vp, _sz, _vp, _vp, _vp, _i32, _i32, _i64, _i64, _i64, _vp]),
    "b200_scan_workspace_bytes": (_i32, [_i32, _i32, _i32, _i64, _i64, _i64, _c.POINTER(_sz)]),
    "b200_scan": (_i32, [_vp, _sz, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _vp, _vp]),
    "b200_scan_carry": (_i32, [_vp, _sz, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _vp, _vp, _i32, _vp]),
    "b200_sort_workspace_bytes": (_i32, [_i32, _i32, _i64, _i64, _c.POINTER(_sz)]),
    "b200_sort_keys": (_i32, [_vp, _sz, _vp, _i32, _i64, _i64, _i32, _vp]),
    "b200_sortperm": (_i32, [_vp, _sz, _vp, _vp, _i32, _i32, _i64, _i64, _i32, _i32, _vp]),
    "b200_sort_dims_workspace_bytes": (_i32, [_i32, _i32, _i64, _i64, _i64, _c.POINTER(_sz)]),
    "b200_sort_keys_dims": (_i32, [_vp, _sz, _vp, _i32, _i64, _i64, _i64, _i32, _vp]),
    "b200_sortperm_dims": (_i32, [_vp, _sz, _vp, _vp, _i32, _i32, _i64, _i64, _i64, _i32, _i32, _vp]),
    "b200_sort_keys_from": (_i32, [_vp, _sz, _vp, _vp, _i32, _i64, _i32, _vp, _vp]),
    "b200_mgpu_plan_bytes": (_i32, [_c.POINTER(_sz)]),
    "b200_mgpu_hist16": (_i32, [_vp, _i32, _i64, _i32, _i32, _vp, _vp]),
    "b200_mgpu_partition_pairs": (_i32, [_vp, _sz, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _i32, _vp]),
    "b200_mgpu_partition_index": (_i32, [_vp, _sz, _vp, _i32, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _vp]),
    "b200_mgpu_widen_index": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "b200_mgpu_plan": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "b200_mgpu_scatter_workspace_bytes": (_i32, [_i64, _c.POINTER(_sz)]),
    "b200_mgpu_scatter": (_i32, [_vp, _sz, _vp, _i32, _i64, _i32, _vp, _vp, _i32, _vp]),
    "b200_sort_pairs": (_i32, [_vp, _sz, _vp, _vp, _i32, _i32, _i64, _i32, _vp]),
    "b200_search_sorted": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp]),
    "b200_partition_workspace_bytes": (_i32, [_i64, _i32, _c.POINTER(_sz)]),
    "b200_partition": (_i32, [_vp, _sz, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _vp]),
    "b200_partition_count": (_i32, [_vp, _sz, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _vp]),
    "b200_partition_scatter": (_i32, [_vp, _sz, _vp, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "b200_select_kth_workspace_bytes": (_i32, [_i32, _i64, _c.POINTER(_sz)]),
    "b200_select_kth": (_i32, [_vp, _sz, _vp, _i32, _i64, _i64, _i32, _vp, _vp]),
    "b200_select_workspace_bytes": (_i32, [_i64, _c.POINTER(_sz)]),
    "b200_select_flagged": (_i32, [_vp, _sz, _vp, _vp, _i32, _vp, _i64, _vp]),
}

_lib = None


def load():
    """Load the library once; raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"libb200arrays.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C cuda.jl_b200/csrc -j`). There is no CPU or library fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the build lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    """Mirror of the @checked wrapper (lib/curand/src/libcurand.jl:10-32): status -> exception."""
    if status != 0:
        lib = load()
        text = lib.b200_status_string(status).decode()
        raise B200ArraysError(status, text, lib.b200_last_cuda_error() if status == 4 else 0)
