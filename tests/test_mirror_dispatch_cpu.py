"""Walks the reference's own call chains through the host-side mirror WITHOUT a GPU:
which C-ABI entry point does each public call reach, and with which arguments?  The library is replaced by a
recording stub; device buffers are CPU torch tensors.

Reference chains (CUDACore/src/sorting.jl):
  partialsort(c, k)      -> partialsort!(copy(c), k; kw...)                 :1046-1048
  partialsort!(c, k; kw) -> partialsort!(c, k, alg; kw...)                  :1042-1044
  partialsort!(c,k,alg)  -> sort!(c, alg; lt, by, rev); copy(c[k])          :1015-1020   (alg POSITIONAL)
  sort!(c; alg, kw...)   -> sort!(c, alg; kw...)                            :1007-1009
  sortperm(c; dims, kw)  -> sortperm!(reshape(1:n), c; initialized=true...) :1067-1070   (hooked directly by the graft)
"""
import ctypes

import numpy as np
import pytest
import torch

import b200arrays as b
from b200arrays import dtypes as dt


class StubLib:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append((name, args))
            if name.endswith("workspace_bytes"):
                args[-1]._obj.value = 64          # ctypes.byref(c_size_t)
            return 0
        return fn


@pytest.fixture()
def stub(monkeypatch):
    lib = StubLib()
    import sys
    from b200arrays import _lib
    monkeypatch.setattr(_lib, "load", lambda: lib)
    for mod in (sys.modules[b.sort_.__module__], sys.modules[b.mapreducedim_.__module__]):
        monkeypatch.setattr(mod, "current_stream_handle", lambda dev: 0)
    return lib


def cpu_array(a):
    a = np.asarray(a)
    flat = np.ascontiguousarray(a.reshape(-1, order="F"))
    return b.CuArray(torch.from_numpy(flat.view(np.uint8).copy()), a.shape, dt.from_np(a.dtype))


def names(lib):
    return [c[0] for c in lib.calls]


def test_partialsort_range_reaches_positional_sort(stub):
    c = cpu_array(np.arange(10, 0, -1, dtype=np.float32))
    out = b.partialsort(c, range(2, 5), rev=True)            # -> partialsort!(copy(c), k) -> sort!(c, alg; lt, by, rev)
    assert names(stub) == ["b200_sort_dims_workspace_bytes", "b200_sort_keys_dims"]
    args = stub.calls[1][1]
    assert args[3] == dt.F32 and args[4:7] == (1, 10, 1) and args[7] == 1      # (pre, n, post), descending
    assert out.dims == (3,)                                                   # copy(c[k])
    np.testing.assert_array_equal(b.Array(c), np.arange(10, 0, -1, dtype=np.float32))   # partialsort copies first


def test_partialsort_bang_positional_alg(stub):
    c = cpu_array(np.arange(6, dtype=np.int32))
    b.partialsort_alg_(c, 3, b.BitonicSort)                   # the method Base's keyword wrapper calls
    assert names(stub) == ["b200_sort_dims_workspace_bytes", "b200_sort_keys_dims"]
    with pytest.raises(b.NotGraftedError):
        b.partialsort_alg_(c, 3, b.QuickSort)                 # reference's own QuickSort method (sorting.jl:1022-1040)


def test_sort_keyword_wrapper_and_fallbacks(stub):
    c = cpu_array(np.zeros((4, 5), dtype=np.uint32))
    b.sort_(c, dims=2, rev=False)
    assert stub.calls[1][1][4:7] == (4, 5, 1)
    with pytest.raises(b.NotGraftedError):
        b.sort_(c, dims=1, by=abs)
    with pytest.raises(b.NotGraftedError):
        b.sort_(c, dims=1, order="x")                         # unknown keyword: the reference method decides
    with pytest.raises(b.NotGraftedError):
        b.sort_(cpu_array(np.zeros(4, dtype=np.float32)), alg=b.QuickSort)


def test_partialsort_scalar_selects_without_sorting(stub):
    c = cpu_array(np.arange(8, dtype=np.float32))
    b.partialsort(c, 3)
    assert names(stub) == ["b200_select_kth_workspace_bytes", "b200_select_kth"]
    assert stub.calls[1][1][5] == 2                           # 0-based k


def test_sortperm_nd_uses_linear_indices(stub):
    A = cpu_array(np.zeros((3, 4), dtype=np.float32))
    ix = b.sortperm(A, dims=2)
    assert ix.eltype == dt.I64 and ix.dims == (3, 4)
    name, args = stub.calls[-1]
    assert name == "b200_sortperm_dims" and args[6:9] == (3, 4, 1) and args[10] == 1     # linear = 1
    v = cpu_array(np.zeros(5, dtype=np.float32))
    b.sortperm(v, rev=True)
    assert stub.calls[-1][1][9:11] == (1, 0)                  # descending, per-segment indices
    with pytest.raises(b.NotGraftedError):
        b.sortperm_(b.CuArray(torch.zeros(40, dtype=torch.uint8), (5,), dt.I64), v, initialized=True)
    with pytest.raises(b.NotGraftedError):
        b.sortperm(v, alg=b.BitonicSort)                      # bitonic_sort! has no `alg` keyword (sorting.jl:909)


def test_integer_and_with_gpuarrays_neutral_is_not_grafted(stub):
    """GPUArrays.neutral_element(&, T) is one(T): not neutral for integers (SURVEY §8 a1'), so `reduce(&, ints)` without
    an explicit init must stay on the reference path; Bool `all` and an explicit all-ones init are grafted."""
    assert b.neutral_element(b.and_, dt.I32) == 1 and b.true_neutral(b.and_, dt.I32) == -1
    assert b.true_neutral(b.and_, dt.U8) == 255 and b.true_neutral(b.and_, dt.BOOL) == 1
    A = cpu_array(np.arange(16, dtype=np.int32))
    with pytest.raises(b.NotGraftedError):
        b.mapreduce(b.identity, b.and_, A, dims=1)
    R = cpu_array(np.zeros(1, dtype=np.int32))
    b.mapreducedim_(b.identity, b.and_, R, A, init=np.int32(-1))
    assert names(stub)[-1] == "b200_reduce"
    flags = cpu_array(np.ones(16, dtype=np.bool_))
    R2 = cpu_array(np.zeros(1, dtype=np.bool_))
    b.mapreducedim_(b.identity, b.and_, R2, flags, init=True)
    assert names(stub)[-1] == "b200_reduce"
