"""Host-side mirror of the Base hooks in CUDACore/src/sorting.jl:993-1070.

`sort_` is `Base.sort!(c::AnyCuArray; alg, lt, by, rev, dims)`, `sortperm_` is
`Base.sortperm!(ix, A; initialized, ...)`, etc.  Grafted for `lt=isless`, `by=identity`
(kwargs do not dispatch, so each hook guards at run time); anything else
raises NotGraftedError where Julia would `invoke` the reference's bitonic method.
Device work: libb200arrays' onesweep radix sort (b200_sort_keys / b200_sortperm).
"""
import ctypes

import numpy as np

from . import _lib
from . import dtypes as dt
from .cuarray import CuArray, Workspace, current_stream_handle
from .errors import ArgumentError, NotGraftedError
from .mapreduce import identity


class _Token:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


isless = _Token("isless")
BitonicSort = _Token("CUDA.BitonicSort")      # sorting.jl:984-990: default alg token; the graft answers for it
QuickSort = _Token("CUDA.QuickSort")          # not exercised on cc >= 9.0 (test/core/sorting.jl:256,282): not grafted
SORTPERM_INDEX = 100


_SORT_KW = ("lt", "by", "rev", "dims")


def _graftable(kwargs):
    """lib/B200Arrays/src/sorting.jl `graftable_sort`: only known keywords, default lt / by, Bool rev."""
    for k in kwargs:
        if k not in _SORT_KW:
            raise NotGraftedError(f"keyword {k!r}: the reference method decides (MethodError for unknown keywords)")
    if kwargs.get("lt", isless) is not isless or kwargs.get("by", identity) is not identity:
        raise NotGraftedError("sort with a custom lt/by stays on the reference's bitonic kernels")
    if not isinstance(kwargs.get("rev", False), bool):
        raise NotGraftedError("rev must be a Bool")


def _factor(c, dims):
    """(pre, n, post) of sorting along `dims` (1-based); vectors: dims in (None, 1)."""
    if c.ndims <= 1:
        if dims not in (None, 1):
            raise ArgumentError("dimension out of range")
        return 1, c.length, 1
    if dims is None:
        raise NotGraftedError("sort! of a multidimensional array needs dims (Base raises UndefKeywordError)")
    if not 1 <= dims <= c.ndims:
        raise ArgumentError("dimension out of range")
    pre = int(np.prod(c.dims[:dims - 1], dtype=np.int64))
    n = int(c.dims[dims - 1])
    post = int(np.prod(c.dims[dims:], dtype=np.int64))
    return pre, n, post


def sort_alg_(c, alg, **kwargs):
    """Base.sort!(c::CuArray{T}, alg::BitonicSortAlg; kwargs...) — THE grafted method (lib/B200Arrays/src/sorting.jl):
    the reference's `sort!(c; alg, kw...)` (sorting.jl:1007-1009) and `partialsort!(c, k, alg; ...)` (sorting.jl:1015-1020)
    both pass `alg` positionally and land here; QuickSort dispatches to the reference's own method (sorting.jl:993)."""
    if alg is not BitonicSort:
        raise NotGraftedError("alg=QuickSort stays on the reference path (sorting.jl:993-1001)")
    _graftable(kwargs)
    pre, n, post = _factor(c, kwargs.get("dims"))
    rev = kwargs.get("rev", False)
    if n <= 1 or pre * post == 0:
        return c
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.b200_sort_dims_workspace_bytes(c.eltype, 0, pre, n, post, ctypes.byref(nbytes)))
    ws = Workspace(nbytes.value, c.buf.device)
    _lib.check(lib.b200_sort_keys_dims(ws.ptr, ws.nbytes, c.pointer, c.eltype, pre, n, post, 1 if rev else 0,
                                       current_stream_handle(c.buf.device)))
    return c


def sort_(c, alg=BitonicSort, **kwargs):
    """Base.sort!(c::AnyCuArray; alg=BitonicSort, kwargs...) = sort!(c, alg; kwargs...) (sorting.jl:1007-1009)."""
    return sort_alg_(c, alg, **kwargs)


def sort(c, **kw):
    """Base.sort(c::AnyCuArray; kw...) = sort!(copy(c); kw...) (sorting.jl:1011-1013)."""
    return sort_(c.copy(), **kw)


def _index_k(c, k):
    """copy(c[k]) for a 1-based integer or range k."""
    if isinstance(k, (int, np.integer)):
        if not 1 <= k <= c.length:
            raise IndexError(f"BoundsError: attempt to access {c.length}-element CuArray at index [{k}]")
        sz = dt.SIZES[c.eltype]
        return CuArray(c.buf[(k - 1) * sz:k * sz].clone(), (1,), c.eltype).to_host()[0]   # @allowscalar
    k = range(k.start, k.stop, k.step or 1) if isinstance(k, slice) else k
    if len(k) and (k.step != 1 or k[0] < 1 or k[-1] > c.length):
        raise NotGraftedError("only unit-stride in-bounds ranges are grafted for partialsort")
    sz = dt.SIZES[c.eltype]
    lo = (k[0] - 1) if len(k) else 0
    return CuArray(c.buf[lo * sz:(lo + len(k)) * sz].clone(), (len(k),), c.eltype)


def partialsort_alg_(c, k, alg, lt=isless, by=identity, rev=False):
    """Base.partialsort!(c::AnyCuVector, k, alg::BitonicSortAlg; lt, by, rev) = sort!(c, alg; lt, by, rev); copy(c[k])
    (sorting.jl:1015-1020) — reference code, reaches the grafted positional sort! method."""
    if c.ndims != 1:
        raise NotGraftedError("partialsort! is defined for vectors")
    sort_alg_(c, alg, lt=lt, by=by, rev=rev)
    return _index_k(c, k)


def partialsort_(c, k, alg=BitonicSort, **kwargs):
    """Base.partialsort!(c, k; alg=BitonicSort, kwargs...) = partialsort!(c, k, alg; kwargs...) (sorting.jl:1042-1044)."""
    return partialsort_alg_(c, k, alg, **kwargs)


def partialsort(c, k, alg=BitonicSort, **kwargs):
    """Base.partialsort(c::AnyCuArray, k; kw...) = partialsort!(copy(c), k; kw...) (sorting.jl:1046-1048).  For a scalar k nothing needs to be
    sorted or copied: b200_select_kth (radix select, c untouched) finds the element a
    stable sort would put at position k.  Ranges go through the full sort like the reference."""
    if not isinstance(k, (int, np.integer)):
        return partialsort_(c.copy(), k, alg=alg, **kwargs)
    if c.ndims != 1:
        raise NotGraftedError("partialsort is defined for vectors")
    if alg is not BitonicSort or "dims" in kwargs:
        raise NotGraftedError("alg=QuickSort / dims stay on the reference path")
    _graftable(kwargs)
    rev = kwargs.get("rev", False)
    if not 1 <= k <= c.length:
        raise IndexError(f"BoundsError: attempt to access {c.length}-element CuArray at index [{k}]")
    lib = _lib.load()
    dev = c.buf.device
    out = CuArray.undef(c.eltype, (1,), dev)
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.b200_select_kth_workspace_bytes(c.eltype, c.length, ctypes.byref(nbytes)))
    ws = Workspace(nbytes.value, dev)
    _lib.check(lib.b200_select_kth(ws.ptr, ws.nbytes, c.pointer, c.eltype, c.length, int(k) - 1, 1 if rev else 0,
                                   out.pointer, current_stream_handle(dev)))
    return out.to_host()[0]                                                    # @allowscalar, like the reference


def sortperm_(ix, A, initialized=False, **kwargs):
    """Base.sortperm!(ix::AnyCuArray, A::AnyCuArray; initialized=false, kw...) (sorting.jl:1051-1061).  The reference
    forwards kw to bitonic_sort!(; by, lt, rev, dims), which has no `alg` keyword."""
    if ix.dims != A.dims:
        def axes(d):
            return "(" + ", ".join(f"Base.OneTo({s})" for s in d) + ("," if len(d) == 1 else "") + ")"
        raise ArgumentError("index array must have the same size/axes as the source array, "
                            f"{axes(ix.dims)} != {axes(A.dims)}")              # sorting.jl:1052-1054
    if A.length == 0:
        return ix                                                              # sorting.jl:1055
    _graftable(kwargs)
    rev, dims = kwargs.get("rev", False), kwargs.get("dims")
    if initialized:
        # a caller-supplied starting permutation breaks ties by index VALUE (sorting.jl:642-643)
        raise NotGraftedError("sortperm!(...; initialized=true) stays on the reference path")
    if ix.eltype not in (dt.I32, dt.I64, dt.U32, dt.U64):
        raise NotGraftedError("index eltype must be a 32/64-bit integer")
    pre, n, post = _factor(A, dims)
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.b200_sort_dims_workspace_bytes(A.eltype, 1, pre, n, post, ctypes.byref(nbytes)))
    ws = Workspace(nbytes.value, A.buf.device)
    _lib.check(lib.b200_sortperm_dims(ws.ptr, ws.nbytes, A.pointer, ix.pointer, A.eltype, ix.eltype, pre, n, post,
                                      1 if rev else 0, 1 if A.ndims > 1 else 0, current_stream_handle(A.buf.device)))
    return ix


def sortperm(c, dims=None, **kw):
    """Base.sortperm(c::AnyCuVector) -> 1-based Int64 (sorting.jl:1063-1065); N-d needs dims and
    returns linear indices (sorting.jl:1067-1070)."""
    if c.ndims > 1 and dims is None:
        raise ArgumentError("sortperm of a multidimensional array needs dims")  # Base: UndefKeywordError
    ix = CuArray.undef(dt.I64, c.dims, c.buf.device)
    return sortperm_(ix, c, dims=dims, **kw)
