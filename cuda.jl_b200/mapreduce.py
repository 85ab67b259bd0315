"""Host-side mirror of CUDACore/src/mapreduce.jl and of the Base / GPUArrays callers
that reach it (SURVEY §3.1, §8 a1/a1').

`mapreducedim_` is `GPUArrays.mapreducedim!(f, op, R, A; init)` (mapreduce.jl:168);
`sum`, `prod`, `maximum`, ... restate what Base + GPUArrays._mapreduce do before
calling it: output eltype, neutral `init`, reduced output shape, scalar read-back
for `dims=:`.  All device work goes through libb200arrays' `b200_reduce`.
"""
import builtins as _bi
import ctypes

import numpy as np

from . import _lib
from . import dtypes as dt
from .cuarray import CuArray, Workspace, current_stream_handle
from .errors import ArgumentError, DimensionMismatch, NotGraftedError


# --- operator / map tokens (Julia function objects the methods are registered for) ---
class _Fn:
    def __init__(self, name, code=None):
        self.name, self.code = name, code

    def __repr__(self):
        return self.name


identity = _Fn("identity", 0)
abs2 = _Fn("abs2", 1)
add_sum = _Fn("Base.add_sum", 0)
mul_prod = _Fn("Base.mul_prod", 1)
add = _Fn("+", 0)
mul = _Fn("*", 1)
max_ = _Fn("max", 2)
min_ = _Fn("min", 3)
and_ = _Fn("&", 4)
or_ = _Fn("|", 5)
extrema_rf = _Fn("Base._extrema_rf", 6)

_OPS = (add_sum, mul_prod, add, mul, max_, min_, and_, or_, extrema_rf)
_MAPS = (identity, abs2)
INIT_NEUTRAL, INIT_FROM_OUT = 0, 1


def neutral_element(op, eltype):
    """GPUArrays.neutral_element(op, T) for the grafted ops (SURVEY §8 a1').  NB: for `&` this is one(T), which is
    the true neutral element only for Bool — see `true_neutral`."""
    npdt = dt.np_dtype(eltype)
    isf = eltype in dt.FLOATS
    if op in (add, add_sum, or_):
        return npdt.type(0)
    if op in (mul, mul_prod, and_):
        return npdt.type(1)
    if op is max_:
        return npdt.type(-np.inf) if isf else (npdt.type(False) if eltype == dt.BOOL else np.iinfo(npdt).min)
    if op is min_:
        return npdt.type(np.inf) if isf else (npdt.type(True) if eltype == dt.BOOL else np.iinfo(npdt).max)
    raise NotGraftedError(f"no neutral element for {op}")


def true_neutral(op, eltype):
    """The element e with op(e, x) == x for every x (lib/B200Arrays/src/mapreduce.jl `true_neutral`).  Differs from
    GPUArrays.neutral_element for integer `&` (all-ones, not one(T)): the reference seeds every thread with the
    value it is given (mapreduce.jl:114,148), so `reduce(&, ints)` with GPUArrays' one(T) is not a plain AND and stays
    on the reference path."""
    if op is and_ and eltype != dt.BOOL:
        if eltype in dt.FLOATS:
            raise NotGraftedError("bitwise ops on floats")
        npdt = dt.np_dtype(eltype)
        return npdt.type(~np.uint64(0)) if eltype in dt.UNSIGNED else npdt.type(-1)
    return neutral_element(op, eltype)


def _is_neutral(op, eltype, init):
    try:
        ne = true_neutral(op, eltype)
    except NotGraftedError:
        return False
    a = np.asarray(init).astype(dt.np_dtype(eltype))
    return a.tobytes() == np.asarray(ne, dtype=dt.np_dtype(eltype)).tobytes()


def _factor(adims, rdims):
    """(pre, red, post) factorisation of the reduced dims (those with size(R,d)==1 != size(A,d),
    mapreduce.jl:196).  Only one contiguous group of reduced dims is grafted."""
    nd = len(adims)
    rdims = tuple(rdims) + (1,) * (nd - len(rdims))  # pad R with singleton dims (mapreduce.jl:187-190)
    for d in range(nd):
        if rdims[d] != adims[d] and rdims[d] != 1:
            raise DimensionMismatch(f"cannot reduce {adims} into {rdims}")  # Base.check_reducedims
    if _bi.any(r != 1 for r in rdims[nd:]):
        raise DimensionMismatch(f"cannot reduce {adims} into {rdims}")
    reduced = [d for d in range(nd) if rdims[d] == 1 and adims[d] != 1]
    if not reduced:
        return int(np.prod(adims, dtype=np.int64)), 1, 1
    lo, hi = reduced[0], reduced[-1]
    for d in range(lo, hi + 1):
        if adims[d] != 1 and d not in reduced:
            raise NotGraftedError(f"reduced dims {tuple(x + 1 for x in reduced)} of {adims} are not contiguous")
    pre = int(np.prod(adims[:lo], dtype=np.int64))
    red = int(np.prod(adims[lo:hi + 1], dtype=np.int64))
    post = int(np.prod(adims[hi + 1:], dtype=np.int64))
    return pre, red, post


def _call_reduce(A, R, opcode, mapcode, pre, red, post, init_mode):
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.b200_reduce_workspace_bytes(A.eltype, R.eltype, opcode, mapcode, pre, red, post,
                                               ctypes.byref(nbytes)))
    ws = Workspace(nbytes.value, A.buf.device)
    stream = current_stream_handle(A.buf.device)
    _lib.check(lib.b200_reduce(ws.ptr, ws.nbytes, A.pointer, R.pointer, A.eltype, R.eltype, opcode, mapcode,
                               pre, red, post, init_mode, stream))


def mapreducedim_(f, op, R, A, init=None):
    """GPUArrays.mapreducedim!(f, op, R::CuArray, A::CuArray; init) (mapreduce.jl:168-301).

    Returns the R it was given (mapreduce.jl:211,300).  Grafted for f in {identity, abs2},
    the built-in ops, and `init` either `None` (fold R's prior contents once) or the
    operator's neutral element; anything else raises NotGraftedError (Julia: falls
    back to the reference's generic method).
    """
    if f not in _MAPS or op not in _OPS:
        raise NotGraftedError(f"mapreducedim!({f}, {op}, ...) stays on the reference's generic kernels")
    if not isinstance(R, CuArray) or not isinstance(A, CuArray):
        raise NotGraftedError("only dense CuArray inputs and outputs are grafted")
    if op is extrema_rf:
        if f is not identity:
            raise NotGraftedError("extrema with a map is not grafted")
        if R.dims[0] != 2:
            raise DimensionMismatch("extrema output must carry (min, max) pairs along a leading dim of 2")
        rdims = R.dims[1:]
    else:
        rdims = R.dims
    pre, red, post = _factor(A.dims, rdims)
    if A.length == 0:
        return R  # mapreduce.jl:175
    if init is None:
        mode = INIT_FROM_OUT
    elif op is extrema_rf or _is_neutral(op, R.eltype, init):
        mode = INIT_NEUTRAL
    else:
        raise NotGraftedError("non-neutral init gives launch-shape-dependent results in the reference "
                              "(mapreduce.jl:114,148); not grafted")
    _call_reduce(A, R, op.code, f.code, pre, red, post, mode)
    return R


# --- Base / GPUArrays callers ---------------------------------------------------------
def _out_eltype(f, op, eltype):
    """Base.promote_op(op, ...) restated for the grafted (f, op, T) matrix (SURVEY §8 a1')."""
    if op in (add_sum, mul_prod):
        return dt.widen_add(eltype)
    if op in (and_, or_) and eltype in dt.FLOATS:
        raise NotGraftedError("bitwise ops on floats")
    return eltype


def _reduced_dims(adims, dims):
    if dims is None:
        return tuple(1 for _ in adims)
    ds = (dims,) if isinstance(dims, int) else tuple(dims)
    for d in ds:
        if d < 1:
            raise ArgumentError("region dimension(s) must be ≥ 1")
    return tuple(1 if (i + 1) in ds else s for i, s in enumerate(adims))


def mapreduce(f, op, A, dims=None, init=None):
    """Base.mapreduce(f, op, A::AnyGPUArray; dims, init) -> GPUArrays._mapreduce.

    dims=None is Julia's `dims=:`: returns a host scalar (`@allowscalar R[]`)."""
    if f not in _MAPS or op not in _OPS:
        raise NotGraftedError(f"mapreduce({f}, {op}, ...) stays on the reference's generic kernels")
    et = _out_eltype(f, op, A.eltype)
    if op is extrema_rf:
        rd = _reduced_dims(A.dims, dims)
        if A.length == 0:
            raise ArgumentError("reducing over an empty collection is not allowed")
        R = CuArray.undef(et, (2,) + rd, A.buf.device)
        mapreducedim_(f, op, R, A, init=0)
        if dims is None:
            h = R.to_host().reshape(-1)
            return (h[0], h[1])
        return R
    if init is None:
        try:
            init = neutral_element(op, et)
        except NotGraftedError:
            raise
        if A.length == 0 and op in (max_, min_):
            raise ArgumentError("reducing over an empty collection is not allowed")
    rd = _reduced_dims(A.dims, dims)
    R = CuArray.undef(et, rd, A.buf.device)
    if A.length == 0 or R.length == 0:
        # _mapreduce fills R with init for empty inputs
        if R.length:
            R.typed_fill(init) if hasattr(R, "typed_fill") else _fill(R, init)
        return R.to_host().reshape(-1)[0] if dims is None else R
    mapreducedim_(f, op, R, A, init=init)
    if dims is None:
        return R.to_host().reshape(-1)[0]
    return R


def _fill(R, value):
    host = np.full(R.length, value, dtype=dt.np_dtype(R.eltype))
    R.buf.copy_(CuArray(host).buf)


def reduce(op, A, dims=None, init=None):
    return mapreduce(identity, op, A, dims=dims, init=init)


def _with_f(default_op):
    def fn(*args, dims=None):
        if len(args) == 2:
            f, A = args
        else:
            (A,) = args
            f = identity
        return mapreduce(f, default_op, A, dims=dims)
    return fn


sum = _with_f(add_sum)      # noqa: A001  (Base.sum)
prod = _with_f(mul_prod)
maximum = _with_f(max_)
minimum = _with_f(min_)


def extrema(A, dims=None):
    return mapreduce(identity, extrema_rf, A, dims=dims)


def count(A, dims=None):
    """count(::CuArray{Bool}) = mapreduce(identity, add_sum, A; init=0) -> Int64."""
    if A.eltype != dt.BOOL:
        raise NotGraftedError("count with a predicate closure stays on the reference path")
    return mapreduce(identity, add_sum, A, dims=dims)


def any(A, dims=None):  # noqa: A001
    if A.eltype != dt.BOOL:
        raise NotGraftedError("any with a predicate closure stays on the reference path")
    r = mapreduce(identity, or_, A, dims=dims)
    return bool(r) if dims is None else r


def all(A, dims=None):  # noqa: A001
    if A.eltype != dt.BOOL:
        raise NotGraftedError("all with a predicate closure stays on the reference path")
    r = mapreduce(identity, and_, A, dims=dims)
    return bool(r) if dims is None else r


def _bang(op):
    def fn(*args):
        """Base.sum!(R, A) family: R is (re)initialised to the neutral element, then
        mapreducedim!(f, op, R, A) with init === nothing folds into it."""
        if len(args) == 3:
            f, R, A = args
        else:
            R, A = args
            f = identity
        _fill(R, neutral_element(op, R.eltype))
        return mapreducedim_(f, op, R, A, init=None)
    return fn


sum_ = _bang(add_sum)
prod_ = _bang(mul_prod)
maximum_ = _bang(max_)
minimum_ = _bang(min_)


# --- findmax / findmin / argmax / argmin ------------------------------------
def _findminmax(A, dims, is_max):
    """Base.findmax(A; dims) / findmin on a CuArray (reference: GPUArrays' findminmax through
    mapreducedim!; tests test/core/array.jl:767-769,782-784).  Returns (values, linear 1-based
    Int64 indices); with dims=None a host (value, index) pair.  (Julia returns CartesianIndex for
    N-d inputs: the linear index carries the same information.)"""
    if A.length == 0:
        raise ArgumentError("reducing over an empty collection is not allowed")
    rd = _reduced_dims(A.dims, dims)
    pre, red, post = _factor(A.dims, rd)
    vals = CuArray.undef(A.eltype, rd, A.buf.device)
    idx = CuArray.undef(dt.I64, rd, A.buf.device)
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.b200_findminmax_workspace_bytes(A.eltype, pre, red, post, ctypes.byref(nbytes)))
    ws = Workspace(nbytes.value, A.buf.device)
    _lib.check(lib.b200_findminmax(ws.ptr, ws.nbytes, A.pointer, vals.pointer, idx.pointer, A.eltype,
                                   1 if is_max else 0, pre, red, post, current_stream_handle(A.buf.device)))
    if dims is None:
        return vals.to_host().reshape(-1)[0], int(idx.to_host().reshape(-1)[0])
    return vals, idx


def findmax(A, dims=None):
    return _findminmax(A, dims, True)


def findmin(A, dims=None):
    return _findminmax(A, dims, False)


def argmax(A, dims=None):
    r = _findminmax(A, dims, True)
    return r[1]


def argmin(A, dims=None):
    r = _findminmax(A, dims, False)
    return r[1]
