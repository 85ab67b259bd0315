"""Host-side mirror of CUDACore/src/accumulate.jl: `scan_` is `scan!` (accumulate.jl:135-199),
`accumulate_` / `accumulate` / `cumsum` / `cumprod` restate the Base hooks at :204-234.
All device work is one call to libb200arrays' `b200_scan` (single-pass look-back scan)."""
import ctypes

import numpy as np

from . import _lib
from . import dtypes as dt
from .cuarray import CuArray, Workspace, current_stream_handle
from .errors import ArgumentError, DimensionMismatch, NotGraftedError
from .mapreduce import add, add_sum, and_, max_, min_, mul, mul_prod, or_

_SCAN_OPS = {add: 0, add_sum: 0, mul: 1, mul_prod: 1, max_: 2, min_: 3, and_: 4, or_: 5}


def _factor(dims_tuple, dims):
    pre = int(np.prod(dims_tuple[:dims - 1], dtype=np.int64))
    n = int(dims_tuple[dims - 1])
    post = int(np.prod(dims_tuple[dims:], dtype=np.int64))
    return pre, n, post


def scan_(f, output, input, dims, init=None, exclusive=False, carry=None, carry_count=0, neutral=None):
    """CUDACore.scan!(f, output, input; dims, init) (accumulate.jl:135).  `init` is the x of
    Some(x).  `exclusive=True` is the extra variant BASELINE config 3 names.  `carry` (a device CuArray of the
    accumulator type) with `carry_count` is the sharded scan's device-side carry (b200_scan_carry): its first
    carry_count values are folded into the seed on the device, no host round trip."""
    if f not in _SCAN_OPS:
        raise NotGraftedError(f"scan!({f}, ...) stays on the reference's generic kernels")
    # scan!'s `neutral` keyword defaults to GPUArrays.neutral_element(f, T) (accumulate.jl:136) and the reference kernel
    # loads every element as op(neutral, x) (accumulate.jl:39-43): for integer `&` that default is one(T), so the
    # reference's own result is x & 1-contaminated.  The graft computes the plain scan, hence it only answers when the
    # caller's neutral IS neutral (lib/B200Arrays/src/accumulate.jl has the same guard); Bool `&` is unaffected.
    from .mapreduce import and_, true_neutral
    if f is and_ and output.eltype != dt.BOOL:
        if neutral is None or np.asarray(neutral).astype(dt.np_dtype(output.eltype)).tobytes() != \
                np.asarray(true_neutral(and_, output.eltype), dtype=dt.np_dtype(output.eltype)).tobytes():
            raise NotGraftedError("accumulate(&, ::CuArray{<:Integer}) with GPUArrays' neutral element one(T) stays on the "
                                  "reference path")
    if not (isinstance(dims, (int, np.integer)) and dims > 0):
        raise ArgumentError("dims must be a positive integer")            # accumulate.jl:137
    if output.dims != input.dims:
        raise DimensionMismatch("shape of B must match A")                 # accumulate.jl:139
    if dims > input.ndims:                                                  # accumulate.jl:140: copyto!
        if output.eltype != input.eltype:
            raise NotGraftedError("converting copy")
        output.buf.copy_(input.buf)
        return output
    if input.dims[dims - 1] == 0 or input.length == 0:                      # accumulate.jl:141
        return output
    pre, n, post = _factor(input.dims, dims)
    lib = _lib.load()
    init_buf = None
    if init is not None:
        init_buf = np.asarray(init).astype(dt.np_dtype(output.eltype)).reshape(1).copy()
    nbytes = ctypes.c_size_t(0)
    opcode = _SCAN_OPS[f]
    _lib.check(lib.b200_scan_workspace_bytes(input.eltype, output.eltype, opcode, pre, n, post, ctypes.byref(nbytes)))
    ws = Workspace(nbytes.value, input.buf.device)
    if carry is not None and carry_count > 0:
        _lib.check(lib.b200_scan_carry(ws.ptr, ws.nbytes, input.pointer, output.pointer, input.eltype, output.eltype,
                                       opcode, 1 if exclusive else 0, pre, n, post,
                                       init_buf.ctypes.data if init_buf is not None else None,
                                       carry.pointer, int(carry_count), current_stream_handle(input.buf.device)))
        return output
    _lib.check(lib.b200_scan(ws.ptr, ws.nbytes, input.pointer, output.pointer, input.eltype, output.eltype, opcode,
                             1 if exclusive else 0, pre, n, post,
                             init_buf.ctypes.data if init_buf is not None else None,
                             current_stream_handle(input.buf.device)))
    return output


def accumulate_(op, B, A, dims=None, init=None, exclusive=False):
    """Base.accumulate!(op, B, A; dims, init) -> Base._accumulate! (accumulate.jl:204-214)."""
    if dims is None:
        if A.ndims != 1:
            raise ArgumentError("Keyword argument dims must be provided for multidimensional arrays")
        dims = 1
    return scan_(op, B, A, dims, init=init, exclusive=exclusive)


def _promote_op(op, eltype, init=None):
    if op in (add_sum, mul_prod):
        return dt.widen_add(eltype)
    if op in (add, mul) and eltype == dt.BOOL:
        return dt.I64                      # promote_op(+, Bool, Bool) == Int
    if init is not None and isinstance(init, np.generic):
        it = dt.from_np(init.dtype)
        if it != eltype:
            # promote_op(op, typeof(init), eltype(A)) for the float/int widths we graft
            return dt.from_np(np.promote_types(init.dtype, dt.np_dtype(eltype))) if eltype != dt.BF16 else eltype
    return eltype


def accumulate(op, A, dims=None, **kw):
    """Base.accumulate(op, A::AnyCuArray; dims, init) (accumulate.jl:219-234)."""
    if dims is None and A.ndims != 1:
        flat = accumulate(op, A.vec().copy(), **kw)                        # A[:] flattening branch (:221-224)
        return flat.reshape(A.dims)
    exclusive = kw.pop("exclusive", False)
    if not kw:
        out = A.similar(_promote_op(op, A.eltype))
    elif tuple(kw.keys()) == ("init",):
        out = A.similar(_promote_op(op, A.eltype, kw["init"]))
    else:
        bad = [k for k in kw if k != "init"]
        raise ArgumentError("accumulate does not support the keyword arguments [" +
                            ", ".join(":" + k for k in bad) + "]")          # accumulate.jl:231
    return accumulate_(op, out, A, dims=dims, exclusive=exclusive, **kw)


def cumsum(A, dims=None):
    """Base.cumsum: out = similar(A, promote_op(add_sum, T, T)); accumulate!(add_sum, out, A; dims)."""
    if dims is None:
        if A.ndims != 1:
            raise ArgumentError("cumsum on a multidimensional array needs dims")   # Base: UndefKeywordError
        dims = 1
    out = A.similar(dt.widen_add(A.eltype))
    return accumulate_(add_sum, out, A, dims=dims)


def cumprod(A, dims=None):
    if dims is None:
        if A.ndims != 1:
            raise ArgumentError("cumprod on a multidimensional array needs dims")
        dims = 1
    out = A.similar(dt.widen_add(A.eltype))
    return accumulate_(mul_prod, out, A, dims=dims)


def cumsum_(B, A, dims=None):
    return accumulate_(add_sum, B, A, dims=dims)


def cumprod_(B, A, dims=None):
    return accumulate_(mul_prod, B, A, dims=dims)
