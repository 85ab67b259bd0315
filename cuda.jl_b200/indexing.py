"""Host-side mirror of `Base.findall(bools::AnyCuArray{Bool})` (CUDACore/src/indexing.jl:26-66),
the in-tree caller of `cumsum`: one fused scan + compact pass instead of
cumsum -> host read-back -> scatter."""
import ctypes

from . import _lib
from . import dtypes as dt
from .cuarray import CuArray, Workspace, current_stream_handle
from .errors import NotGraftedError


def findall(bools):
    """1-based linear indices (Int64) of the true elements, in order.  (For N-d inputs the
    reference converts to CartesianIndices on top of the same linear indices.)"""
    if bools.eltype != dt.BOOL:
        raise NotGraftedError("findall with a predicate closure stays on the reference path")
    n = bools.length
    dev = bools.buf.device
    out = CuArray.undef(dt.I64, (n,), dev)
    count = CuArray.undef(dt.I64, (1,), dev)
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.b200_select_workspace_bytes(n, ctypes.byref(nbytes)))
    ws = Workspace(nbytes.value, dev)
    _lib.check(lib.b200_select_flagged(ws.ptr, ws.nbytes, bools.pointer if n else None, out.pointer if n else None,
                                       dt.I64, count.pointer, n, current_stream_handle(dev)))
    m = int(count.to_host()[0])            # the reference reads indices[end] back the same way (indexing.jl:30)
    return CuArray(out.buf[:m * 8], (m,), dt.I64)
