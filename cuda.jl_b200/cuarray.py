"""CuArray — dense column-major device array, the host-side stand-in for
CUDACore's `CuArray{T,N}` (CUDACore/src/array.jl:74-106).

Storage is a flat torch uint8 CUDA buffer (torch is the allocator / stream
plumbing here, like CUDA.jl's stream-ordered pool, CUDACore/src/memory.jl:698-801);
`dims` is the Julia-style shape, first dimension fastest.  `pointer` is what
`unsafe_convert(CuPtr{T}, ::CuArray)` hands to a ccall (array.jl:467-468).
"""
import math

import numpy as np
import torch

from . import dtypes as dt


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("b200arrays needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    return torch.device("cuda", torch.cuda.current_device())


class CuArray:
    __slots__ = ("buf", "dims", "eltype")

    def __init__(self, data, dims=None, eltype=None):
        """CuArray(host_ndarray) copies host -> device (column-major);
        CuArray(torch_uint8_buffer, dims, eltype) wraps device storage."""
        if isinstance(data, np.ndarray) or np.isscalar(data) or isinstance(data, (list, tuple, range)):
            a = np.asarray(data)
            if a.dtype == np.dtype(object):
                raise TypeError("unsupported host array")
            if eltype is not None:
                a = a.astype(dt.np_dtype(eltype))
            self.eltype = dt.from_np(a.dtype)
            self.dims = tuple(int(d) for d in a.shape)
            flat = np.ascontiguousarray(a.reshape(-1, order="F"))
            host = torch.from_numpy(flat.view(np.uint8).copy())
            self.buf = host.to(_device())
        else:
            assert isinstance(data, torch.Tensor) and data.dtype == torch.uint8 and data.dim() == 1
            self.buf = data
            self.dims = tuple(int(d) for d in dims)
            self.eltype = int(eltype)
            assert self.buf.numel() == self.length * dt.SIZES[self.eltype]

    # -- construction helpers (similar / zeros) --------------------------------
    @staticmethod
    def undef(eltype, dims, device=None):
        """CuArray{T}(undef, dims...) — uninitialised pool memory."""
        dims = tuple(int(d) for d in (dims if isinstance(dims, (tuple, list)) else (dims,)))
        n = math.prod(dims) if dims else 1
        buf = torch.empty(n * dt.SIZES[eltype], dtype=torch.uint8, device=device or _device())
        return CuArray(buf, dims, eltype)

    @staticmethod
    def zeros(eltype, dims, device=None):
        a = CuArray.undef(eltype, dims, device)
        a.buf.zero_()
        return a

    @staticmethod
    def from_torch(t, dims=None, eltype=None):
        """Wrap a contiguous torch CUDA tensor WITHOUT copying; its memory order is
        taken as column-major over `dims` (default: 1-D)."""
        assert t.is_cuda and t.is_contiguous()
        if eltype is None:
            eltype = _TORCH_TO_DT[t.dtype]
        if dims is None:
            dims = (t.numel(),)
        return CuArray(t.reshape(-1).view(torch.uint8), dims, eltype)

    def similar(self, eltype=None, dims=None):
        """similar(A, T, dims) (array.jl:185-190)."""
        return CuArray.undef(self.eltype if eltype is None else eltype, self.dims if dims is None else dims,
                             self.buf.device)

    # -- queries ------------------------------------------------------------------
    @property
    def length(self):
        return math.prod(self.dims) if self.dims else 1

    @property
    def ndims(self):
        return len(self.dims)

    def size(self, d=None):
        if d is None:
            return self.dims
        return self.dims[d - 1] if d <= len(self.dims) else 1  # 1-based like Julia

    @property
    def pointer(self):
        return self.buf.data_ptr()

    @property
    def nbytes(self):
        return self.buf.numel()

    def reshape(self, *dims):
        dims = dims[0] if len(dims) == 1 and isinstance(dims[0], (tuple, list)) else dims
        assert math.prod(dims) == self.length
        return CuArray(self.buf, tuple(dims), self.eltype)

    def vec(self):
        return self.reshape((self.length,))

    def copy(self):
        return CuArray(self.buf.clone(), self.dims, self.eltype)

    def typed(self):
        """The storage as a typed 1-D torch tensor (plumbing for collectives)."""
        if self.buf.numel() == 0:
            return torch.empty(0, dtype=_DT_TO_TORCH[self.eltype], device=self.buf.device)
        return self.buf.view(_DT_TO_TORCH[self.eltype])

    # -- host transfer: Array(c) ----------------------------------------------------
    def to_host(self):
        host = self.buf.cpu().numpy()
        return host.view(dt.np_dtype(self.eltype)).reshape(self.dims, order="F")

    def __repr__(self):
        return f"CuArray{{{dt.NAMES[self.eltype]}}}{self.dims}"


def Array(c):
    """Array(::CuArray) — device -> host copy (synchronises like the reference's)."""
    return c.to_host()


_DT_TO_TORCH = {
    dt.F16: torch.float16, dt.BF16: torch.bfloat16, dt.F32: torch.float32, dt.F64: torch.float64,
    dt.I32: torch.int32, dt.I64: torch.int64, dt.U32: torch.uint32, dt.U64: torch.uint64,
    dt.BOOL: torch.bool, dt.I8: torch.int8, dt.U8: torch.uint8, dt.I16: torch.int16, dt.U16: torch.uint16,
}
_TORCH_TO_DT = {v: k for k, v in _DT_TO_TORCH.items()}


class Workspace:
    """with_workspace (CUDACore/src/utils/call.jl:23-103): a temporary from the
    stream-ordered caching allocator, sized by the library's *_workspace_bytes query."""

    def __init__(self, nbytes, device):
        self.nbytes = int(nbytes)
        self.buf = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device) if nbytes else None

    @property
    def ptr(self):
        return self.buf.data_ptr() if self.buf is not None else None


def current_stream_handle(device):
    """The task-local CuStream handle (lib/cudadrv/state.jl:289): torch's current stream here."""
    return torch.cuda.current_stream(device).cuda_stream
