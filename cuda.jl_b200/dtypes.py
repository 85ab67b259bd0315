"""Element types: Julia name <-> b200 enum <-> numpy dtype (include/b200arrays.h b200_dtype_t)."""
import numpy as np

try:  # BFloat16 on the host side (BFloat16s.jl in the reference, CUDACore/Project.toml:46)
    import ml_dtypes

    _bf16 = np.dtype(ml_dtypes.bfloat16)
except Exception:  # pragma: no cover - ml_dtypes is in the image
    ml_dtypes = None
    _bf16 = None

F16, BF16, F32, F64, I32, I64, U32, U64, BOOL, I8, U8, I16, U16 = range(13)

NAMES = {
    F16: "Float16", BF16: "BFloat16", F32: "Float32", F64: "Float64",
    I32: "Int32", I64: "Int64", U32: "UInt32", U64: "UInt64", BOOL: "Bool",
    I8: "Int8", U8: "UInt8", I16: "Int16", U16: "UInt16",
}
SIZES = {F16: 2, BF16: 2, F32: 4, F64: 8, I32: 4, I64: 8, U32: 4, U64: 8, BOOL: 1, I8: 1, U8: 1, I16: 2, U16: 2}

_NP = {
    F16: np.dtype(np.float16), F32: np.dtype(np.float32), F64: np.dtype(np.float64),
    I32: np.dtype(np.int32), I64: np.dtype(np.int64), U32: np.dtype(np.uint32), U64: np.dtype(np.uint64),
    BOOL: np.dtype(np.bool_), I8: np.dtype(np.int8), U8: np.dtype(np.uint8),
    I16: np.dtype(np.int16), U16: np.dtype(np.uint16),
}
if _bf16 is not None:
    _NP[BF16] = _bf16
_FROM_NP = {v: k for k, v in _NP.items()}

FLOATS = (F16, BF16, F32, F64)
SIGNED = (I8, I16, I32, I64)
UNSIGNED = (U8, U16, U32, U64)
INTS = SIGNED + UNSIGNED


def np_dtype(dt):
    return _NP[dt]


def from_np(npdt):
    npdt = np.dtype(npdt)
    if npdt not in _FROM_NP:
        raise TypeError(f"unsupported element type {npdt}")
    return _FROM_NP[npdt]


def widen_add(dt):
    """eltype of Base.add_sum / mul_prod results: BitSignedSmall -> Int, BitUnsignedSmall -> UInt, Bool -> Int."""
    if dt in (I8, I16, I32, BOOL):
        return I64
    if dt in (U8, U16, U32):
        return U64
    return dt
