"""Parity of b200_scan (through the host mirror of accumulate.jl) against the CPU oracle.

Follows test/core/array.jl:425-456 ("accumulate"): sizes {0,1,2,3,10,10_000,16384,16385},
vectors and n x 2 matrices, with / without init, N-d with dims on Int, in-place, cumsum /
cumprod, the bad-kwarg error text.  Integers bit-exact; floats within
4*log2(n)*eps(T)*cumsum|x| of the Float64 reference.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(0x5CA9)


def npdt(name):
    return oracle.bfloat16 if name == "bfloat16" else np.dtype(name)


def rand(name, shape):
    d = npdt(name)
    if oracle.is_float(d):
        return np.asfortranarray(RNG.random(size=shape).astype(np.float32).astype(d))
    if d == np.bool_:
        return np.asfortranarray(RNG.random(size=shape) < 0.5)
    ii = np.iinfo(d)
    return np.asfortranarray(RNG.integers(ii.min, ii.max, size=shape, dtype=d, endpoint=True))


def check_scan(b, opname, A, dims=None, init=None, exclusive=False, out_name=None):
    op = {"add": b.add, "mul": b.mul, "max": b.max_, "min": b.min_, "and": b.and_, "or": b.or_}[opname]
    out_dt = npdt(out_name) if out_name else A.dtype
    d = b.CuArray(A)
    kw = {}
    if init is not None:
        kw["init"] = init
    if exclusive:
        kw["exclusive"] = True
    if out_name:
        out = d.similar(b.dtypes.from_np(out_dt))
        R = b.accumulate_(op, out, d, dims=dims, **kw)
    else:
        R = b.accumulate(op, d, dims=dims, **kw)
    got = b.Array(R)
    if dims is None and A.ndim > 1:
        ref = oracle.accumulate(opname, A.reshape(-1, order="F"), 1, init, out_dt, exclusive).reshape(A.shape, order="F")
        Aflat, d_ = A.reshape(-1, order="F"), 1
        tol_shape = lambda t: t.reshape(A.shape, order="F")
    else:
        d_ = 1 if dims is None else dims
        ref = oracle.accumulate(opname, A, d_, init, out_dt, exclusive)
        Aflat = A
        tol_shape = lambda t: t
    assert got.shape == A.shape and got.dtype == out_dt
    if oracle.is_float(out_dt) and opname in ("add", "mul"):
        if d_ > A.ndim:
            np.testing.assert_array_equal(got.astype(np.float64), ref.astype(np.float64))
            return
        tol = tol_shape(oracle.scan_tolerance(Aflat, d_, out_dt, init))
        if exclusive:  # the bound of the preceding prefix, shifted like the data
            tol = np.concatenate([np.take(tol, [0], axis=d_ - 1) * 0 + tol.max() * 0 + np.finfo(np.float64).tiny + (abs(float(init)) if init is not None else 0) * oracle.eps(out_dt),
                                  np.take(tol, range(0, tol.shape[d_ - 1] - 1), axis=d_ - 1)], axis=d_ - 1) if tol.shape[d_ - 1] else tol
        err = np.abs(got.astype(np.float64) - ref)
        assert np.all(err <= tol + oracle.eps(out_dt) * np.abs(ref)), f"max err {err.max()}"
    else:
        np.testing.assert_array_equal(got.astype(np.float64) if oracle.is_float(out_dt) else got,
                                      ref.astype(np.float64) if oracle.is_float(out_dt) else ref)


SIZES = (0, 1, 2, 3, 10, 10_000, 16384, 16384 + 1)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_accumulate_reference_sizes(b, n, dtype):
    """test/core/array.jl:426-430."""
    check_scan(b, "add", rand(dtype, (n,)))
    check_scan(b, "add", rand(dtype, (n, 2)))
    check_scan(b, "add", rand(dtype, (n,)), init=npdt(dtype).type(RNG.random()))


@pytest.mark.parametrize("sizes,dims", [((2,), 2), ((3, 4, 5), 2), ((1, 70, 50, 20), 3)])
def test_accumulate_multidimensional_int(b, sizes, dims):
    """test/core/array.jl:432-446 (rand(Int, sizes), with and without init)."""
    A = rand("int64", sizes)
    check_scan(b, "add", A, dims=dims)
    check_scan(b, "add", A)
    y = RNG.integers(-2**62, 2**62, dtype=np.int64)
    check_scan(b, "add", A, dims=dims, init=np.int64(y))
    check_scan(b, "add", A, init=np.int64(y))


def test_in_place_and_specialized(b):
    """test/core/array.jl:448-453."""
    x = rand("float64", (2,))
    d = b.CuArray(x)
    b.accumulate_(b.add, d, d.copy())
    np.testing.assert_allclose(b.Array(d), np.cumsum(x))
    d2 = b.CuArray(x)
    b.accumulate_(b.add, d2, d2)                      # exact aliasing
    np.testing.assert_allclose(b.Array(d2), np.cumsum(x))
    np.testing.assert_allclose(b.Array(b.cumsum(b.CuArray(x))), np.cumsum(x))
    np.testing.assert_allclose(b.Array(b.cumprod(b.CuArray(x))), np.cumprod(x))
    big = rand("int64", (1 << 20,))
    db = b.CuArray(big)
    b.accumulate_(b.add, db, db)
    np.testing.assert_array_equal(b.Array(db), oracle.accumulate("add", big))


def test_bad_kwarg_error_text(b):
    """test/core/array.jl:455."""
    with pytest.raises(b.ArgumentError) as e:
        b.accumulate(b.add, b.CuArray(np.zeros(1024, dtype=np.float32)), bad_kwarg="bad")
    assert str(e.value) == "accumulate does not support the keyword arguments [:bad_kwarg]"
    with pytest.raises(b.ArgumentError):
        b.scan_(b.add, b.CuArray(np.zeros(4)), b.CuArray(np.zeros(4)), 0)
    with pytest.raises(b.DimensionMismatch):
        b.scan_(b.add, b.CuArray(np.zeros(5)), b.CuArray(np.zeros(4)), 1)


@pytest.mark.parametrize("dtype", ["float16", "bfloat16", "float32", "float64", "int32", "int64", "uint32", "uint64",
                                   "int8", "uint8", "int16", "uint16"])
@pytest.mark.parametrize("n", [4095, 4096, 4097, 100_003, (1 << 21) + 5])
def test_cumsum_types(b, dtype, n):
    A = rand(dtype, (n,))
    if dtype in ("float16", "bfloat16"):
        A = (A.astype(np.float32) * (0.01 if n > 5000 else 1.0)).astype(npdt(dtype))
    d = b.CuArray(A)
    R = b.cumsum(d)
    out_dt = oracle.widen_add(A.dtype)
    got = b.Array(R)
    assert got.dtype == out_dt
    ref = oracle.accumulate("add", A, 1, None, out_dt)
    if oracle.is_float(out_dt):
        tol = oracle.scan_tolerance(A, 1, out_dt)
        assert np.all(np.abs(got.astype(np.float64) - ref) <= tol + oracle.eps(out_dt) * np.abs(ref))
    else:
        np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("op", ["max", "min", "mul", "and", "or"])
@pytest.mark.parametrize("dtype", ["float32", "int32", "uint64", "float64", "int8", "uint16"])
def test_other_ops(b, op, dtype):
    if op in ("and", "or") and dtype.startswith("float"):
        pytest.skip("bitwise on floats")
    A = rand(dtype, (70_001,))
    if op == "mul" and oracle.is_float(npdt(dtype)):
        A = (1 + (A - 0.5) * 1e-3).astype(npdt(dtype))[:5000]
    if op == "and":
        # integer `&`: GPUArrays' default neutral one(T) is not neutral and the reference kernel applies it to every element
        # (accumulate.jl:39-43), so accumulate(&, ints) stays on the reference path; scan! with the true neutral is grafted
        A = A | np.uint64(0x00FF00FF).astype(npdt(dtype))
        d = b.CuArray(A)
        with pytest.raises(b.NotGraftedError):
            b.accumulate(b.and_, d)
        out = b.scan_(b.and_, d.similar(), d, 1, neutral=b.true_neutral(b.and_, d.eltype))
        np.testing.assert_array_equal(b.Array(out), oracle.accumulate("and", A, 1, None, A.dtype, False))
        return
    check_scan(b, op, A)
    check_scan(b, op, np.asfortranarray(A[:69993 if A.size > 69993 else 4998].reshape(3, -1)), dims=2)


def test_bool_scans(b):
    M = rand("bool", (50_001,))
    d = b.CuArray(M)
    np.testing.assert_array_equal(b.Array(b.cumsum(d)), np.cumsum(M, dtype=np.int64))
    np.testing.assert_array_equal(b.Array(b.accumulate(b.or_, d)), np.logical_or.accumulate(M))
    np.testing.assert_array_equal(b.Array(b.accumulate(b.and_, d)), np.logical_and.accumulate(M))
    np.testing.assert_array_equal(b.Array(b.accumulate(b.add, d)), np.cumsum(M, dtype=np.int64))


@pytest.mark.parametrize("shape,dims", [((4096, 7), 1), ((4097, 3), 1), ((1000, 1000), 1), ((1000, 1000), 2),
                                        ((33, 65, 17), 2), ((33, 65, 17), 3), ((33, 65, 17), 1), ((5, 20000), 2),
                                        ((16385, 2), 1), ((2, 100_000), 1)])
@pytest.mark.parametrize("dtype", ["int64", "float32", "int32"])
def test_dims(b, shape, dims, dtype):
    check_scan(b, "add", rand(dtype, shape), dims=dims)


@pytest.mark.parametrize("shape,dims", [((3, 1_000_000), 1), ((3, 1_000_000), 2), ((2, 70_001), 2), ((5, 40_000), 2),
                                        ((100, 3_001), 2), ((128, 1_000), 2), ((2, 30_000, 3), 2), ((1, 9_999, 7), 2),
                                        ((17, 300_000), 1), ((128, 5_000), 1), ((129, 5_000), 1)])
@pytest.mark.parametrize("dtype", ["int64", "float32", "uint8"])
def test_reference_benchmark_shapes(b, shape, dims, dtype):
    """perf/array.jl:8-9,72-90 "L" shapes (3 x 1_000_000) and neighbours: very many very short segments
    (thread per segment) and few long columns (scan_fewcols_kernel), inclusive / exclusive / init."""
    A = rand(dtype, shape)
    if dtype == "float32":
        A = (A * 0.01).astype(np.float32)
    check_scan(b, "add", A, dims=dims)
    check_scan(b, "max", A, dims=dims)
    if shape[0] <= 5:
        check_scan(b, "add", A, dims=dims, exclusive=True, init=npdt(dtype).type(3))


@pytest.mark.parametrize("dtype", ["int64", "float32", "int32", "float64"])
@pytest.mark.parametrize("n", [1, 5, 4096, 8193, 300_001])
def test_exclusive(b, dtype, n):
    A = rand(dtype, (n,))
    check_scan(b, "add", A, exclusive=True)
    check_scan(b, "add", A, exclusive=True, init=npdt(dtype).type(3))
    if n > 1:
        check_scan(b, "add", np.asfortranarray(rand(dtype, (n // 2 + 1, 2))), dims=2, exclusive=True)


def test_widening_out(b):
    A = rand("int32", (200_001,))
    check_scan(b, "add", A, out_name="int64")
    Af = rand("float32", (200_001,))
    check_scan(b, "add", Af, out_name="float64")


def test_unaligned_view(b):
    base = rand("int32", (100_003,))
    d = b.CuArray(base)
    v = b.CuArray(d.buf[4:], (base.size - 1,), d.eltype)       # 4-byte offset: scalar-chunk path
    np.testing.assert_array_equal(b.Array(b.accumulate(b.add, v)), oracle.accumulate("add", base[1:]))
    M = rand("int32", (4099, 5))                                # odd segment length: segments unaligned
    check_scan(b, "add", M, dims=1)
    # widening scans: the input chunk is narrower than 16 bytes, so an 8-byte-aligned Int32 view or 8-byte-multiple
    # segments pass the chunk test but must not reach the bulk-load (16-byte) kernels
    out = b.CuArray.undef(b.dtypes.I64, (base.size - 2,))
    v8 = b.CuArray(d.buf[8:], (base.size - 2,), d.eltype)
    np.testing.assert_array_equal(b.Array(b.cumsum_(out, v8)), np.cumsum(base[2:].astype(np.int64)))
    M2 = rand("int32", (70_002, 3))                             # 280008-byte segments: 8- but not 16-byte multiples
    check_scan(b, "add", M2, dims=1, out_name="int64")


@pytest.mark.parametrize("n", [0, 1, 5, 4095, 4096, 4097, 100_003, (1 << 22) + 11])
@pytest.mark.parametrize("p", [0.0, 0.5, 1.0, 0.01])
def test_findall_fused_select(b, n, p):
    """findall(::CuArray{Bool}) (CUDACore/src/indexing.jl:26-66): scan + compact in one pass."""
    m = RNG.random(n) < p
    got = b.Array(b.findall(b.CuArray(m)))
    np.testing.assert_array_equal(got, oracle.findall(m))
    if n >= 4096:
        M = np.asfortranarray(m[: (n // 64) * 64].reshape(64, -1))
        np.testing.assert_array_equal(b.Array(b.findall(b.CuArray(M))), oracle.findall(M))


@pytest.mark.parametrize("dtype", ["int64", "float32", "int32"])
@pytest.mark.parametrize("shape,dims", [((300, 4097), 2), ((2048, 1500), 2), ((7, 1000, 9), 2), ((1, 5000, 3), 2),
                                        ((129, 129, 129), 3)])
def test_strided_lookback_scan(b, dtype, shape, dims):
    """accumulate(op, A; dims) with dims > 1 on few, long columns: per-column look-back chains."""
    A = rand(dtype, shape)
    check_scan(b, "add", A, dims=dims)
    check_scan(b, "add", A, dims=dims, exclusive=True, init=npdt(dtype).type(2))
    if dtype != "float32":
        check_scan(b, "max", A, dims=dims)
