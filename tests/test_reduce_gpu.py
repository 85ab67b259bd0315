"""Parity of b200_reduce (through the host mirror) against the CPU oracle.

Modelled on the reference's own tests: test/core/array.jl:619-628 (issue 919, inference),
:756-804 (large map reduce, #1169), :1018-1030 (issue 2595), test/core/device/bfloat16.jl:72-78,
and the GPUArrays reduction testsuite's (op x eltype x dims) matrix.
Bar: bit-exact for integers / Bool / max / min; floats within
4*log2(n)*eps(T)*sum|x| of the Float64 pairwise reference (BASELINE north_star).
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(0xB200)

FLOATS = ["float16", "bfloat16", "float32", "float64"]
INTS = ["int32", "int64", "uint32", "uint64"]


def npdt(name):
    return oracle.bfloat16 if name == "bfloat16" else np.dtype(name)


def rand(name, shape):
    d = npdt(name)
    if oracle.is_float(d):
        return RNG.uniform(-1, 1, size=shape).astype(np.float32).astype(d)
    if d == np.bool_:
        return RNG.random(size=shape) < 0.5
    ii = np.iinfo(d)
    return RNG.integers(ii.min, ii.max, size=shape, dtype=d, endpoint=True)


def check(b, f, op, A, dims, exact=None, init=None):
    """Run mapreduce on device and compare with the oracle."""
    fn = {"add_sum": b.add_sum, "mul_prod": b.mul_prod, "add": b.add, "mul": b.mul, "max": b.max_, "min": b.min_,
          "and": b.and_, "or": b.or_}[op]
    ff = {"identity": b.identity, "abs2": b.abs2}[f]
    oop = {"add_sum": "add", "mul_prod": "mul"}.get(op, op)
    out_dt = oracle.widen_add(A.dtype) if op in ("add_sum", "mul_prod") else A.dtype
    d = b.CuArray(A)
    R = b.mapreduce(ff, fn, d, dims=dims, init=init)
    ds = range(1, A.ndim + 1) if dims is None else ((dims,) if isinstance(dims, int) else dims)
    rdims = tuple(1 if (i + 1) in ds else s for i, s in enumerate(A.shape))
    ref = oracle.mapreducedim(f, oop, A, rdims, out_dt)
    got = np.asarray(R) if dims is None else b.Array(R)
    got = got.reshape(rdims)
    if oracle.is_float(out_dt) and oop in ("add", "mul"):
        if oop == "add":
            tol = oracle.sum_tolerance(f, A, rdims, out_dt)
            err = np.abs(got.astype(np.float64) - ref)
            assert np.all(err <= tol), f"max err {err.max()} tol {tol.min()}"
        else:
            np.testing.assert_allclose(got.astype(np.float64), ref, rtol=64 * oracle.eps(out_dt), atol=1e-30)
    else:
        assert got.dtype == np.dtype(out_dt)
        np.testing.assert_array_equal(got.astype(np.float64) if oracle.is_float(out_dt) else got,
                                      ref.astype(np.float64) if oracle.is_float(out_dt) else ref)


@pytest.mark.parametrize("dtype", FLOATS + INTS)
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1000, 4097, 65536 + 7, 1 << 20, (1 << 22) + 3])
def test_sum_1d(b, dtype, n):
    check(b, "identity", "add_sum", rand(dtype, (n,)), None)


@pytest.mark.parametrize("dtype", FLOATS + INTS)
@pytest.mark.parametrize("op", ["max", "min", "add", "mul_prod"])
def test_ops_1d(b, dtype, op):
    A = rand(dtype, (100003,))
    if op == "mul_prod" and oracle.is_float(npdt(dtype)):
        A = (1 + A.astype(np.float32) * 1e-3).astype(npdt(dtype))[:2000]
    check(b, "identity", op, A, None)


@pytest.mark.parametrize("dtype", INTS)
@pytest.mark.parametrize("op", ["and", "or"])
def test_bitwise(b, dtype, op):
    A = rand(dtype, (70001,))
    init = None
    if op == "and":
        A = A | np.uint64(0x0F0F0F0F).astype(npdt(dtype))
        # GPUArrays.neutral_element(&, T) is one(T), NOT neutral for integers: `reduce(&, ints)` without an explicit init
        # must stay on the reference path (SURVEY section 8 a1'); with the true neutral element (all ones) it is grafted
        with pytest.raises(b.NotGraftedError):
            b.mapreduce(b.identity, b.and_, b.CuArray(A))
        init = b.true_neutral(b.and_, b.dtypes.from_np(A.dtype))
    check(b, "identity", op, A, None, init=init)
    check(b, "identity", op, np.asfortranarray(A[:69993].reshape(7, 9999)), 2, init=init)


@pytest.mark.parametrize("dtype", FLOATS + INTS)
def test_sum_abs2(b, dtype):
    check(b, "abs2", "add_sum", rand(dtype, (50001,)), None)
    check(b, "abs2", "add", rand(dtype, (129, 65)), 1)


SMALL_INTS = ["int8", "uint8", "int16", "uint16"]


@pytest.mark.parametrize("dtype", SMALL_INTS)
def test_small_integers(b, dtype):
    """Int8/UInt8/Int16/UInt16: sum widens to (U)Int64 (Base.add_sum), reduce(+/*), maximum, minimum, &, | keep the
    element type and wrap; the contiguous sum runs on the IDP.4A / IDP.2A path."""
    for n in (1, 31, 4097, (1 << 22) + 5):
        A = rand(dtype, (n,))
        check(b, "identity", "add_sum", A, None)
        check(b, "identity", "add_sum", A[3:] if n > 3 else A, None)          # unaligned head
    A = rand(dtype, (70001,))
    for op in ("max", "min", "add", "mul", "and", "or", "mul_prod"):
        # integer `&` is grafted only with the TRUE neutral element as init (all ones), see test_bitwise
        check(b, "identity", op, A, None, init=b.true_neutral(b.and_, b.dtypes.from_np(A.dtype)) if op == "and" else None)
    check(b, "abs2", "add_sum", A, None)
    M = np.asfortranarray(A[:69993].reshape(7, 9999))
    for dims in (1, 2):
        for op in ("add_sum", "max", "add"):
            check(b, "identity", op, M, dims)
    M = np.asfortranarray(rand(dtype, (4099, 33)))
    for dims in (1, 2, (1, 2)):
        check(b, "identity", "add_sum", M, dims)


SHAPES_DIMS = [
    ((16, 16), 1), ((16, 16), 2), ((1000, 3), 1), ((1000, 3), 2), ((3, 1000), 1), ((3, 1000), 2),
    ((257, 129), 1), ((257, 129), 2), ((4096, 512), 1), ((4096, 512), 2), ((512, 4096), 2),
    ((32, 32, 32), 1), ((32, 32, 32), 2), ((32, 32, 32), 3), ((32, 32, 32), (1, 2)), ((32, 32, 32), (2, 3)),
    ((85, 132, 10), 1), ((1, 70, 50, 20), 3), ((5, 1, 7), 2), ((2, 3, 100000), 2), ((100000, 2), 2),
    ((7, 200000), 1), ((3, 5, 7, 11), (2, 3)),
    # perf/array.jl:8-9 "L" shapes and neighbours: few long columns (narrow column tiles, parallel final fold)
    ((3, 1_000_000), 1), ((3, 1_000_000), 2), ((2, 300_001), 2), ((5, 100_000), 2), ((33, 40_000), 2), ((2, 50_000, 3), 2),
]


@pytest.mark.parametrize("shape,dims", SHAPES_DIMS)
@pytest.mark.parametrize("dtype", ["float32", "int64", "bfloat16", "int32", "float64"])
def test_sum_dims(b, shape, dims, dtype):
    check(b, "identity", "add_sum", np.asfortranarray(rand(dtype, shape)), dims)


@pytest.mark.parametrize("shape,dims", SHAPES_DIMS[:12])
@pytest.mark.parametrize("op", ["max", "min"])
@pytest.mark.parametrize("dtype", ["float32", "float16", "uint32"])
def test_maxmin_dims(b, shape, dims, op, dtype):
    check(b, "identity", op, np.asfortranarray(rand(dtype, shape)), dims)


def test_max_nan_and_signed_zero(b):
    A = rand("float32", (100000,))
    A[77777] = np.nan
    assert np.isnan(b.maximum(b.CuArray(A)))
    assert np.isnan(b.minimum(b.CuArray(A)))
    Z = np.array([-0.0, 0.0, -0.0], dtype=np.float32)
    assert not np.signbit(b.maximum(b.CuArray(Z)))
    assert np.signbit(b.minimum(b.CuArray(Z)))
    Z64 = np.array([-0.0, 0.0, -0.0, np.nan], dtype=np.float64)
    assert np.isnan(b.maximum(b.CuArray(Z64)))
    assert np.signbit(b.minimum(b.CuArray(Z64[:3])))
    M = np.asfortranarray(rand("float32", (300, 200)))
    M[5, 7] = np.nan
    got = b.Array(b.maximum(b.CuArray(M), dims=1)).reshape(-1)
    assert np.isnan(got[7]) and not np.isnan(np.delete(got, 7)).any()


def test_count_any_all(b):
    M = np.asfortranarray(RNG.random((1000, 777)) < 0.5)
    d = b.CuArray(M)
    assert b.count(d) == M.sum()
    np.testing.assert_array_equal(b.Array(b.count(d, dims=1)), M.sum(axis=0, keepdims=True).astype(np.int64))
    np.testing.assert_array_equal(b.Array(b.count(d, dims=2)), M.sum(axis=1, keepdims=True).astype(np.int64))
    assert b.any(d) is True and b.all(d) is False
    T = np.ones(5000, dtype=bool)
    assert b.all(b.CuArray(T)) is True
    T[4999] = False
    assert b.all(b.CuArray(T)) is False
    assert b.any(b.CuArray(np.zeros(5000, dtype=bool))) is False
    np.testing.assert_array_equal(b.Array(b.any(d, dims=2)), M.any(axis=1, keepdims=True))
    np.testing.assert_array_equal(b.Array(b.all(d, dims=1)), M.all(axis=0, keepdims=True))


@pytest.mark.parametrize("dtype", ["float32", "int32", "bfloat16", "float64"])
def test_extrema(b, dtype):
    A = rand(dtype, (123457,))
    mn, mx = b.extrema(b.CuArray(A))
    assert mn == A.min() and mx == A.max()
    M = np.asfortranarray(rand(dtype, (300, 50)))
    R = b.Array(b.extrema(b.CuArray(M), dims=1))
    np.testing.assert_array_equal(R[0].astype(np.float64), M.min(axis=0, keepdims=True).astype(np.float64))
    np.testing.assert_array_equal(R[1].astype(np.float64), M.max(axis=0, keepdims=True).astype(np.float64))


def test_large_map_reduce_serial_regime(b):
    """test/core/array.jl:756-788: (threshold+5) x 31, `==` on Float32 sum => index-order accumulation."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    big = 1024 * sms + 5
    a = RNG.random((big, 31), dtype=np.float32)
    a = np.asfortranarray(a)
    c = b.CuArray(a)
    np.testing.assert_array_equal(b.Array(b.minimum(c, dims=2)), a.min(axis=1, keepdims=True))
    exp = oracle.mapreducedim("identity", "add", a, (big, 1), np.float32, exact_order=True)
    np.testing.assert_array_equal(b.Array(b.sum(c, dims=2)), exp)
    # same regime with the reduced dimension leading (pre == 1)
    at = np.asfortranarray(a.T.copy())
    exp_t = oracle.mapreducedim("identity", "add", at, (1, big), np.float32, exact_order=True)
    np.testing.assert_array_equal(b.Array(b.sum(b.CuArray(at), dims=1)), exp_t)
    ai = np.asfortranarray(RNG.integers(-2**62, 2**62, size=(big, 31), dtype=np.int64))
    ci = b.CuArray(ai)
    np.testing.assert_array_equal(b.Array(b.sum(ci, dims=2)), ai.sum(axis=1, keepdims=True))
    np.testing.assert_array_equal(b.Array(b.minimum(ci, dims=2)), ai.min(axis=1, keepdims=True))


@pytest.mark.parametrize("shape", [(32, 32, 32), (256, 256, 256), (85, 1320, 100)])
def test_issue_1169_sum_bang(b, shape):
    """test/core/array.jl:790-804: sum!(R(1,Ny,Nz), A) on Float64 randn, `≈`."""
    A = np.asfortranarray(RNG.standard_normal(shape))
    R = b.CuArray(np.zeros((1,) + shape[1:]))
    out = b.sum_(R, b.CuArray(A))
    assert out is R
    np.testing.assert_allclose(b.Array(R), A.sum(axis=0, keepdims=True), rtol=1e-12, atol=1e-12)


def test_reference_literal_pins(b):
    # issue 919 (test/core/array.jl:619-622): sum!(zeros(1,1) view, ones(1,4096)) == [4096f0]
    R = b.CuArray(np.zeros((1, 1), dtype=np.float32))
    b.sum_(R, b.CuArray(np.ones((1, 4096), dtype=np.float32)))
    assert b.Array(R).reshape(-1).tolist() == [4096.0]
    # issue 2595 (test/core/array.jl:1018-1024): sum!(Float32 zeros(1), Float64 ones(2)) == [2f0]
    a = b.CuArray(np.zeros(1, dtype=np.float32))
    b.sum_(a, b.CuArray(np.ones(2, dtype=np.float64)))
    assert b.Array(a).tolist() == [2.0] and b.Array(a).dtype == np.float32
    # BFloat16 reductions (test/core/device/bfloat16.jl:72-78)
    x = b.CuArray(np.array([1, 2, 3, 4], dtype=np.float32).astype(oracle.bfloat16))
    assert float(b.sum(x)) == 10.0 and float(b.prod(x)) == 24.0 and float(b.sum(b.abs2, x)) == 30.0
    # mapreduce inference test (array.jl:624-628): returns the very R it was given
    inp = b.CuArray(np.ones(512, dtype=np.float32))
    out = inp.similar(dims=(1,))
    assert b.mapreducedim_(b.identity, b.add, out, inp, init=np.float32(0)) is out
    assert b.Array(out).tolist() == [512.0]


def test_init_from_out_folds_once(b):
    A = rand("int64", (1 << 21,))
    R = b.CuArray(np.array([5], dtype=np.int64))
    b.mapreducedim_(b.identity, b.add, R, b.CuArray(A), init=None)
    with np.errstate(over="ignore"):
        assert b.Array(R)[0] == np.int64(5) + A.sum(dtype=np.int64)


def test_empty_and_errors(b):
    E = b.CuArray(np.zeros((0,), dtype=np.float32))
    assert b.sum(E) == 0.0
    R = b.CuArray(np.array([7.0], dtype=np.float32))
    assert b.mapreducedim_(b.identity, b.add, R, E, init=np.float32(0)) is R
    assert b.Array(R)[0] == 7.0                       # untouched (mapreduce.jl:175)
    with pytest.raises(b.NotGraftedError):
        b.mapreducedim_(lambda x: x, b.add, R, E)
    with pytest.raises(b.NotGraftedError):
        b.mapreducedim_(b.identity, b.add, R, b.CuArray(np.ones(4, dtype=np.float32)), init=np.float32(3))
    with pytest.raises(b.DimensionMismatch):
        b.mapreducedim_(b.identity, b.add, b.CuArray(np.zeros(3, dtype=np.float32)),
                        b.CuArray(np.ones(4, dtype=np.float32)), init=np.float32(0))


def test_unaligned_views(b):
    """Pointers offset by an odd number of elements (contiguous views stay CuArrays, array.jl:379-386)."""
    import torch
    base = rand("float32", (100003,))
    d = b.CuArray(base)
    for off in (1, 3, 5):
        v = b.CuArray(d.buf[off * 4:], (base.size - off,), d.eltype)
        ref = base[off:].astype(np.float64).sum()
        assert abs(float(b.sum(v)) - ref) <= 1e-3
    h = rand("bfloat16", (9, 4001))
    dh = b.CuArray(np.asfortranarray(h))
    v = b.CuArray(dh.buf[2 * 9:], (9, 4000), dh.eltype)      # drop first column: 18-byte offset
    got = b.Array(b.sum(v, dims=2)).astype(np.float64).reshape(-1)
    ref = h[:, 1:].astype(np.float64).sum(axis=1)
    assert np.all(np.abs(got - ref) <= 0.05 * np.abs(h[:, 1:].astype(np.float64)).sum(axis=1))


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16", "bfloat16", "int32", "int64", "uint32", "uint64"])
def test_findmax_findmin(b, dtype):
    """findmax / findmin (test/core/array.jl:767-769,782-784): first extremal element wins."""
    a = rand(dtype, (100_003,))
    a[::1000] = a[7]                        # ties: the first one must win
    for is_max, fn in ((True, b.findmax), (False, b.findmin)):
        v, i = fn(b.CuArray(a))
        rv, ri = oracle.findminmax(a, None, is_max)
        assert i == ri and np.asarray(v).reshape(1).tobytes() == np.asarray(rv).reshape(1).tobytes()
    for shape, dims in (((300, 257), 1), ((300, 257), 2), ((33, 65, 17), 2), ((5, 40_000), 2), ((40_000, 5), 1), ((2, 3, 50_000), 3)):
        A = np.asfortranarray(rand(dtype, shape))
        for is_max, fn in ((True, b.findmax), (False, b.findmin)):
            V, I = fn(b.CuArray(A), dims=dims)
            rv, ri = oracle.findminmax(A, dims, is_max)
            np.testing.assert_array_equal(b.Array(I), ri)
            assert np.ascontiguousarray(b.Array(V)).tobytes() == np.ascontiguousarray(rv).tobytes()
    assert b.argmax(b.CuArray(a)) == oracle.findminmax(a, None, True)[1]


def test_findmax_nan_and_zero_rules(b):
    a = np.array([1.0, -0.0, 0.0, np.nan, 5.0, np.nan, -7.0], dtype=np.float32)
    assert b.findmax(b.CuArray(a))[1] == 4 and np.isnan(b.findmax(b.CuArray(a))[0])      # first NaN
    assert b.findmin(b.CuArray(a))[1] == 4 and np.isnan(b.findmin(b.CuArray(a))[0])
    z = np.array([0.0, -0.0, 0.0, -0.0], dtype=np.float64)
    assert b.findmax(b.CuArray(z))[1] == 1 and b.findmin(b.CuArray(z))[1] == 2
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    big = 1024 * sms + 5                                                                    # "large map reduce" shape
    A = np.asfortranarray(RNG.random((big, 31), dtype=np.float32))
    V, I = b.findmax(b.CuArray(A), dims=2)
    rv, ri = oracle.findminmax(A, 2, True)
    np.testing.assert_array_equal(b.Array(V), rv)
    np.testing.assert_array_equal(b.Array(I), ri)
