"""Parity of the onesweep radix sort (through the host mirror of sorting.jl's Base hooks)
against the CPU oracle.  Mirrors test/core/sorting.jl:285-403:
  * sort!/sort validity = same multiset + issorted (check_equivalence, :169-171); here the
    stronger element-wise equality with the oracle's stable sort is asserted as well;
  * sortperm == Base.sortperm exactly (:248-253), incl. rev, Float32/Float64 1e6, Int32/UInt64
    index eltypes, empty input, N-d dims=1;
  * sizes 1,2,3,4, 2^16 + {0,1,31,32,33,127,128,129} (:351-362); Int8 4e6 (:307);
  * partialsort k scalar and ranges (:368-372); ArgumentError on mismatched axes (:401).
Bit-exact everywhere.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(0x50F7)


def npdt(name):
    return oracle.bfloat16 if name == "bfloat16" else np.dtype(name)


def rand(name, n):
    d = npdt(name)
    if oracle.is_float(d):
        return RNG.random(n).astype(np.float32).astype(d) if d.itemsize < 8 else RNG.random(n)
    if d == np.bool_:
        return RNG.random(n) < 0.5
    ii = np.iinfo(d)
    return RNG.integers(ii.min, ii.max, size=n, dtype=d, endpoint=True)


def bits(a):
    a = np.asarray(a)
    return a.view(np.dtype(f"u{a.dtype.itemsize}")) if a.dtype != np.bool_ else a.astype(np.uint8)


ALL_TYPES = ["int64", "int32", "int8", "float64", "float32", "float16", "bfloat16", "uint8", "uint16", "int16",
             "uint32", "uint64", "bool"]


@pytest.mark.parametrize("dtype", ALL_TYPES)
@pytest.mark.parametrize("rev", [False, True])
def test_sort_types(b, dtype, rev):
    a = rand(dtype, 10000)
    c = b.CuArray(a)
    out = b.sort_(c, rev=rev)
    assert out is c
    got = b.Array(c)
    assert oracle.is_valid_sort(a, got, rev=rev)
    np.testing.assert_array_equal(bits(got), bits(oracle.sort(a, rev=rev)))
    s = b.sort(b.CuArray(a), rev=rev)
    np.testing.assert_array_equal(bits(b.Array(s)), bits(oracle.sort(a, rev=rev)))


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, (1 << 16), (1 << 16) + 1, (1 << 16) + 31, (1 << 16) + 32, (1 << 16) + 33,
                               (1 << 16) + 127, (1 << 16) + 128, (1 << 16) + 129, 4095, 4096, 4097])
def test_sort_sizes(b, n):
    a = rand("float32", n)
    got = b.Array(b.sort_(b.CuArray(a)))
    np.testing.assert_array_equal(bits(got), bits(oracle.sort(a)))


@pytest.mark.parametrize("dtype,n,f", [("int32", 1_000_000, "sorted"), ("float32", 1_000_000, "sorted"),
                                       ("int32", 1_000_000, "reverse"), ("float32", 1_000_000, "reverse"),
                                       ("int8", 4_000_000, "random"), ("uint8", 100_000, "skew"),
                                       ("float32", 1 << 22, "random"), ("int64", 1_000_000, "random"),
                                       ("float64", 1_000_000, "random")])
@pytest.mark.parametrize("rev", [False, True])
def test_sort_orderings(b, dtype, n, f, rev):
    d = npdt(dtype)
    if f == "sorted":
        a = np.arange(1, n + 1).astype(d)
    elif f == "reverse":
        a = (-np.arange(1, n + 1)).astype(d)
    elif f == "skew":
        a = np.round(255 * RNG.random(n) ** 3).astype(d)
    else:
        a = rand(dtype, n)
    got = b.Array(b.sort_(b.CuArray(a), rev=rev))
    assert oracle.is_valid_sort(a, got, rev=rev)
    np.testing.assert_array_equal(bits(got), bits(oracle.sort(a, rev=rev)))


def test_float_total_order(b):
    """isless: -Inf < ... < -0.0 < +0.0 < ... < +Inf < NaN."""
    special = np.array([np.nan, -np.nan, np.inf, -np.inf, 0.0, -0.0, 1.0, -1.0, 5e-324, -5e-324], dtype=np.float64)
    for dtype in ("float64", "float32", "float16", "bfloat16"):
        a = np.concatenate([special.astype(np.float32 if dtype != "float64" else np.float64), RNG.standard_normal(5000)]).astype(npdt(dtype))
        RNG.shuffle(a)
        for rev in (False, True):
            got = b.Array(b.sort_(b.CuArray(a), rev=rev))
            assert oracle.is_valid_sort(a, got, rev=rev)
            ref = oracle.sort(a, rev=rev)
            nan = np.isnan(ref.astype(np.float64))
            np.testing.assert_array_equal(np.isnan(got.astype(np.float64)), nan)
            np.testing.assert_array_equal(bits(got)[~nan], bits(ref)[~nan])
            p = b.Array(b.sortperm(b.CuArray(a), rev=rev))
            np.testing.assert_array_equal(p, oracle.sortperm(a, rev=rev))


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("rev", [False, True])
def test_sortperm_1m(b, dtype, rev):
    """test/core/sorting.jl:381-387: 1e6 Float32s have ~9.4e5 unique values, stability is non-trivial."""
    a = rand(dtype, 1_000_000)
    I = b.sortperm(b.CuArray(a), rev=rev)
    assert I.eltype == b.dtypes.I64
    np.testing.assert_array_equal(b.Array(I), oracle.sortperm(a, rev=rev))


@pytest.mark.parametrize("dtype", ["int32", "uint8", "int8", "bool", "uint16", "float16", "int64", "uint64"])
def test_sortperm_stability_few_unique(b, dtype):
    a = rand(dtype, 300_001)
    if dtype not in ("bool", "uint8", "int8"):
        a = (a.astype(np.float64) % 17).astype(npdt(dtype))
    for rev in (False, True):
        np.testing.assert_array_equal(b.Array(b.sortperm(b.CuArray(a), rev=rev)), oracle.sortperm(a, rev=rev))


def test_sortperm_empty_and_small(b):
    assert b.Array(b.sortperm(b.CuArray(np.zeros(0, dtype=np.float32)))).size == 0
    for n in (1, 2, 3, 33):
        a = rand("float32", n)
        np.testing.assert_array_equal(b.Array(b.sortperm(b.CuArray(a))), oracle.sortperm(a))


def test_sortperm_bang_index_eltypes(b):
    """test/core/sorting.jl:395-403."""
    a = rand("float32", 1_000_000)
    ix = b.CuArray(np.arange(1, a.size + 1, dtype=np.int32))
    out = b.sortperm_(ix, b.CuArray(a))
    assert out is ix
    np.testing.assert_array_equal(b.Array(ix), oracle.sortperm(a).astype(np.int32))
    ix0 = b.CuArray(np.zeros(0, dtype=np.int32))
    assert b.sortperm_(ix0, b.CuArray(np.zeros(0, dtype=np.float32))) is ix0
    ai = rand("int64", 1_000_000)
    ixu = b.CuArray(np.arange(1, ai.size + 1, dtype=np.uint64))
    b.sortperm_(ixu, b.CuArray(ai), initialized=False)
    np.testing.assert_array_equal(b.Array(ixu), oracle.sortperm(ai).astype(np.uint64))
    with pytest.raises(b.ArgumentError) as e:
        b.sortperm_(b.CuArray(np.arange(1, 4)), b.CuArray(np.arange(1, 5)))
    assert "index array must have the same size/axes as the source array" in str(e.value)
    with pytest.raises(b.NotGraftedError):
        b.sortperm_(b.CuArray(np.arange(1, 5)), b.CuArray(np.arange(1, 5)), initialized=True)
    # keys are never modified (sorting.jl:857-858)
    d = b.CuArray(a)
    b.sortperm(d)
    np.testing.assert_array_equal(b.Array(d), a)


def test_partialsort(b):
    """test/core/sorting.jl:368-372."""
    a = rand("int64", 100_000)
    ref = oracle.sort(a)
    for k in (1, 100_000, 50_000):
        c = b.CuArray(a)
        assert b.partialsort_(c, k) == ref[k - 1]
        assert oracle.is_valid_sort(a, b.Array(c))
    for k in (range(10_000, 20_001), range(1, 100_001)):
        c = b.CuArray(a)
        out = b.partialsort_(c, k)
        assert out.dims == (len(k),)
        np.testing.assert_array_equal(b.Array(out), ref[k[0] - 1:k[-1]])
    np.testing.assert_array_equal(b.Array(b.partialsort(b.CuArray(a), range(5, 10), rev=True)), oracle.sort(a, rev=True)[4:9])


def test_dims1_segments(b):
    """sort!(A; dims=1) / sortperm(A; dims=1) (test/core/sorting.jl:390,392 shapes, scaled)."""
    A = np.asfortranarray(rand("float32", 10_000 * 16).reshape(10_000, 16))
    c = b.CuArray(A)
    b.sort_(c, dims=1)
    np.testing.assert_array_equal(b.Array(c), oracle.sort_dims(A, 1))
    np.testing.assert_array_equal(b.Array(b.sortperm(b.CuArray(A), dims=1)), oracle.sortperm_dims(A, 1))
    B3 = np.asfortranarray(rand("int32", 100 * 16 * 8).reshape(100, 16, 8))
    c3 = b.CuArray(B3)
    b.sort_(c3, dims=1, rev=True)
    np.testing.assert_array_equal(b.Array(c3), oracle.sort_dims(B3, 1, rev=True))
    np.testing.assert_array_equal(b.Array(b.sortperm(b.CuArray(B3), dims=1)), oracle.sortperm_dims(B3, 1))


def test_guards(b):
    c = b.CuArray(rand("float32", 100))
    with pytest.raises(b.NotGraftedError):
        b.sort_(c, by=lambda x: abs(x - 0.5))
    with pytest.raises(b.NotGraftedError):
        b.sort_(c, alg=b.QuickSort)
    with pytest.raises(b.ArgumentError):
        b.sort_(b.CuArray(np.zeros((4, 4), dtype=np.float32)), dims=3)


def test_repetition_no_races(b):
    """test/core/sorting.jl:273-278: repeat to hunt races."""
    a = (np.arange(100_000, 0, -1) % 256).astype(np.uint8)
    ref = oracle.sort(a)
    for _ in range(10):
        np.testing.assert_array_equal(b.Array(b.sort_(b.CuArray(a))), ref)


def test_large_sort_and_sortperm(b):
    n = (1 << 24) + 12345
    a = rand("uint32", n)
    got = b.Array(b.sort_(b.CuArray(a)))
    np.testing.assert_array_equal(got, np.sort(a))
    f = rand("float32", n)
    np.testing.assert_array_equal(b.Array(b.sortperm(b.CuArray(f))), oracle.sortperm(f))


@pytest.mark.parametrize("shape,dims,dtype,rev", [
    ((4, 50000, 4), 2, "int32", False),          # test/core/sorting.jl:314
    ((2, 2, 50000), 3, "int32", True),           # :315
    ((100_000, 4), 1, "float32", False), ((4, 100_000), 2, "float32", False),     # :326-327 shapes
    ((100, 16), 1, "float32", False), ((100, 16), 2, "float32", True), ((100, 256, 64), 1, "float32", False),
    ((3, 5000, 7), 2, "float64", False), ((33, 4096, 3), 2, "uint16", True), ((7, 4097, 2), 2, "int8", False),
    ((5, 3, 1000), 2, "int64", False), ((1, 1, 9), 3, "float32", False),
])
def test_sort_any_dims(b, shape, dims, dtype, rev):
    """sort!(A; dims) / sortperm(A; dims) along any dimension."""
    A = np.asfortranarray(rand(dtype, int(np.prod(shape))).reshape(shape))
    c = b.CuArray(A)
    assert b.sort_(c, dims=dims, rev=rev) is c
    np.testing.assert_array_equal(bits(b.Array(c)), bits(oracle.sort_dims(A, dims, rev=rev)))
    got = b.Array(b.sortperm(b.CuArray(A), dims=dims, rev=rev))
    np.testing.assert_array_equal(got, oracle.sortperm_dims(A, dims, rev=rev) if not rev else _sortperm_dims_rev(A, dims))


def _sortperm_dims_rev(A, dims):
    ax = dims - 1
    lin = (np.arange(A.size, dtype=np.int64) + 1).reshape(A.shape, order="F")
    M, L = np.moveaxis(A, ax, -1), np.moveaxis(lin, ax, -1)
    out = np.empty(M.shape, dtype=np.int64)
    for idx in np.ndindex(M.shape[:-1]):
        out[idx] = L[idx][oracle.sortperm(M[idx], rev=True) - 1]
    return np.moveaxis(out, -1, ax)


def test_sortperm_dims_reference_shapes(b):
    """test/core/sorting.jl:388-392: (100_000, 16) dims=1 and dims=2, (100, 256, 256) dims=1 (scaled)."""
    A = np.asfortranarray(rand("float32", 100_000 * 16).reshape(100_000, 16))
    for d in (1, 2):
        np.testing.assert_array_equal(b.Array(b.sortperm(b.CuArray(A), dims=d)), oracle.sortperm_dims(A, d))
    B3 = np.asfortranarray(rand("float32", 100 * 256 * 32).reshape(100, 256, 32))
    np.testing.assert_array_equal(b.Array(b.sortperm(b.CuArray(B3), dims=1)), oracle.sortperm_dims(B3, 1))


@pytest.mark.parametrize("n", [5, 100, 1000, 4095, 4096])
@pytest.mark.parametrize("dtype", ["float32", "int64", "uint8", "float16"])
def test_small_block_sort(b, n, dtype):
    a = rand(dtype, n)
    np.testing.assert_array_equal(bits(b.Array(b.sort_(b.CuArray(a)))), bits(oracle.sort(a)))
    np.testing.assert_array_equal(b.Array(b.sortperm(b.CuArray(a), rev=True)), oracle.sortperm(a, rev=True))


@pytest.mark.parametrize("dtype", ["float32", "uint32", "int64", "float16", "uint8"])
@pytest.mark.parametrize("nb", [1, 2, 8, 16])
@pytest.mark.parametrize("rev", [False, True])
def test_partition_by_splitters(b, dtype, nb, rev):
    """b200_partition: stable multi-way partition used by the multi-GPU sample sort."""
    ops = b.multigpu.DeviceOps()
    for n in (0, 5, 4096, 100_003):
        a = rand(dtype, n)
        if dtype.startswith("float") and n > 10:
            a[3] = np.nan
        spl = oracle.sort(rand(dtype, nb - 1), rev=rev) if nb > 1 else np.zeros(0, dtype=a.dtype)
        pay = np.arange(n, dtype=np.int64)
        kout, pout, counts = ops.partition(b.CuArray(a), b.CuArray(pay), b.CuArray(spl), nb, rev)
        ke, se = oracle.total_order_key(a), oracle.total_order_key(spl)
        if rev:
            ke, se = ~ke, ~se
        bid = np.searchsorted(se, ke, side="right") if nb > 1 else np.zeros(n, dtype=np.int64)
        order = np.argsort(bid, kind="stable")
        np.testing.assert_array_equal(b.Array(counts), np.bincount(bid, minlength=nb).astype(np.int64))
        np.testing.assert_array_equal(b.Array(pout), pay[order])
        np.testing.assert_array_equal(bits(b.Array(kout)), bits(a[order]))
        k2, p2, c2 = ops.partition(b.CuArray(a), None, b.CuArray(spl), nb, rev)
        assert p2 is None
        np.testing.assert_array_equal(bits(b.Array(k2)), bits(a[order]))


@pytest.mark.parametrize("dtype", ["uint32", "float32", "int64", "uint8"])
@pytest.mark.parametrize("nb", [2, 8, 16])
def test_partition_two_phase_scatter(b, dtype, nb):
    """b200_partition_count + b200_partition_scatter: every bucket goes to its own destination buffer at a
    caller-chosen element offset (the multi-GPU sort points these at peer receive buffers)."""
    ops = b.multigpu.DeviceOps()
    for n in (1, 4097, 250_007):
        a = rand(dtype, n)
        spl = oracle.sort(rand(dtype, nb - 1))
        pay = np.arange(n, dtype=np.int64)
        ka, sa = b.CuArray(a), b.CuArray(spl)
        counts, ws = ops.partition_count(ka, sa, nb, False)
        ke, se = oracle.total_order_key(a), oracle.total_order_key(spl)
        bid = np.searchsorted(se, ke, side="right")
        want = np.bincount(bid, minlength=nb).astype(np.int64)
        np.testing.assert_array_equal(b.Array(counts), want)
        base = [3 * i + 1 for i in range(nb)]
        dk = [b.CuArray(np.zeros(int(want[i]) + base[i] + 2, dtype=a.dtype)) for i in range(nb)]
        dp = [b.CuArray(np.full(int(want[i]) + base[i] + 2, -1, dtype=np.int64)) for i in range(nb)]
        ops.partition_scatter(ws, ka, b.CuArray(pay), sa, nb, False, counts,
                              [d.pointer for d in dk], [d.pointer for d in dp], base)
        for i in range(nb):
            sel = np.nonzero(bid == i)[0]
            gotk, gotp = b.Array(dk[i]), b.Array(dp[i])
            np.testing.assert_array_equal(bits(gotk[base[i]:base[i] + len(sel)]), bits(a[sel]))
            np.testing.assert_array_equal(gotp[base[i]:base[i] + len(sel)], pay[sel])
            assert (gotp[:base[i]] == -1).all() and (gotp[base[i] + len(sel):] == -1).all()   # nothing outside the slot


@pytest.mark.parametrize("dtype", ["float32", "uint32", "int64", "float64", "float16", "uint8", "int16"])
@pytest.mark.parametrize("rev", [False, True])
def test_partialsort_radix_select(b, dtype, rev):
    """partialsort(c, k) for scalar k through b200_select_kth == sort(c)[k] (the reference: full sort, then
    c[k], sorting.jl:1015-1022); the input is not modified."""
    for n in (1, 2, 1000, 4097, 300_003):
        a = rand(dtype, n)
        if dtype.startswith("float") and n > 10:
            a[1] = np.nan; a[2] = -0.0; a[3] = 0.0; a[4] = np.inf; a[5] = -np.inf
            a[n // 2:n // 2 + n // 8] = a[0]                      # a block of duplicates
        ref = oracle.sort(a, rev=rev)
        d = b.CuArray(a)
        for k in sorted({min(k, n) for k in (1, 2, n // 3 + 1, n // 2 + 1, max(n - 1, 1), n)}):
            got = b.partialsort(d, k, rev=rev)
            np.testing.assert_array_equal(bits(np.asarray([got], dtype=a.dtype)), bits(ref[k - 1:k]))
        np.testing.assert_array_equal(bits(b.Array(d)), bits(a))
    with pytest.raises(IndexError):
        b.partialsort(b.CuArray(rand(dtype, 5)), 6)


@pytest.mark.parametrize("dtype", ["float32", "uint32", "int64", "float64", "int16"])
@pytest.mark.parametrize("case", ["random", "dups", "constant", "sorted"])
def test_partialsort_sampled_path(b, dtype, case):
    """n >= 2^22: b200_select_kth brackets rank k from a sorted sample, compacts the bracket in ONE read and finishes on the
    candidates; a bracket that misses (or overflows its buffer: constant input) falls back to the streaming passes on the
    device.  Same answer as sort(c)[k] either way, input untouched."""
    n = (1 << 22) + 12345
    a = rand(dtype, n)
    if case == "dups":
        a[: n // 2] = a[7]                                        # half the array is one value: bracket = N/2 candidates
    elif case == "constant":
        a[:] = a[3]
    elif case == "sorted":
        a = np.sort(a)
    if dtype.startswith("float") and case == "random":
        a[1] = np.nan; a[2] = -0.0; a[3] = 0.0; a[4] = np.inf; a[5] = -np.inf
    d = b.CuArray(a)
    for rev in (False, True):
        ref = oracle.sort(a, rev=rev)
        for k in (1, 2, n // 1000, n // 3 + 1, n // 2, n - 1, n):
            got = b.partialsort(d, k, rev=rev)
            np.testing.assert_array_equal(bits(np.asarray([got], dtype=a.dtype)), bits(ref[k - 1:k]))
    np.testing.assert_array_equal(bits(b.Array(d)), bits(a))
