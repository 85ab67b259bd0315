"""CPU tests of the oracle: (1) the reference's literal-valued test expectations
(tests/golden/reference_literals.json, each citing file:line), (2) committed oracle regression
vectors, (3) semantic properties the reference's tests rely on, (4) the C restatement
(oracle/oracle_cpu.c) against the numpy oracle."""
import json
import os

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def _input(spec):
    dt = oracle.bfloat16 if spec["eltype"] == "bfloat16" else np.dtype(spec["eltype"])
    if "values" in spec:
        return np.array(spec["values"], dtype=np.float32).astype(dt)
    return np.full(spec["dims"], spec["fill"], dtype=dt)


def test_reference_literal_pins():
    cases = json.load(open(os.path.join(HERE, "golden", "reference_literals.json")))["cases"]
    checked = 0
    for c in cases:
        if c["call"] in ("sum!", "mapreducedim"):
            A = _input(c["input"])
            odt = np.dtype(c.get("out_eltype", c["input"]["eltype"]))
            rd = tuple(c["out_dims"]) + (1,) * (A.ndim - len(c["out_dims"]))
            R = oracle.mapreducedim("identity", "add", A, rd, odt)
            assert np.asarray(R).astype(odt).reshape(-1).tolist() == c["expected"], c["name"]
            checked += 1
        elif c["call"] in ("sum", "prod", "sum_abs2"):
            A = _input(c["input"])
            f, op = {"sum": ("identity", "add"), "prod": ("identity", "mul"), "sum_abs2": ("abs2", "add")}[c["call"]]
            R = oracle.mapreducedim(f, op, A, (1,), A.dtype)
            assert float(np.asarray(R).reshape(-1)[0]) == c["expected"][0], c["name"]
            checked += 1
    assert checked == 6


def test_reference_deterministic_cases_pin_the_oracle():
    """Sort / partialsort / sortperm / accumulate on the reference tests' own deterministic inputs (and constructed
    tie / NaN / signed-zero vectors): expected outputs are closed forms built in tests/golden/cases.py."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import cases as G
    n_checked = 0
    for c in G.load():
        call = c["call"]
        if call == "sort":
            a = G.make_input(c["input"])
            np.testing.assert_array_equal(oracle.sort(a, rev=c["rev"]), G.make_expected(c), err_msg=c["name"])
        elif call == "partialsort!":
            a = G.make_input(c["input"])
            srt = oracle.sort(a)
            k = c["k"]
            got = srt[k - 1:k] if isinstance(k, int) else srt[k[0] - 1:k[1]]
            np.testing.assert_array_equal(got, G.make_expected(c), err_msg=c["name"])
        elif call == "sortperm":
            a = G.make_input(c["input"])
            np.testing.assert_array_equal(oracle.sortperm(a, rev=c["rev"]), G.make_expected(c), err_msg=c["name"])
        elif call in ("accumulate", "accumulate_init"):
            for n in c["sizes"]:
                a = G.make_input(c["input"], n=n)
                init = a.dtype.type(c["init"]) if "init" in c else None
                np.testing.assert_array_equal(oracle.accumulate("add", a, 1, init), G.make_expected(c, n=n), err_msg=c["name"])
        elif call == "accumulate_cols":
            for n in c["sizes"]:
                a = G.make_input(c["input"], shape=(n, 2))
                got = oracle.accumulate("add", a.reshape(-1, order="F"), 1).reshape(a.shape, order="F")   # A[:] branch
                for col in range(2):
                    np.testing.assert_array_equal(got[:, col], np.arange(1, n + 1) + col * n, err_msg=c["name"])
        elif call == "accumulate_dims":
            shape = tuple(c["shape"])
            a = G.make_input(c["input"], shape=shape)
            np.testing.assert_array_equal(oracle.accumulate("add", a, c["dims"]), G.make_expected(c, shape=shape), err_msg=c["name"])
        else:
            continue
        n_checked += 1
    assert n_checked == 25


def test_oracle_regression_vectors():
    g = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
    a, f, i64, u8 = g["i32"], g["f32"], g["i64"], g["u8"]
    np.testing.assert_array_equal(oracle.mapreducedim("identity", "add", a, (1, 23), np.int64), g["i32_sum_d1"])
    np.testing.assert_array_equal(oracle.mapreducedim("identity", "add", a, (37, 1), np.int64), g["i32_sum_d2"])
    np.testing.assert_array_equal(oracle.mapreducedim("identity", "add", a, (37, 1), np.int32), g["i32_plus_d2_wrap"])
    np.testing.assert_array_equal(oracle.mapreducedim("abs2", "add", a, (1, 1), np.int64), g["i32_abs2_wrap"])
    np.testing.assert_array_equal(oracle.mapreducedim("identity", "max", a, (1, 23), np.int32), g["i32_max_d1"])
    np.testing.assert_array_equal(oracle.sort(f).view(np.uint32), g["f32_sort"].view(np.uint32))
    np.testing.assert_array_equal(oracle.sort(f, rev=True).view(np.uint32), g["f32_sort_rev"].view(np.uint32))
    np.testing.assert_array_equal(oracle.sortperm(f), g["f32_sortperm"])
    np.testing.assert_array_equal(oracle.sortperm(f, rev=True), g["f32_sortperm_rev"])
    np.testing.assert_array_equal(oracle.accumulate("add", i64, 2), g["i64_cumsum_d2"])
    np.testing.assert_array_equal(oracle.accumulate("add", i64, 2, init=np.int64(12345)), g["i64_cumsum_d2_init"])
    np.testing.assert_array_equal(oracle.accumulate("max", i64, 3), g["i64_cummax_d3"])
    np.testing.assert_array_equal(oracle.sortperm(u8), g["u8_sortperm"])


def test_semantics_julia_base():
    # add_sum widening (SURVEY §8 a1'): Int32 -> Int64, UInt32 -> UInt64, Bool -> Int64; `+` wraps in T
    assert oracle.widen_add(np.int32) == np.int64 and oracle.widen_add(np.uint32) == np.uint64
    assert oracle.widen_add(np.bool_) == np.int64 and oracle.widen_add(np.float32) == np.float32
    a = np.array([2**31 - 1, 1], dtype=np.int32)
    assert oracle.mapreducedim("identity", "add", a, (1,), np.int32)[0] == np.int32(-2**31)
    assert oracle.mapreducedim("identity", "add", a, (1,), np.int64)[0] == 2**31
    # abs2 of Int32 wraps in Int32 before widening
    b = np.array([65536], dtype=np.int32)
    assert oracle.mapreducedim("abs2", "add", b, (1,), np.int64)[0] == 0
    # max/min: NaN-propagating, -0.0 < +0.0
    z = np.array([-0.0, 0.0], dtype=np.float32)
    assert not np.signbit(oracle.mapreducedim("identity", "max", z, (1,), np.float32)[0])
    assert np.signbit(oracle.mapreducedim("identity", "min", z, (1,), np.float32)[0])
    assert np.isnan(oracle.mapreducedim("identity", "max", np.array([1.0, np.nan]), (1,), np.float64)[0])
    # isless order and stable sortperm with rev (sorting.jl:582-592)
    v = np.array([1.0, np.nan, -0.0, 0.0, -np.inf, 1.0], dtype=np.float64)
    assert oracle.sortperm(v).tolist() == [5, 3, 4, 1, 6, 2]
    assert oracle.sortperm(v, rev=True).tolist() == [2, 1, 6, 4, 3, 5]
    # scan: init folded once into every element (accumulate.jl:90-92); dims > ndims copies (:140)
    x = np.array([1, 2, 3], dtype=np.int64)
    assert oracle.accumulate("add", x, 1, init=np.int64(10)).tolist() == [11, 13, 16]
    assert oracle.accumulate("add", x, 2).tolist() == [1, 2, 3]
    assert oracle.accumulate("add", x, 1, exclusive=True).tolist() == [0, 1, 3]
    # the order pin: sequential Float32 accumulation differs from the Float64 reference in general
    r = np.random.default_rng(1).random((50, 31)).astype(np.float32)
    seq = oracle.mapreducedim("identity", "add", r, (50, 1), np.float32, exact_order=True)
    acc = np.zeros(50, dtype=np.float32)
    for j in range(31):
        acc = (acc + r[:, j]).astype(np.float32)
    np.testing.assert_array_equal(seq.reshape(-1), acc)
    assert oracle.findall(np.array([[True, False], [True, True]])).tolist() == [1, 2, 4]


def test_c_restatement_matches_numpy_oracle():
    import oracle_cpu as oc
    rng = np.random.default_rng(7)
    a = rng.random(300_001).astype(np.float32)
    ref = a.astype(np.float64).sum()
    assert abs(float(oc.sum_f32(a)) - ref) <= float(oracle.sum_tolerance("identity", a, (1,), np.float32)[0])
    A = np.asfortranarray(rng.uniform(-1, 1, size=(513, 257)).astype(np.float32))
    for d, rd in ((1, (1, 257)), (2, (513, 1))):
        got = oc.sum_dims(A, d).astype(np.float64)
        assert np.all(np.abs(got - oracle.mapreducedim("identity", "add", A, rd, np.float32))
                      <= oracle.sum_tolerance("identity", A, rd, np.float32))
    np.testing.assert_array_equal(oc.sum_dims(A, 2), oracle.mapreducedim("identity", "add", A, (513, 1), np.float32, exact_order=True))
    assert oc.maximum_f32(A) == A.max()
    H = np.asfortranarray(A.astype(oracle.bfloat16))
    assert oc.maximum_bf16(H.view(np.uint16)) == np.float32(H.astype(np.float32).max())
    hb = oc.sum_dims(np.asfortranarray(H.view(np.uint16)), 1).view(oracle.bfloat16).astype(np.float64)
    refb = oracle.mapreducedim("identity", "add", H, (1, 257), oracle.bfloat16)
    assert np.all(np.abs(hb - refb) <= 513 * 2.0 ** -8 * np.abs(H.astype(np.float64)).sum(axis=0, keepdims=True))
    Bm = rng.random(10_001) < 0.5
    assert oc.count_bool(Bm) == int(Bm.sum())
    c = oc.cumsum(a).astype(np.float64)
    assert np.all(np.abs(c - np.cumsum(a.astype(np.float64))) <= oracle.scan_tolerance(a, 1, np.float32))
    i = rng.integers(-2**62, 2**62, size=100_000, dtype=np.int64)
    np.testing.assert_array_equal(oc.cumsum(i), oracle.accumulate("add", i))
    f = rng.standard_normal(100_000).astype(np.float32)
    f[[5, 7, 9]] = [np.nan, -0.0, 0.0]
    np.testing.assert_array_equal(oc.sort(f).view(np.uint32), oracle.sort(f).view(np.uint32))
    np.testing.assert_array_equal(oc.sortperm(f), oracle.sortperm(f))
    u = rng.integers(0, 1000, size=100_000, dtype=np.uint32)
    np.testing.assert_array_equal(oc.sort(u), np.sort(u))
    np.testing.assert_array_equal(oc.sortperm(u), oracle.sortperm(u))
