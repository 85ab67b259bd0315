"""The CUDA path against the reference-held deterministic cases of tests/golden/reference_literals.json
(sort!, partialsort!, sortperm, accumulate): expected outputs are closed forms built in tests/golden/cases.py,
never the oracle's or the library's own output.  Bit-exact.  The same cases pin the oracle on the CPU in
tests/test_oracle_golden.py::test_reference_deterministic_cases_pin_the_oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import cases as G  # noqa: E402

CASES = [c for c in G.load() if c["call"] in ("sort", "partialsort!", "sortperm", "sortperm_error", "accumulate", "accumulate_init",
                                              "accumulate_cols", "accumulate_dims")]


@pytest.mark.parametrize("case", CASES, ids=[c["name"][:60] for c in CASES])
def test_reference_deterministic_case(b, case):
    c, call = case, case["call"]
    if call == "sort":
        d = b.CuArray(G.make_input(c["input"]))
        assert b.sort_(d, rev=c["rev"]) is d                                    # sort! returns its argument
        np.testing.assert_array_equal(b.Array(d), G.make_expected(c))
    elif call == "partialsort!":
        d = b.CuArray(G.make_input(c["input"]))
        k = c["k"]
        out = b.partialsort_(d, k if isinstance(k, int) else range(k[0], k[1] + 1))
        got = np.atleast_1d(out if isinstance(k, int) else b.Array(out))
        np.testing.assert_array_equal(got, G.make_expected(c))
        np.testing.assert_array_equal(b.Array(d), np.sort(G.make_input(c["input"])))   # partialsort! leaves c sorted
    elif call == "sortperm":
        a = G.make_input(c["input"])
        ix = b.sortperm(b.CuArray(a), rev=c["rev"])
        assert ix.eltype == b.dtypes.I64
        np.testing.assert_array_equal(b.Array(ix), G.make_expected(c))
        if a.size:                                                               # Int32 / UInt64 index eltypes (sorting.jl:395-403)
            for et in (b.dtypes.I32, b.dtypes.U64):
                ixe = b.sortperm_(b.CuArray.undef(et, a.shape), b.CuArray(a), rev=c["rev"])
                np.testing.assert_array_equal(b.Array(ixe).astype(np.int64), G.make_expected(c))
    elif call == "sortperm_error":
        with pytest.raises(b.ArgumentError) as e:
            b.sortperm_(b.CuArray(np.arange(1, c["ix_len"] + 1)), b.CuArray(np.arange(1, c["a_len"] + 1)))
        assert str(e.value) == c["expected_text"]
    elif call in ("accumulate", "accumulate_init"):
        for n in c["sizes"]:
            a = G.make_input(c["input"], n=n)
            init = a.dtype.type(c["init"]) if "init" in c else None
            got = b.Array(b.accumulate(b.add, b.CuArray(a), init=init))
            np.testing.assert_array_equal(got, G.make_expected(c, n=n))
    elif call == "accumulate_cols":
        for n in c["sizes"]:
            a = np.asfortranarray(G.make_input(c["input"], shape=(n, 2)))
            got = b.Array(b.accumulate(b.add, b.CuArray(a)))                     # no dims: A[:] flattening branch
            for col in range(2):
                np.testing.assert_array_equal(got[:, col], np.arange(1, n + 1) + col * n)
    elif call == "accumulate_dims":
        shape = tuple(c["shape"])
        a = np.asfortranarray(G.make_input(c["input"], shape=shape))
        got = b.Array(b.accumulate(b.add, b.CuArray(a), dims=c["dims"]))
        np.testing.assert_array_equal(got, G.make_expected(c, shape=shape))
