"""Regenerates tests/golden/oracle_vectors.npz: small seeded inputs with the outputs of
oracle/oracle.py.  These pin the ORACLE against accidental change (they are oracle outputs,
not reference outputs: the reference is Julia and cannot run here — its literal-valued pins
are tests/golden/reference_literals.json).  Usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import oracle  # noqa: E402

rng = np.random.default_rng(0xB200)
out = {}
a = rng.integers(-2**31, 2**31 - 1, size=(37, 23), dtype=np.int32)
out["i32"] = a
out["i32_sum_d1"] = oracle.mapreducedim("identity", "add", a, (1, 23), np.int64)
out["i32_sum_d2"] = oracle.mapreducedim("identity", "add", a, (37, 1), np.int64)
out["i32_plus_d2_wrap"] = oracle.mapreducedim("identity", "add", a, (37, 1), np.int32)
out["i32_abs2_wrap"] = oracle.mapreducedim("abs2", "add", a, (1, 1), np.int64)
out["i32_max_d1"] = oracle.mapreducedim("identity", "max", a, (1, 23), np.int32)
f = rng.standard_normal(1001).astype(np.float32)
f[[3, 500]] = [np.nan, -0.0]
f[7] = 0.0
f[9] = np.inf
f[11] = -np.inf
out["f32"] = f
out["f32_sort"] = oracle.sort(f)
out["f32_sort_rev"] = oracle.sort(f, rev=True)
out["f32_sortperm"] = oracle.sortperm(f)
out["f32_sortperm_rev"] = oracle.sortperm(f, rev=True)
i64 = rng.integers(-2**62, 2**62, size=(5, 70, 6), dtype=np.int64)
out["i64"] = i64
out["i64_cumsum_d2"] = oracle.accumulate("add", i64, 2)
out["i64_cumsum_d2_init"] = oracle.accumulate("add", i64, 2, init=np.int64(12345))
out["i64_cummax_d3"] = oracle.accumulate("max", i64, 3)
few = rng.integers(0, 5, size=300).astype(np.uint8)
out["u8"] = few
out["u8_sortperm"] = oracle.sortperm(few)
np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
print("wrote", len(out), "arrays")
