"""Builds the inputs and the closed-form expected outputs of tests/golden/reference_literals.json.
Expected values come from the FORMULA named in the JSON (never from the oracle or the library), so the same
cases can pin both: tests/test_oracle_golden.py (oracle, CPU) and tests/test_golden_gpu.py (CUDA path, GPU)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    return json.load(open(os.path.join(HERE, "reference_literals.json")))["cases"]


def _special(v):
    return {"nan": np.nan, "inf": np.inf, "-inf": -np.inf}.get(v, v) if isinstance(v, str) else v


def make_input(spec, n=None, shape=None):
    dt = np.dtype(spec["eltype"])
    if "values" in spec:
        return np.array([_special(v) for v in spec["values"]], dtype=dt)
    n = spec.get("n", n)
    pat = spec["pattern"]
    if pat == "ones":
        return np.ones(shape if shape is not None else n, dtype=dt)
    i = np.arange(1, n + 1, dtype=np.int64)
    if pat == "iota":                      # map(x -> T(x), 1:N)
        return i.astype(dt)
    if pat == "neg_iota":                  # map(x -> T(-x), 1:N)
        return (-i).astype(dt)
    if pat == "rev_iota_mod256":           # map(x -> x % UInt8, reverse(1:N))
        return (i[::-1] % 256).astype(dt)
    if pat == "mod7":
        return (i % 7).astype(dt)
    raise ValueError(pat)


def make_expected(case, n=None, shape=None):
    exp = case["expected"]
    if isinstance(exp, list):
        return np.array(exp)
    n = case["input"].get("n", n)
    dt = np.dtype(case["input"]["eltype"])
    pat = exp["pattern"]
    i = np.arange(1, (n or 0) + 1, dtype=np.int64)
    if pat == "iota":
        return i.astype(dt)
    if pat == "iota_desc":
        return i[::-1].astype(dt)
    if pat == "neg_iota_asc":              # -N, ..., -1
        return (-i[::-1]).astype(dt)
    if pat == "neg_iota":                  # -1, ..., -N (descending)
        return (-i).astype(dt)
    if pat == "mod256_sorted":             # value v occurs #{k in 1..N : k mod 256 == v} times
        counts = [(n - v) // 256 + (1 if v else 0) if n >= v else 0 for v in range(256)]
        counts = [n // 256 if v == 0 else (n - v) // 256 + 1 for v in range(256)]
        return np.repeat(np.arange(256), counts).astype(dt)
    if pat == "range":
        return np.arange(exp["lo"], exp["hi"] + 1).astype(dt)
    if pat == "mod7_perm":                 # stable: residue 0 (indices 7, 14, ...), then residue 1 (1, 8, ...), ...
        return np.concatenate([np.arange(r if r else 7, n + 1, 7) for r in range(7)]).astype(np.int64)
    if pat == "iota_plus":
        return i.astype(dt) + dt.type(exp["offset"])
    if pat == "triangular":
        return (i * (i + 1) // 2).astype(dt)
    if pat == "index_along_dims":          # ones along dims: the running count = index along that dimension
        d = case["dims"] - 1
        idx = np.arange(1, shape[d] + 1).reshape([-1 if a == d else 1 for a in range(len(shape))])
        return np.broadcast_to(idx, shape).astype(dt)
    raise ValueError(pat)
