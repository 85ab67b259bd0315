import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import b200arrays as b, oracle
rng = np.random.default_rng(0)
for dtype in (np.int64, np.float32, np.int32, np.uint64):
    for n in (4095, 4096, 8192, 8193, 16384, 24577, 100_003, (1 << 21) + 5, (1 << 24) + 7):
        a = rng.integers(-1000, 1000, size=n).astype(dtype) if np.dtype(dtype).kind != "f" else rng.random(n).astype(dtype)
        print("n", n, np.dtype(dtype).name, end=" ", flush=True)
        got = b.Array(b.accumulate(b.add, b.CuArray(a)))
        ref = np.cumsum(a.astype(np.float64)) if np.dtype(dtype).kind == "f" else np.cumsum(a, dtype=dtype)
        ok = np.allclose(got, ref, rtol=1e-4) if np.dtype(dtype).kind == "f" else np.array_equal(got, ref)
        print("ok" if ok else "MISMATCH", flush=True)
M = np.asfortranarray(rng.integers(-5, 5, size=(20000, 7)).astype(np.int64))
print("segments", end=" ", flush=True)
print(np.array_equal(b.Array(b.accumulate(b.add, b.CuArray(M), dims=1)), np.cumsum(M, axis=0)), flush=True)
