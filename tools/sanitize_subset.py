"""A small, fast pass over every kernel family for compute-sanitizer (memcheck / racecheck), SURVEY section 5:
    compute-sanitizer --tool memcheck  python tools/sanitize_subset.py
    compute-sanitizer --tool racecheck python tools/sanitize_subset.py
Sizes are tiny (the tools slow kernels down 10-100x); the hybrid MSD sort is forced onto a 2^18-key input through the
tuning build's knobs (B200_MSD_MIN_LOG2N / B200_MSD_MIN_AVG), so every look-back / ticket / atomic-rank kernel runs.
Results are checked against numpy so a silent corruption also fails the run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200arrays as b  # noqa: E402
import oracle  # noqa: E402

torch.cuda.set_device(0)
rng = np.random.default_rng(3)
ok = True


def check(name, cond):
    global ok
    ok = ok and bool(cond)
    print(f"{name:50s} {'ok' if cond else 'FAIL'}", flush=True)


u = rng.integers(0, 2**32, size=(1 << 18) + 77, dtype=np.uint64).astype(np.uint32)
check("sort! UInt32 2^18 (hybrid forced in the tuning build)", np.array_equal(b.Array(b.sort(b.CuArray(u))), np.sort(u)))
srt = (np.arange(u.size, dtype=np.uint64) * 16000).astype(np.uint32)
check("sort! presorted UInt32 (uniform-row paths)", np.array_equal(b.Array(b.sort(b.CuArray(srt))), srt))
dup = (rng.integers(0, 2**16, size=u.size, dtype=np.uint32) * np.uint32(65537))
check("sort! heavy duplicates (16-bit-bin leaf)", np.array_equal(b.Array(b.sort(b.CuArray(dup))), np.sort(dup)))
rf = rng.random((1 << 18) + 5, dtype=np.float32)            # uniform [0, 1): half the keys share one top byte -> heavy-digit ranking
check("sort! Float32 uniform [0,1) (heavy digits, 16-bit leaf with B200_MSD_CAP8)", np.array_equal(b.Array(b.sort(b.CuArray(rf))), np.sort(rf)))
big = (rng.integers(0, 64, size=1 << 23, dtype=np.uint32) * np.uint32(67108859))      # 64 values x 2^17 copies: spill path of the 16-bit leaf
check("sort! 64 values x 131072 copies (spill table)", np.array_equal(b.Array(b.sort(b.CuArray(big))), np.sort(big)))
del big
ps = rng.standard_normal((1 << 22) + 3).astype(np.float32)
check("partialsort (sampled bracket) 2^22", b.partialsort(b.CuArray(ps), 1 << 21) == np.sort(ps)[(1 << 21) - 1])
del ps
f = rng.standard_normal(200_003).astype(np.float32)
f[[3, 4, 5]] = [np.nan, -0.0, 0.0]
check("sort! Float32 (LSD passes)", np.array_equal(b.Array(b.sort(b.CuArray(f))).view(np.uint32), oracle.sort(f).view(np.uint32)))
check("sortperm Float32", np.array_equal(b.Array(b.sortperm(b.CuArray(f))), oracle.sortperm(f)))
i64 = rng.integers(-2**40, 2**40, size=300_001, dtype=np.int64)
check("sort! Int64", np.array_equal(b.Array(b.sort(b.CuArray(i64))), np.sort(i64)))
check("cumsum Int64 (chain2)", np.array_equal(b.Array(b.cumsum(b.CuArray(i64))), np.cumsum(i64)))
g = rng.random(500_000).astype(np.float32)
cs = b.Array(b.cumsum(b.CuArray(g))).astype(np.float64)
check("cumsum Float32", np.all(np.abs(cs - np.cumsum(g.astype(np.float64))) <= oracle.scan_tolerance(g, 1, np.float32)))
M = np.asfortranarray(rng.integers(-1000, 1000, size=(513, 257), dtype=np.int64))
check("cumsum dims=2 (strided look-back)", np.array_equal(b.Array(b.cumsum(b.CuArray(M), dims=2)), np.cumsum(M, axis=1)))
check("sum dims=1", np.array_equal(b.Array(b.sum(b.CuArray(M), dims=1)), M.sum(axis=0, keepdims=True)))
check("sum dims=2", np.array_equal(b.Array(b.sum(b.CuArray(M), dims=2)), M.sum(axis=1, keepdims=True)))
check("maximum", b.maximum(b.CuArray(g)) == g.max())
check("extrema", tuple(b.extrema(b.CuArray(i64))) == (i64.min(), i64.max()))
check("findmax", tuple(b.findmax(b.CuArray(g))) == (g.max(), int(g.argmax()) + 1))
flags = rng.random(100_001) < 0.3
check("findall", np.array_equal(b.Array(b.findall(b.CuArray(flags))), np.flatnonzero(flags) + 1))
check("partialsort (radix select)", b.partialsort(b.CuArray(f[~np.isnan(f)]), 1234) == np.sort(f[~np.isnan(f)])[1233])
torch.cuda.synchronize()
print("ALL OK" if ok else "FAILURES", flush=True)
sys.exit(0 if ok else 1)
