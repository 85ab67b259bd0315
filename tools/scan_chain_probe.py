"""Correctness + timing of the contiguous scan (look-back vs self-chained variants).
usage: scan_chain_probe.py [check] [time]   (B200ARRAYS_LIB / B200_SCAN_CHAIN select the variant)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import b200arrays as b
from b200arrays import dtypes as dt
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
tag = os.path.basename(os.environ.get("B200ARRAYS_LIB", "default")) + " CHAIN=" + os.environ.get("B200_SCAN_CHAIN", "1")
modes = sys.argv[1:] or ["check", "time"]

if "check" in modes:
    g = torch.Generator(device=dev); g.manual_seed(5)
    bad = 0
    for n in (1, 5, 4096, 16384, 16385, 65536 * 3 + 1, (1 << 22) + 3, (1 << 26) + 12345, 1 << 28):
        I = torch.randint(-2**20, 2**20, (n,), device=dev, generator=g, dtype=torch.int64)
        got = b.accumulate(b.add, b.CuArray.from_torch(I)).buf.view(torch.int64)[:n]
        ok = bool(torch.equal(got, torch.cumsum(I, 0)))
        F = torch.rand(n, device=dev, generator=g, dtype=torch.float32)
        gotf = b.accumulate(b.add, b.CuArray.from_torch(F)).buf.view(torch.float32)[:n]
        ref = torch.cumsum(F.double(), 0)
        err = float(((gotf.double() - ref).abs() / ref.abs().clamp_min(1e-30)).max())
        I32 = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, generator=g, dtype=torch.int32)
        got32 = b.accumulate(b.add, b.CuArray.from_torch(I32)).buf.view(torch.int32)[:n]
        ok32 = bool(torch.equal(got32, torch.cumsum(I32, 0, dtype=torch.int32)))
        okf = err < 4 * max(np.log2(max(n, 2)), 1) * 2.0 ** -23
        bad += (not ok) + (not ok32) + (not okf)
        print(f"[{tag}] n={n}: int64 {'ok' if ok else 'MISMATCH'} int32 {'ok' if ok32 else 'MISMATCH'} f32 relerr {err:.2e} {'ok' if okf else 'BAD'}", flush=True)
    # segmented (post > 1) with several tiles per segment, exclusive, init
    for (rows, cols) in ((50_001, 37), (16384, 64), (200_000, 3), (16384 * 5, 300)):
        M = torch.randint(-1000, 1000, (cols, rows), device=dev, generator=g, dtype=torch.int64)     # column-major rows x cols
        cM = b.CuArray.from_torch(M, (rows, cols))
        got = b.accumulate(b.add, cM, dims=1).buf.view(torch.int64).reshape(cols, rows)
        ok = bool(torch.equal(got, torch.cumsum(M, 1)))
        got2 = b.accumulate(b.add, cM, dims=1, init=7).buf.view(torch.int64).reshape(cols, rows)
        ok2 = bool(torch.equal(got2, torch.cumsum(M, 1) + 7))
        bad += (not ok) + (not ok2)
        print(f"[{tag}] segments {rows}x{cols}: {'ok' if ok else 'MISMATCH'} init {'ok' if ok2 else 'MISMATCH'}", flush=True)
    print(f"[{tag}] CHECK {'PASSED' if bad == 0 else 'FAILED'}", flush=True)

if "time" in modes:
    def tmed(fn, reps):
        ts = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
        return float(np.median(ts)), float(min(ts))
    for lg in [int(x) for x in os.environ.get("PROBE_LG", "28,30").split(",")]:
        n = 1 << lg
        V = torch.rand(n, device=dev, dtype=torch.float32)
        cV = b.CuArray.from_torch(V); O = b.CuArray.undef(dt.F32, (n,))
        fn = lambda: b.accumulate_(b.add, O, cV)
        fn(); fn(); med, mn = tmed(fn, 7)
        print(f"[{tag}] cumsum 2^{lg} F32: med {med:.3f} ms min {mn:.3f} ms  {8 * n / med / 1e6:.0f} GB/s", flush=True)
        del V, cV, O; torch.cuda.empty_cache()
        I = torch.randint(-2**20, 2**20, (n,), device=dev, dtype=torch.int64)
        cI = b.CuArray.from_torch(I); O = b.CuArray.undef(dt.I64, (n,))
        fn = lambda: b.accumulate_(b.add, O, cI)
        fn(); fn(); med, mn = tmed(fn, 7)
        print(f"[{tag}] cumsum 2^{lg} I64: med {med:.3f} ms min {mn:.3f} ms  {16 * n / med / 1e6:.0f} GB/s", flush=True)
        del I, cI, O; torch.cuda.empty_cache()
