"""Device-timed microbenchmarks of single library calls (CUDA events, L2 flushed by
input size or an explicit flush).  Usage: python tools/microbench.py [case ...]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200arrays as b  # noqa: E402
from b200arrays import dtypes as dt  # noqa: E402

PEAK = 6550.1


def timeit(fn, reps=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def report(name, nbytes, tmin, tmed):
    print(f"{name:44s} min {tmin:8.4f} ms  med {tmed:8.4f} ms  {nbytes / tmin / 1e6:8.1f} GB/s "
          f"({nbytes / tmin / 1e6 / PEAK * 100:5.1f}% of measured copy peak, {nbytes / tmin / 1e6 / 8000 * 100:5.1f}% of 8 TB/s)",
          flush=True)


def torch_dtype(e):
    from b200arrays.cuarray import _DT_TO_TORCH
    return _DT_TO_TORCH[e]


def rand_array(e, dims):
    n = int(np.prod(dims))
    td = torch_dtype(e)
    if e in dt.FLOATS:
        t = torch.rand(n, device="cuda", dtype=torch.float32).mul_(2).sub_(1).to(td)
    elif e == dt.BOOL:
        t = torch.rand(n, device="cuda") < 0.5
    elif e in (dt.U32, dt.U64):
        t = torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", dtype=torch.int32 if e == dt.U32 else torch.int64).view(td)
    else:
        t = torch.randint(-2**20, 2**20, (n,), device="cuda", dtype=td)
    return b.CuArray.from_torch(t, dims, e)


def reduce_cases():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for e, name in ((dt.F32, "F32"), (dt.BF16, "BF16")):
        A = rand_array(e, (16384, 16384))
        nb = A.nbytes
        for dims in (1, 2):
            R = b.CuArray.undef(e, (1, 16384) if dims == 1 else (16384, 1))
            report(f"sum dims={dims} 16384^2 {name}", nb,
                   *timeit(lambda: b.mapreducedim_(b.identity, b.add_sum, R, A, init=0), flush=flush))
        R1 = b.CuArray.undef(e, (1, 1))
        report(f"maximum 16384^2 {name}", nb, *timeit(lambda: b.mapreducedim_(b.identity, b.max_, R1, A, init=-np.inf), flush=flush))
        del A
    M = rand_array(dt.BOOL, (16384, 16384))
    Rc = b.CuArray.undef(dt.I64, (1, 1))
    report("count 16384^2 Bool", M.nbytes, *timeit(lambda: b.mapreducedim_(b.identity, b.add_sum, Rc, M, init=0), flush=flush))
    del M
    V = rand_array(dt.F32, (1 << 30,))
    Rs = b.CuArray.undef(dt.F32, (1,))
    report("sum 2^30 F32", V.nbytes, *timeit(lambda: b.mapreducedim_(b.identity, b.add_sum, Rs, V, init=0)))
    Vt = V.typed()
    report("  torch.sum 2^30 F32 (library comparator)", V.nbytes, *timeit(lambda: torch.sum(Vt)))
    del V, Vt
    H = rand_array(dt.BF16, (1 << 30,))
    Rh = b.CuArray.undef(dt.F32, (1,))
    report("sum(abs2) 2^30 BF16 -> F32", H.nbytes, *timeit(lambda: b.mapreducedim_(b.abs2, b.add, Rh, H, init=0)))


def scan_cases():
    for e, name in ((dt.F32, "F32"), (dt.I64, "I64")):
        V = rand_array(e, (1 << 30,))
        O = b.CuArray.undef(e, (1 << 30,))
        report(f"cumsum 2^30 {name}", 2 * V.nbytes, *timeit(lambda: b.accumulate_(b.add, O, V), reps=10))
        Vt, Ot = V.typed(), O.typed()
        report(f"  torch.cumsum 2^30 {name} (library comparator)", 2 * V.nbytes, *timeit(lambda: torch.cumsum(Vt, 0, out=Ot), reps=5))
        del V, O, Vt, Ot


def scan_dims_cases():
    for e, name in ((dt.F32, "F32"), (dt.I64, "I64")):
        A = rand_array(e, (16384, 16384 if e == dt.F32 else 8192))
        O = b.CuArray.undef(e, A.dims)
        for d in (1, 2):
            report(f"cumsum dims={d} {A.dims} {name}", 2 * A.nbytes, *timeit(lambda: b.accumulate_(b.add, O, A, dims=d), reps=5))
        del A, O


def sort_cases():
    n = 1 << 30
    for e, name in ((dt.U32, "U32"), (dt.F32, "F32")):
        src = rand_array(e, (n,))
        K = src.copy()
        def run():
            K.buf.copy_(src.buf)
            b.sort_(K)
        def base():
            K.buf.copy_(src.buf)
        t0 = timeit(base, reps=5)[0]
        tmin, tmed = timeit(run, reps=5)
        tmin, tmed = tmin - t0, tmed - t0
        print(f"sort! 2^30 {name}: {tmin:.3f} ms (med {tmed:.3f})  {n / tmin / 1e6:.1f} Gkeys/s  "
              f"{2 * 4 * n * 4 / tmin / 1e6:.0f} GB/s algorithmic ({2 * 4 * n * 4 / tmin / 1e6 / 8000 * 100:.1f}% of 8 TB/s roofline)", flush=True)
        tp = timeit(lambda: b.sortperm(src), reps=3)
        print(f"sortperm 2^30 {name}: {tp[0]:.3f} ms  {n / tp[0] / 1e6:.1f} Gkeys/s", flush=True)
        if e == dt.F32:
            st = src.typed()
            tt = timeit(lambda: torch.sort(st), reps=3)
            print(f"  torch.sort 2^30 F32 (library comparator): {tt[0]:.3f} ms  {n / tt[0] / 1e6:.1f} Gkeys/s", flush=True)
        del src, K


if __name__ == "__main__":
    which = sys.argv[1:] or ["reduce"]
    torch.cuda.set_device(0)
    for w in which:
        {"reduce": reduce_cases, "scan": scan_cases, "scan_dims": scan_dims_cases, "sort": sort_cases}[w]()
