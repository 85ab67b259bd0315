"""Correctness + timing sweep of keys-only sort! through the C ABI, with the hybrid-MSD control words read back.

usage: python tools/sort_check.py [log2n ...]   (default 27 28 30)
Checker: torch.sort on the device (a library used as a test oracle only).  Prints one line per case:
  n, dtype, distribution, rev | mode nne tiles2 max_seg (SortCtl, first words of the workspace) | ms, Gkeys/s | ok
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200arrays as b  # noqa: E402
from b200arrays import _lib, dtypes as dt  # noqa: E402

lib = _lib.load()
torch.cuda.set_device(0)
stream = torch.cuda.current_stream().cuda_stream


def gen(n, dist, seed=1):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    if dist == "uniform":
        return torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
    if dist == "narrow8":      # keys in 1/8 of the key space (a multi-GPU shard after the exchange)
        x = torch.randint(0, 2**29, (n,), device="cuda", dtype=torch.int32, generator=g)
        return x + (3 << 29)
    if dist == "lowbits":      # only 20 significant bits: heavy duplicates, top 16 bits nearly constant
        return torch.randint(0, 2**20, (n,), device="cuda", dtype=torch.int32, generator=g)
    if dist == "normalf":      # bit pattern of normal(0,1) floats
        return torch.randn(n, device="cuda", generator=g).view(torch.int32)
    if dist == "dup16":        # 65536 distinct values spread over the key space
        return torch.randint(0, 2**16, (n,), device="cuda", dtype=torch.int32, generator=g) * 65537
    if dist == "const":
        return torch.full((n,), 123456789, device="cuda", dtype=torch.int32)
    if dist == "hi16only":     # low 16 bits zero: every leaf segment is one byte-1 group of identical-low keys
        return torch.randint(-2**15, 2**15 - 1, (n,), device="cuda", dtype=torch.int32, generator=g) << 16
    if dist == "randf":        # bit pattern of uniform [0, 1) floats: half the keys share one top byte, 64 copies per value at 2^30
        return torch.rand(n, device="cuda", generator=g).view(torch.int32)
    if dist == "skew2":        # half the keys uniform, half inside one top byte
        x = torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
        y = torch.randint(0, 2**24, (n,), device="cuda", dtype=torch.int32, generator=g) + (0x5A << 24)
        return torch.where(x > 0, x, y)
    if dist == "bigval":       # 1024 distinct values over the key space, n / 1024 copies each (> 65535 copies from 2^27 keys)
        return (torch.randint(0, 1024, (n,), device="cuda", dtype=torch.int32, generator=g) * 4194301) ^ -2**31
    if dist == "sorted":
        return torch.arange(n, device="cuda", dtype=torch.int64).mul_(3).to(torch.int32)
    raise ValueError(dist)


def expected(src, e, rev):
    if e == dt.U32:
        r = torch.sort(src ^ -2**31, descending=rev).values ^ -2**31
    elif e == dt.I32:
        r = torch.sort(src, descending=rev).values
    else:  # F32: isless order: -Inf < ... < -0.0 < +0.0 < ... < Inf < NaN (all NaNs last; first under rev)
        f = src.view(torch.float32)
        key = src.clone()
        neg = key < 0
        key = torch.where(neg, ~key, key | -2**31) ^ -2**31      # signed-comparable total order
        nan = torch.isnan(f)
        key = torch.where(nan, torch.full_like(key, 2**31 - 1), key)
        idx = torch.sort(key, descending=rev, stable=True).indices
        r = src[idx]
    return r


def run(n, e, dist, rev, reps=3):
    src = gen(n, dist)
    keys = src.clone()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.b200_sort_workspace_bytes(e, -1, n, 1, ctypes.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    ws.fill_(0xA5)                       # the library must not rely on zeroed workspace
    ts = []
    for _ in range(reps):
        keys.copy_(src)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.b200_sort_keys(ws.data_ptr(), nbytes.value, keys.data_ptr(), e, n, 1, 1 if rev else 0, stream))
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ctl = ws[:16].view(torch.int32).tolist()
    want = expected(src, e, rev)
    if e == dt.F32:
        f, w = keys.view(torch.float32), want.view(torch.float32)
        nanok = bool((torch.isnan(f) == torch.isnan(w)).all())
        ok = nanok and bool((keys[~torch.isnan(f)] == want[~torch.isnan(w)]).all())
    else:
        ok = bool((keys == want).all())
    nbad = 0 if ok else int((keys != want).sum())
    print(f"n={n:>11d} {dt.NAMES[e]:4s} {dist:9s} rev={int(rev)} | mode={ctl[0]} nne={ctl[1]} tiles2={ctl[2]} max_seg={ctl[3]} "
          f"| min {min(ts):8.3f} ms {n / min(ts) / 1e6:7.1f} Gkeys/s | ok={ok} bad={nbad}", flush=True)
    del ws, keys, src, want
    return ok


def main():
    logs = [int(a) for a in sys.argv[1:]] or [27, 28, 30]
    all_ok = True
    if os.environ.get("SORT_CHECK_ONE"):            # listed distributions only, UInt32 (ncu launch lists, A/B timing)
        ok = True
        for dist in os.environ["SORT_CHECK_ONE"].split(","):          # "dist" (UInt32) or "f32:dist" / "i32:dist"
            e = dt.U32
            if ":" in dist:
                pre, dist = dist.split(":")
                e = {"f32": dt.F32, "i32": dt.I32, "u32": dt.U32}[pre]
            ok &= run(1 << logs[0], e, dist, False, reps=int(os.environ.get("SORT_CHECK_REPS", "2")))
        sys.exit(0 if ok else 1)
    for lg in logs:
        n = 1 << lg
        for dist in ("uniform", "narrow8", "lowbits", "normalf", "dup16", "const", "hi16only", "sorted", "randf", "skew2", "bigval"):
            all_ok &= run(n, dt.U32, dist, False)
        all_ok &= run(n, dt.U32, "uniform", True)
        all_ok &= run(n, dt.I32, "uniform", False)
        all_ok &= run(n, dt.F32, "uniform", False)
        all_ok &= run(n, dt.F32, "normalf", True)
        all_ok &= run(n, dt.F32, "randf", False)
        all_ok &= run(n, dt.F32, "randf", True)
        all_ok &= run(n - 12345, dt.U32, "uniform", False)
        all_ok &= run(n + 4097, dt.F32, "uniform", False) if lg < 30 else True
    print("ALL OK" if all_ok else "FAILURES", flush=True)
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
