"""The reduce / scan / sort / findall groups of the reference's perf/array.jl (same shapes and group names,
perf/array.jl:5-22,37-41,48-51,72-125,149-153) through the Python mirror; torch as library comparator.
Small, launch-latency-dominated shapes: 512x1000 and 3x1_000_000.  Prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import b200arrays as b

torch.cuda.set_device(0)
rng = np.random.default_rng(123)
m, n, m_long, n_long = 512, 1000, 3, 1_000_000
mat = b.CuArray(np.asfortranarray(rng.random((m, n), dtype=np.float32)))
mat_long = b.CuArray(np.asfortranarray(rng.random((m_long, n_long), dtype=np.float32)))
mat_i = b.CuArray(np.asfortranarray(rng.integers(-10, 11, (m, n), dtype=np.int64)))
mat_long_i = b.CuArray(np.asfortranarray(rng.integers(-10, 11, (m_long, n_long), dtype=np.int64)))
bools = b.CuArray(rng.random(m * n) < 0.5)


def t_(a):          # torch view with the same memory (column-major dims -> reversed row-major shape)
    return a.typed().view(*reversed(a.dims)) if len(a.dims) > 1 else a.typed()


def bench(fn, reps=200):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize()
    reps = max(3, min(reps, int(20.0 / max(e0.elapsed_time(e1), 1e-3))))     # <= ~20 ms per sample
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return round(min(ts), 2)


def bench_graph(fn, calls=20):
    """Device-side time per call: `calls` invocations captured in one CUDA graph and replayed (the library is
    capturable; this removes the Python mirror's ~30 us of host work per call from the figure)."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / calls)
    return round(min(ts), 2)


_bench_host = bench


def bench(fn, reps=200, graph=True):      # [host-driven us, graph-replayed us] for our side, host-driven for torch
    return _bench_host(fn, reps)


def pair(fn_ours, reps_or_fn, fn_torch=None, reps=200):
    if fn_torch is None:
        fn_torch = reps_or_fn
    else:
        reps = reps_or_fn
    return {"b200arrays": [_bench_host(fn_ours, reps), bench_graph(fn_ours)],
            "torch": [_bench_host(fn_torch, reps), bench_graph(fn_torch)]}


out = {}
for name, M, L in (("Float32", mat, mat_long), ("Int64", mat_i, mat_long_i)):
    v = M.vec()
    tm, tl = t_(M), t_(L)
    acc = out.setdefault("accumulate", {}).setdefault(name, {})
    acc["1d"] = pair(lambda: b.accumulate(b.add, v), lambda: torch.cumsum(tm.reshape(-1), 0))
    acc["dims=1"] = pair(lambda: b.accumulate(b.add, M, dims=1), lambda: torch.cumsum(tm, 1))
    acc["dims=2"] = pair(lambda: b.accumulate(b.add, M, dims=2), lambda: torch.cumsum(tm, 0))
    acc["dims=1L"] = pair(lambda: b.accumulate(b.add, L, dims=1), lambda: torch.cumsum(tl, 1))
    acc["dims=2L"] = pair(lambda: b.accumulate(b.add, L, dims=2), lambda: torch.cumsum(tl, 0))
    red = out.setdefault("reduce", {}).setdefault(name, {})
    red["1d"] = pair(lambda: b.mapreduce(b.identity, b.add, v, dims=1), lambda: tm.sum())
    red["dims=1"] = pair(lambda: b.mapreduce(b.identity, b.add, M, dims=1), lambda: tm.sum(1))
    red["dims=2"] = pair(lambda: b.mapreduce(b.identity, b.add, M, dims=2), lambda: tm.sum(0))
    red["dims=1L"] = pair(lambda: b.mapreduce(b.identity, b.add, L, dims=1), lambda: tl.sum(1))
    red["dims=2L"] = pair(lambda: b.mapreduce(b.identity, b.add, L, dims=2), lambda: tl.sum(0))
v = mat.vec()
tv = t_(mat).reshape(-1)
out["sorting"] = {
    "1d": pair(lambda: b.sort(v), 50, lambda: torch.sort(tv)),
    "2d": pair(lambda: b.sort(mat, dims=1), 50, lambda: torch.sort(t_(mat), dim=1)),
    "sortperm": pair(lambda: b.sortperm(v), 50, lambda: torch.argsort(tv, stable=True)),
}
tb = bools.typed()
out["findall"] = {"bool": [bench(lambda: b.findall(bools), 50), bench(lambda: torch.nonzero(tb), 50)]}
out["findmin"] = {"1d": [bench(lambda: b.findmin(v), 50), bench(lambda: torch.min(tv, 0), 50)],
                  "2d": [bench(lambda: b.findmin(mat, dims=1), 50), bench(lambda: torch.min(t_(mat), 1), 50)]}
print(json.dumps({"unit": "us per call, [host-driven loop, CUDA-graph replay = device time]; findall/findmin: [b200arrays, torch] host-driven only (they read a scalar back)", "shapes": "512x1000, 3x1000000 (perf/array.jl)", "results": out}))
