"""The kernels of the device-driven multi-GPU sort (csrc/mgpu_sort.cu) on ONE GPU: G ranks are emulated in one
process through the C ABI — b200_mgpu_hist16 per shard, the G histograms concatenated (what the all-gather
delivers), b200_mgpu_plan per rank, b200_mgpu_scatter of every shard into G ordinary device buffers standing in for
the peers' receive buffers, b200_sort_keys_from per rank — and the concatenation of the ranks' slices is compared
bit for bit with the oracle's sort of the whole input.  (The NVLink peer mapping itself is exercised by bench.py
--gpus N and tests/test_multigpu_gpu.py on multi-GPU boxes; the host logic by tests/test_multigpu_cpu.py.)"""
import ctypes

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def emulate(b, full, G, rev, cuts=None):
    lib = b._lib.load()
    dt = b.dtypes
    et = dt.from_np(full.dtype)
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream().cuda_stream
    cuts = np.linspace(0, full.size, G + 1).astype(int) if cuts is None else cuts
    shards = [b.CuArray(full[cuts[r]:cuts[r + 1]]) for r in range(G)]
    hist_all = torch.empty(G * 65536, dtype=torch.int32, device=dev)
    for r, s in enumerate(shards):
        b._lib.check(lib.b200_mgpu_hist16(s.pointer if s.length else None, et, s.length, int(rev), 0,
                                          hist_all[r * 65536:].data_ptr(), stream))
    nb = ctypes.c_size_t(0)
    b._lib.check(lib.b200_mgpu_plan_bytes(ctypes.byref(nb)))
    plans, hist_recv, infos = [], [], []
    for r in range(G):
        plan = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        hr = torch.empty(65536, dtype=torch.int32, device=dev)
        b._lib.check(lib.b200_mgpu_plan(hist_all.data_ptr(), G, r, plan.data_ptr(), hr.data_ptr(), stream))
        host = plan.cpu().numpy()
        u64 = host[80:488].view(np.uint64)
        infos.append({"spl": host[:68].view(np.uint32)[:G + 1].copy(), "dst_base": u64[:G].copy(), "send": u64[16:16 + G].copy(),
                      "recv": u64[32:32 + G].copy(), "recv_total": int(u64[48]), "max_recv_total": int(u64[49]), "total": int(u64[50])})
        plans.append(plan); hist_recv.append(hr)
    # every rank derives the same plan
    for i in infos[1:]:
        np.testing.assert_array_equal(i["spl"], infos[0]["spl"])
        assert i["max_recv_total"] == infos[0]["max_recv_total"] and i["total"] == full.size
    assert sum(i["recv_total"] for i in infos) == full.size
    assert max(i["recv_total"] for i in infos) == infos[0]["max_recv_total"]
    for r in range(G):
        assert int(infos[r]["send"].sum()) == shards[r].length
        for s in range(G):
            assert infos[r]["recv"][s] == infos[s]["send"][r]
    cap = max(infos[0]["max_recv_total"], 1)
    recv = [torch.full((cap,), 0x5A5A5A5A, dtype=torch.int32, device=dev) for _ in range(G)]
    ptrs = (ctypes.c_void_p * G)(*[t.data_ptr() for t in recv])
    for r, s in enumerate(shards):
        w = ctypes.c_size_t(0)
        b._lib.check(lib.b200_mgpu_scatter_workspace_bytes(s.length, ctypes.byref(w)))
        ws = torch.empty(max(w.value, 1), dtype=torch.uint8, device=dev)
        ws.fill_(0xA5)
        b._lib.check(lib.b200_mgpu_scatter(ws.data_ptr(), w.value, s.pointer if s.length else None, et, s.length, int(rev),
                                           plans[r].data_ptr(), ptrs, G, stream))
    outs = []
    for r in range(G):
        n = infos[r]["recv_total"]
        out = b.CuArray.undef(et, (n,))
        if n:
            w = ctypes.c_size_t(0)
            b._lib.check(lib.b200_sort_workspace_bytes(et, -1, n, 1, ctypes.byref(w)))
            ws = torch.empty(max(w.value, 1), dtype=torch.uint8, device=dev)
            ws.fill_(0xA5)
            b._lib.check(lib.b200_sort_keys_from(ws.data_ptr(), w.value, recv[r].data_ptr(), out.pointer, et, n, int(rev),
                                                 hist_recv[r].data_ptr(), stream))
        outs.append(b.Array(out))
    torch.cuda.synchronize()
    return np.concatenate(outs), infos


@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("dtype,rev", [("uint32", False), ("float32", False), ("int32", True), ("float32", True)])
def test_emulated_ranks_match_the_oracle(b, G, dtype, rev):
    rng = np.random.default_rng(0xD157 + G)
    n = 3_000_017
    if dtype == "float32":
        full = rng.standard_normal(n).astype(np.float32)
        full[[5, 6, 7, 8]] = [np.nan, -0.0, 0.0, np.inf]
    else:
        ii = np.iinfo(np.dtype(dtype))
        full = rng.integers(ii.min, ii.max, size=n, dtype=np.dtype(dtype), endpoint=True)
    got, infos = emulate(b, full, G, rev)
    want = oracle.sort(full, rev=rev)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    # balance: bucket sizes are within one 16-bit bin's population of N / G (uniform keys: a few per cent)
    if dtype != "float32":
        assert max(i["recv_total"] for i in infos) <= 1.05 * n / G


def test_ragged_shards_duplicates_and_empty_ranks(b):
    rng = np.random.default_rng(77)
    full = (rng.integers(0, 50, size=2_000_003, dtype=np.uint32) << np.uint32(20)).astype(np.uint32)      # 50 distinct keys
    cuts = np.array([0, 0, 700_001, 700_001, 2_000_003])                                                    # two empty ranks
    got, infos = emulate(b, full, 4, False, cuts=cuts)
    np.testing.assert_array_equal(got, np.sort(full))
    # equal keys are never split: every distinct key lives on one rank
    edges = np.cumsum([i["recv_total"] for i in infos])[:-1]
    for e in edges:
        if 0 < e < full.size:
            assert got[e - 1] != got[e]
    srt, _ = emulate(b, np.arange(1_000_000, dtype=np.uint32) * 4000, 8, True)                             # presorted, rev
    np.testing.assert_array_equal(srt, (np.arange(1_000_000, dtype=np.uint32) * 4000)[::-1])


def test_dense_slice_takes_the_hybrid_path(b):
    """2^25 keys per emulated rank, 4 ranks: each received slice is dense inside a quarter of the key space, so the local
    sort runs the hybrid MSD path on the plan's histogram (no histogram pass of its own)."""
    rng = np.random.default_rng(5)
    full = rng.integers(0, 2**32, size=4 << 25, dtype=np.uint64).astype(np.uint32)
    got, infos = emulate(b, full, 4, False)
    assert bool((got[1:] >= got[:-1]).all()) and got.size == full.size
    np.testing.assert_array_equal(np.bincount(got >> np.uint32(24), minlength=256), np.bincount(full >> np.uint32(24), minlength=256))
    assert int(got.astype(np.uint64).sum()) == int(full.astype(np.uint64).sum())


def emulate_sortperm(b, full, G, rev, tagged=False):
    """sortperm's exchange on one GPU: nan-canonical histogram, plan, stable pairs partition into G local buffers,
    stable local pairs sort; returns the concatenated global 1-based permutation.  tagged: the 4-byte generated payload
    (rank << shift | position) and the widening kernel instead of an Int64 payload read from memory."""
    lib = b._lib.load()
    dt = b.dtypes
    et = dt.from_np(full.dtype)
    dev = torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream().cuda_stream
    cuts = np.linspace(0, full.size, G + 1).astype(int)
    shards = [b.CuArray(full[cuts[r]:cuts[r + 1]]) for r in range(G)]
    hist_all = torch.empty(G * 65536, dtype=torch.int32, device=dev)
    for r, s in enumerate(shards):
        b._lib.check(lib.b200_mgpu_hist16(s.pointer, et, s.length, int(rev), 1, hist_all[r * 65536:].data_ptr(), stream))
    nb = ctypes.c_size_t(0)
    b._lib.check(lib.b200_mgpu_plan_bytes(ctypes.byref(nb)))
    plans, totals = [], []
    for r in range(G):
        plan = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        hr = torch.empty(65536, dtype=torch.int32, device=dev)
        b._lib.check(lib.b200_mgpu_plan(hist_all.data_ptr(), G, r, plan.data_ptr(), hr.data_ptr(), stream))
        u64 = plan.cpu().numpy()[80:488].view(np.uint64)
        plans.append(plan); totals.append((int(u64[48]), int(u64[49])))
        max_shard = int(plan.cpu().numpy()[72:76].view(np.uint32)[0])
        assert max_shard == max(s.length for s in shards)
    shift = max(1, (max_shard - 1).bit_length())
    cap = max(totals[0][1], 1)
    rk = [torch.zeros(cap, dtype=torch.int32, device=dev) for _ in range(G)]
    rp = [torch.zeros(cap, dtype=torch.int32 if tagged else torch.int64, device=dev) for _ in range(G)]
    kp = (ctypes.c_void_p * G)(*[t.data_ptr() for t in rk])
    pp = (ctypes.c_void_p * G)(*[t.data_ptr() for t in rp])
    for r, s in enumerate(shards):
        idx = torch.arange(cuts[r] + 1, cuts[r + 1] + 1, dtype=torch.int64, device=dev)
        w = ctypes.c_size_t(0)
        b._lib.check(lib.b200_partition_workspace_bytes(s.length, G, ctypes.byref(w)))
        ws = torch.empty(w.value + 256, dtype=torch.uint8, device=dev)
        ws.fill_(0xA5)
        if tagged:
            b._lib.check(lib.b200_mgpu_partition_index(ws.data_ptr(), w.value + 256, s.pointer, et, s.length, int(rev),
                                                       plans[r].data_ptr(), r, shift, kp, pp, G, stream))
        else:
            b._lib.check(lib.b200_mgpu_partition_pairs(ws.data_ptr(), w.value + 256, s.pointer, idx.data_ptr(), et, s.length,
                                                       int(rev), plans[r].data_ptr(), kp, pp, G, stream))
    out = []
    ops = b.multigpu.DeviceOps()
    offsets = torch.zeros(16, dtype=torch.int64, device=dev)
    offsets[:G] = torch.as_tensor(cuts[:G].astype(np.int64), device=dev)
    for r in range(G):
        n = totals[r][0]
        keys = b.CuArray(rk[r][:n].view(torch.uint8), (n,), et)
        pay = b.CuArray(rp[r][:n].view(torch.uint8), (n,), dt.U32 if tagged else dt.I64)
        ops.sort_pairs(keys, pay, rev)
        out.append(b.Array(ops.widen_index(rp[r][:n].view(torch.uint8), n, shift, offsets)) if tagged else b.Array(pay))
    return np.concatenate(out)


@pytest.mark.parametrize("tagged", [False, True])
@pytest.mark.parametrize("G", [2, 5])
@pytest.mark.parametrize("rev", [False, True])
def test_emulated_sortperm_is_stable_across_ranks(b, G, rev, tagged):
    rng = np.random.default_rng(31 + G)
    full = rng.standard_normal(1_500_007).astype(np.float32)
    full[::97] = 0.25                                   # heavy ties: stability across tiles, ranks and the local sort
    full[[11, 500_000, 1_400_000]] = np.nan             # several NaNs with the same key: they keep ascending index
    full[[12, 13]] = [-0.0, 0.0]
    np.testing.assert_array_equal(emulate_sortperm(b, full, G, rev, tagged), oracle.sortperm(full, rev=rev))
    u = rng.integers(0, 40, size=700_001, dtype=np.uint32) * np.uint32(100_000_000)
    np.testing.assert_array_equal(emulate_sortperm(b, u, G, rev, tagged), oracle.sortperm(u, rev=rev))
