"""Host-side logic of the sharded paths (cuda.jl_b200/multigpu.py) on CPU: world_size-2 and -3
`gloo` groups with a numpy/oracle-backed stand-in for the shard-local kernels.  Checks the carry
exchange of the distributed scan, the partial combination of the distributed reduce and the
splitter / all-to-all / stability logic of the sample sort against the single-array oracle."""
import os
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _mk(arr, b):
    """CuArray container over a CPU buffer (test only: the product constructs them on the GPU)."""
    a = np.ascontiguousarray(arr)
    return b.CuArray(torch.from_numpy(a.view(np.uint8).copy()), (a.size,), b.dtypes.from_np(a.dtype))


class FakeOps:
    """numpy/oracle stand-in for the shard-local library calls."""

    def __init__(self, b, oracle):
        self.b, self.o = b, oracle

    def _name(self, op):
        b = self.b
        return {b.add: "add", b.add_sum: "add", b.mul: "mul", b.mul_prod: "mul", b.max_: "max", b.min_: "min",
                b.and_: "and", b.or_: "or"}[op]

    def reduce(self, f, op, shard, out_eltype):
        a = shard.to_host()
        odt = self.b.dtypes.np_dtype(out_eltype)
        if a.size == 0:
            val = np.asarray(self.b.neutral_element(op, out_eltype), dtype=odt).reshape(1)
        else:
            val = np.asarray(self.o.mapreducedim("identity" if f is self.b.identity else "abs2", self._name(op), a, (1,), odt)).astype(odt).reshape(1)
        return _mk(val, self.b)

    def scan(self, op, out, shard, init, exclusive, carry=None, carry_count=0):
        a = shard.to_host()
        odt = self.b.dtypes.np_dtype(out.eltype)
        if carry is not None and carry_count > 0:        # what b200_scan_carry does on the device
            c = carry.to_host()[:carry_count]
            acc = c[0]
            for x in c[1:]:
                acc = self.o.mapreducedim("identity", self._name(op), np.array([acc, x]), (1,), c.dtype)[0]
            init = acc if init is None else self.o.mapreducedim("identity", self._name(op), np.array([init, acc]), (1,), c.dtype)[0]
            init = np.asarray(init).astype(odt)[()]
        r = np.asarray(self.o.accumulate(self._name(op), a, 1, init, odt, exclusive)).astype(odt)
        out.buf.copy_(torch.from_numpy(np.ascontiguousarray(r).view(np.uint8)))
        return out

    def sort_keys(self, c, rev):
        a = c.to_host()
        c.buf.copy_(torch.from_numpy(np.ascontiguousarray(self.o.sort(a, rev=rev)).view(np.uint8).copy()))
        return c

    def sort_pairs(self, keys, payload, rev):
        k, p = keys.to_host(), payload.to_host()
        perm = self.o.sortperm(k, rev=rev) - 1
        keys.buf.copy_(torch.from_numpy(np.ascontiguousarray(k[perm]).view(np.uint8).copy()))
        payload.buf.copy_(torch.from_numpy(np.ascontiguousarray(p[perm]).view(np.uint8).copy()))

    # ---- numpy stand-ins for csrc/mgpu_sort.cu (same decisions as the kernels; the kernels themselves are tested
    #      on one GPU with G emulated ranks in tests/test_mgpu_sort_gpu.py)
    emulates_device_path = True

    @property
    def MGPU_KEY_TYPES(self):
        d = self.b.dtypes
        return (d.U32, d.I32, d.F32)

    def _bins(self, a, rev):
        k = self.o.total_order_key(a).astype(np.uint32)
        return ((~k if rev else k) >> np.uint32(16)).astype(np.int64)

    def mgpu_hist16(self, shard, rev):
        h = np.bincount(self._bins(shard.to_host(), rev), minlength=65536).astype(np.uint32)
        return _mk(h, self.b)

    def mgpu_plan(self, hist_all, G, r):
        H = hist_all.numpy().view(np.uint32).reshape(G, 65536).astype(np.int64)
        g = H.sum(axis=0)
        excl = np.concatenate([[0], np.cumsum(g)[:-1]])
        total = int(g.sum())
        spl = np.array([0] + [int((excl < (total * b) // G).sum()) for b in range(1, G)] + [65536], dtype=np.int64)
        cnt = np.array([[H[s, spl[b]:spl[b + 1]].sum() for b in range(G)] for s in range(G)], dtype=np.int64)   # [src, dst]
        hist_recv = np.where((np.arange(65536) >= spl[r]) & (np.arange(65536) < spl[r + 1]), g, 0).astype(np.uint32)
        info = {"spl": spl, "dst_base": cnt[:r].sum(axis=0), "send_count": cnt[r], "recv_count": cnt[:, r],
                "recv_total": int(cnt[:, r].sum()), "max_recv_total": int(cnt.sum(axis=0).max()), "total": total}
        return None, _mk(hist_recv, self.b), info

    def mgpu_exchange(self, shard, rev, plan, info, group):
        a = shard.to_host()
        G = dist.get_world_size(group)
        r = dist.get_rank(group)
        bucket = np.searchsorted(info["spl"][1:-1], self._bins(a, rev), side="right")
        raw = a.view(np.uint32)
        send = [torch.from_numpy(raw[bucket == d].astype(np.int64)) for d in range(G)]
        assert [t.numel() for t in send] == [int(x) for x in info["send_count"]]
        recv = [torch.empty(int(c), dtype=torch.int64) for c in info["recv_count"]]
        # gloo has no all_to_all on CPU in every build: pairwise send / recv in rank order
        for src in range(G):
            for dst in range(G):
                if src == dst == r:
                    recv[src].copy_(send[dst])
                elif src == r:
                    dist.send(send[dst], dst, group=group)
                elif dst == r:
                    dist.recv(recv[src], src, group=group)
        got = torch.cat(recv).to(torch.int64).numpy().astype(np.uint32)
        assert got.size == info["recv_total"]
        return torch.from_numpy(got.view(np.uint8).copy())

    def sort_keys_from(self, src_u8, eltype, n, rev, hist_recv):
        a = src_u8.numpy().view(self.b.dtypes.np_dtype(eltype))[:n]
        if n:
            np.testing.assert_array_equal(np.bincount(self._bins(a, rev), minlength=65536), hist_recv.to_host())
        return _mk(self.o.sort(a, rev=rev), self.b)

    def partition(self, keys, payload, splitters, nbuckets, rev):
        k = keys.to_host()
        ke = self.o.total_order_key(k)
        se = self.o.total_order_key(splitters.to_host())
        if rev:
            ke, se = ~ke, ~se
        bid = np.searchsorted(se, ke, side="right") if se.size else np.zeros(k.size, dtype=np.int64)
        order = np.argsort(bid, kind="stable")
        counts = np.bincount(bid, minlength=nbuckets).astype(np.int64)
        kout = _mk(k[order], self.b)
        pout = _mk(payload.to_host()[order], self.b) if payload is not None else None
        return kout, pout, _mk(counts, self.b)

    def lower_bound(self, sorted_keys, needles, rev):
        ks = self.o.total_order_key(sorted_keys.to_host())
        nd = self.o.total_order_key(needles.to_host())
        if rev:
            ks, nd = ~ks, ~nd
        return _mk(np.searchsorted(ks, nd, side="left").astype(np.int64), self.b)


def _worker(rank, world, port, case, ret):
    try:
        if case == "sort_padded":       # force the equal-slot exchange used from 6 ranks on
            os.environ["B200_PADDED_EXCHANGE_MIN_RANKS"] = "2"
            case = "sort"
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import b200arrays as b
        import oracle
        ops = FakeOps(b, oracle)
        mg = b.multigpu
        rng = np.random.default_rng(1234)
        out = {}
        if case == "reduce_scan":
            full_i = rng.integers(-2**40, 2**40, size=100_003, dtype=np.int64)
            full_f = rng.random(70_001).astype(np.float32)
            cuts_i = np.linspace(0, full_i.size, world + 1).astype(int)
            cuts_f = np.linspace(0, full_f.size, world + 1).astype(int)
            si = _mk(full_i[cuts_i[rank]:cuts_i[rank + 1]], b)
            sf = _mk(full_f[cuts_f[rank]:cuts_f[rank + 1]], b)
            out["sum_i"] = int(mg.dist_mapreduce(b.identity, b.add_sum, si, ops=ops))
            out["max_f"] = float(mg.dist_mapreduce(b.identity, b.max_, sf, ops=ops))
            out["sum_f"] = float(mg.dist_mapreduce(b.identity, b.add_sum, sf, ops=ops))
            mn, mx = mg.dist_extrema(si, ops=ops)
            out["ext"] = (int(mn), int(mx))
            out["scan_i"] = mg.dist_accumulate(b.add, si, ops=ops).to_host()
            out["scan_i_excl"] = mg.dist_accumulate(b.add, si, exclusive=True, ops=ops).to_host()
            out["scan_f"] = mg.dist_accumulate(b.add, sf, ops=ops).to_host()
            out["scan_max"] = mg.dist_accumulate(b.max_, si, ops=ops).to_host()
            # Bool and unsigned partials across ranks (max over Bool shards was combined with MIN)
            fb = np.zeros(999, dtype=np.bool_); fb[700] = True          # only the LAST rank's shard holds a True
            cb = np.linspace(0, fb.size, world + 1).astype(int)
            sb = _mk(fb[cb[rank]:cb[rank + 1]], b)
            for name, op in (("bmax", b.max_), ("bmin", b.min_), ("bor", b.or_), ("band", b.and_)):
                out[name] = bool(mg.dist_mapreduce(b.identity, op, sb, ops=ops))
            fu = rng.integers(0, 2**32, size=5000, dtype=np.uint32)
            fu[4000] = 0xFFFFFFF0; fu[10] = 3                           # extremes on different ranks, above 2^31
            cu = np.linspace(0, fu.size, world + 1).astype(int)
            su = _mk(fu[cu[rank]:cu[rank + 1]], b)
            out["umax"] = int(mg.dist_mapreduce(b.identity, b.max_, su, ops=ops))
            out["umin"] = int(mg.dist_mapreduce(b.identity, b.min_, su, ops=ops))
            try:
                mg.dist_mapreduce(b.identity, b.add, sb, ops=ops)
                out["bool_add"] = "no error"
            except b.NotGraftedError:
                out["bool_add"] = "NotGrafted"
        elif case == "sort":
            full = rng.standard_normal(50_000).astype(np.float32)
            full[::97] = 0.25                       # heavy duplicates
            full[5] = np.nan; full[6] = -0.0; full[7] = 0.0
            cuts = np.array([0] + sorted(rng.integers(0, full.size, world - 1).tolist()) + [full.size])   # ragged shards
            shard = _mk(full[cuts[rank]:cuts[rank + 1]], b)
            for rev in (False, True):
                srt, counts = mg.dist_sort(shard, rev=rev, ops=ops)
                out[f"sort_{rev}"] = srt.to_host()
                out[f"perm_{rev}"] = mg.dist_sortperm(shard, rev=rev, ops=ops).to_host()
            u = rng.integers(0, 50, size=20_000, dtype=np.uint32)             # few unique keys
            cu = np.linspace(0, u.size, world + 1).astype(int)
            out["perm_u"] = mg.dist_sortperm(_mk(u[cu[rank]:cu[rank + 1]], b), ops=ops).to_host()
            empty = _mk(np.zeros(0, dtype=np.float32), b)
            out["empty"] = mg.dist_sort(empty, ops=ops)[0].to_host()
        ret[rank] = out
        dist.destroy_process_group()
    except Exception:
        ret[rank] = {"error": traceback.format_exc()}


def _run(world, case):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert not p.is_alive(), "worker hung"
    res = [ret.get(r) for r in range(world)]
    for r in res:
        assert r is not None and "error" not in r, r
    return res


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_reduce_and_scan(world):
    import oracle
    res = _run(world, "reduce_scan")
    rng = np.random.default_rng(1234)
    full_i = rng.integers(-2**40, 2**40, size=100_003, dtype=np.int64)
    full_f = rng.random(70_001).astype(np.float32)
    for r in res:
        assert r["sum_i"] == int(full_i.sum())
        assert r["max_f"] == float(full_f.max())
        assert abs(r["sum_f"] - float(full_f.astype(np.float64).sum())) <= 1e-6 * float(full_f.sum()) * 20
        assert r["ext"] == (int(full_i.min()), int(full_i.max()))
        assert (r["bmax"], r["bmin"], r["bor"], r["band"]) == (True, False, True, False)
        assert r["umax"] == 0xFFFFFFF0 and r["umin"] <= 3
        assert r["bool_add"] == "NotGrafted"
    np.testing.assert_array_equal(np.concatenate([r["scan_i"] for r in res]), np.cumsum(full_i))
    excl = np.concatenate([[0], np.cumsum(full_i)[:-1]])
    np.testing.assert_array_equal(np.concatenate([r["scan_i_excl"] for r in res]), excl)
    np.testing.assert_array_equal(np.concatenate([r["scan_max"] for r in res]), np.maximum.accumulate(full_i))
    got = np.concatenate([r["scan_f"] for r in res]).astype(np.float64)
    ref = np.cumsum(full_f.astype(np.float64))
    assert np.all(np.abs(got - ref) <= oracle.scan_tolerance(full_f, 1, np.float32) + 1e-7 * ref)


@pytest.mark.parametrize("world,case", [(2, "sort"), (3, "sort"), (3, "sort_padded")])
def test_sample_sort(world, case):
    import oracle
    res = _run(world, case)
    rng = np.random.default_rng(1234)
    full = rng.standard_normal(50_000).astype(np.float32)
    full[::97] = 0.25
    full[5] = np.nan; full[6] = -0.0; full[7] = 0.0
    _ = rng.integers(0, full.size, world - 1)
    u = rng.integers(0, 50, size=20_000, dtype=np.uint32)
    for rev in (False, True):
        got = np.concatenate([r[f"sort_{rev}"] for r in res])
        np.testing.assert_array_equal(got.view(np.uint32), oracle.sort(full, rev=rev).view(np.uint32))
        perm = np.concatenate([r[f"perm_{rev}"] for r in res])
        np.testing.assert_array_equal(perm, oracle.sortperm(full, rev=rev))       # stable across ranks
    np.testing.assert_array_equal(np.concatenate([r["perm_u"] for r in res]), oracle.sortperm(u))
    assert all(r["empty"].size == 0 for r in res)


# ---- bench.py: the sharded extras can never cost the headline JSON line ----------------------------
@pytest.mark.parametrize("mode", ["ok", "raise", "hang"])
def test_bench_sharded_extras_guard(mode):
    """bench.py --gpus N runs the sharded configs AFTER the headline measurement; whatever they do (return, raise,
    hang in a collective) rank 0 must still print exactly one JSON line and exit 0."""
    import json
    import subprocess
    import sys
    code = f"""
import sys, time
sys.path.insert(0, {ROOT!r})
import bench
line = {{"metric": "m", "value": 1.0, "extras": {{}}}}
def fn():
    if {mode!r} == "raise": raise RuntimeError("boom")
    if {mode!r} == "hang": time.sleep(60)
    return {{"cumsum": {{"ms": 1.0}}}}
bench.finish_with_sharded_extras(line, 0, fn, timeout=1.0)
print("not reached")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] == 1.0
    sh = d["extras"]["sharded"]
    if mode == "ok":
        assert sh == {"cumsum": {"ms": 1.0}}
    elif mode == "raise":
        assert sh.startswith("failed: RuntimeError: boom")
    else:
        assert sh.startswith("timeout")
