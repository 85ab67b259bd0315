"""Sharded (one process per GPU) versions of the 1-D primitives — SURVEY §8(e).

The reference has NO multi-GPU path for reduce / scan / sort (SURVEY §2a "Collectives: none");
this is new functionality layered on the shard-local kernels, as BASELINE.json's north star
asks: shard-local reductions + allreduce of the partials; shard-local scans + all-gather of the
shard carries; sample sort with splitter selection + all-to-all + local radix sort.
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the plumbing; every pass over the
data is one of libb200arrays' kernels.

Each rank holds a contiguous shard of a conceptual global vector, in rank order.

`ops` is the provider of shard-local operations: `DeviceOps` (the library) by default; the CPU
`gloo` tests inject a numpy-backed stand-in to exercise the host-side logic without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import dtypes as dt
from .cuarray import CuArray, current_stream_handle
from .errors import NotGraftedError

_TRANSPORT = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}


class DeviceOps:
    """Shard-local operations on CuArrays through libb200arrays."""

    def reduce(self, f, op, shard, out_eltype):
        from .mapreduce import mapreducedim_, neutral_element, extrema_rf
        R = CuArray.undef(out_eltype, (1,), shard.buf.device)
        if shard.length == 0:
            from .mapreduce import _fill
            _fill(R, neutral_element(op, out_eltype))
            return R
        init = 0 if op is extrema_rf else neutral_element(op, out_eltype)
        mapreducedim_(f, op, R, shard, init=init)
        return R

    def extrema_pair(self, shard):
        """(min, max) of one shard in ONE pass (B200_OP_EXTREMA); the neutral pair for an empty shard."""
        from .mapreduce import mapreducedim_, identity, extrema_rf, neutral_element, min_, max_, _fill
        R = CuArray.undef(shard.eltype, (2, 1), shard.buf.device)
        if shard.length == 0:
            host = np.array([neutral_element(min_, shard.eltype), neutral_element(max_, shard.eltype)], dtype=dt.np_dtype(shard.eltype))
            R.buf.copy_(CuArray(host).buf)
            return R.reshape((2,))
        mapreducedim_(identity, extrema_rf, R, shard.vec(), init=0)
        return R.reshape((2,))

    def extrema_combine(self, g2, R):
        """g2: (2, G) gathered pairs -> host (min over row 1, max over row 2) with the library's min / max kernels."""
        from .mapreduce import mapreducedim_, identity, min_, max_, neutral_element
        G = g2.dims[1]
        mins = CuArray(g2.buf.view(-1), (2, G), g2.eltype)
        out = CuArray.undef(g2.eltype, (2, 1), g2.buf.device)
        # rows of a column-major (2, G) matrix: reduce dims=2 with min gives (min of mins, min of maxs); likewise max
        mapreducedim_(identity, min_, out, mins, init=neutral_element(min_, g2.eltype))
        lo = out.to_host().reshape(-1)[0]
        mapreducedim_(identity, max_, out, mins, init=neutral_element(max_, g2.eltype))
        hi = out.to_host().reshape(-1)[1]
        return lo, hi

    def scan(self, op, out, shard, init, exclusive, carry=None, carry_count=0):
        from .accumulate import scan_
        return scan_(op, out, shard, 1, init=init, exclusive=exclusive, carry=carry, carry_count=carry_count)

    def sort_keys(self, c, rev):
        from .sorting import sort_
        return sort_(c, rev=rev)

    # -- device-driven keys-only sample sort (csrc/mgpu_sort.cu) ---------------------------------------------
    MGPU_KEY_TYPES = (dt.U32, dt.I32, dt.F32)

    def mgpu_hist16(self, shard, rev, nan_canonical=False):
        """Histogram of the top 16 bits of the order-mapped keys: CuArray{UInt32}(65536).  nan_canonical: sortperm's
        key map (every NaN is one key) instead of sort!'s (NaN payloads kept)."""
        from . import _lib
        lib = _lib.load()
        hist = CuArray.undef(dt.U32, (65536,), shard.buf.device)
        _lib.check(lib.b200_mgpu_hist16(shard.pointer if shard.length else None, shard.eltype, shard.length,
                                        1 if rev else 0, 1 if nan_canonical else 0, hist.pointer,
                                        current_stream_handle(shard.buf.device)))
        return hist

    def mgpu_exchange_pairs(self, shard, payload, rev, plan, info, group):
        """sortperm's exchange: stable partition of (keys, Int64 payload) by the plan's splitters, every bucket stored
        straight into the destination rank's symmetric receive buffers.  Returns uint8 views (keys, payload)."""
        import ctypes
        from . import _lib
        from .cuarray import Workspace
        lib = _lib.load()
        G, r = _world(group)
        dev = shard.buf.device
        cap = max(info["max_recv_total"], 1)
        kbuf, khdl = _symm_buffer("keys", cap * 4, dev, group)
        pbuf, phdl = _symm_buffer("payload", cap * 8, dev, group)
        nb = ctypes.c_size_t(0)
        _lib.check(lib.b200_partition_workspace_bytes(shard.length, G, ctypes.byref(nb)))
        ws = Workspace(nb.value + 256, dev)
        kp = (ctypes.c_void_p * G)(*[int(x) for x in khdl.buffer_ptrs])
        pp = (ctypes.c_void_p * G)(*[int(x) for x in phdl.buffer_ptrs])
        khdl.barrier()
        _lib.check(lib.b200_mgpu_partition_pairs(ws.ptr, ws.nbytes, shard.pointer if shard.length else None,
                                                 payload.pointer if shard.length else None, shard.eltype, shard.length,
                                                 1 if rev else 0, plan.data_ptr(), kp, pp, G, current_stream_handle(dev)))
        khdl.barrier()
        n = info["recv_total"]
        return kbuf[:n * 4], pbuf[:n * 8]

    def mgpu_exchange_index(self, shard, rev, plan, info, group, shift):
        """sortperm's exchange with a generated payload: stable partition of the keys, each carrying the 4-byte tagged
        index (rank << shift) | position in this shard.  Returns uint8 views (keys, tagged UInt32 indices)."""
        import ctypes
        from . import _lib
        from .cuarray import Workspace
        lib = _lib.load()
        G, r = _world(group)
        dev = shard.buf.device
        cap = max(info["max_recv_total"], 1)
        kbuf, khdl = _symm_buffer("keys", cap * 4, dev, group)
        pbuf, phdl = _symm_buffer("index", cap * 4, dev, group)
        nb = ctypes.c_size_t(0)
        _lib.check(lib.b200_partition_workspace_bytes(shard.length, G, ctypes.byref(nb)))
        ws = Workspace(nb.value + 256, dev)
        kp = (ctypes.c_void_p * G)(*[int(x) for x in khdl.buffer_ptrs])
        pp = (ctypes.c_void_p * G)(*[int(x) for x in phdl.buffer_ptrs])
        khdl.barrier()
        _lib.check(lib.b200_mgpu_partition_index(ws.ptr, ws.nbytes, shard.pointer if shard.length else None, shard.eltype,
                                                 shard.length, 1 if rev else 0, plan.data_ptr(), r, shift, kp, pp, G,
                                                 current_stream_handle(dev)))
        khdl.barrier()
        n = info["recv_total"]
        return kbuf[:n * 4], pbuf[:n * 4]

    def widen_index(self, index_u8, n, shift, offsets):
        """tagged UInt32 indices -> global 1-based Int64 indices (offsets: Int64[16] on the device)."""
        from . import _lib
        lib = _lib.load()
        out = CuArray.undef(dt.I64, (n,), index_u8.device)
        _lib.check(lib.b200_mgpu_widen_index(index_u8.data_ptr() if n else None, n, shift, offsets.data_ptr(),
                                             out.pointer if n else None, current_stream_handle(index_u8.device)))
        return out

    def mgpu_plan(self, hist_all, G, r):
        """(plan buffer, hist_recv) on the device from the all-gathered histograms; then ONE small device-to-host
        copy of the plan (the only synchronisation of the sort: the slice length sizes the result)."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        dev = hist_all.device
        nb = ctypes.c_size_t(0)
        _lib.check(lib.b200_mgpu_plan_bytes(ctypes.byref(nb)))
        plan = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        hist_recv = CuArray.undef(dt.U32, (65536,), dev)
        _lib.check(lib.b200_mgpu_plan(hist_all.data_ptr(), G, r, plan.data_ptr(), hist_recv.pointer,
                                      current_stream_handle(dev)))
        host = plan.cpu().numpy()
        u64 = host[80:488].view(np.uint64)                  # dst_base[16], send_count[16], recv_count[16], 3 scalars
        info = {"spl": host[:68].view(np.uint32)[:G + 1].copy(), "dst_base": u64[:G].copy(), "send_count": u64[16:16 + G].copy(),
                "recv_count": u64[32:32 + G].copy(), "recv_total": int(u64[48]), "max_recv_total": int(u64[49]),
                "total": int(u64[50]), "max_shard_len": int(host[72:76].view(np.uint32)[0])}
        return plan, hist_recv, info

    def mgpu_exchange(self, shard, rev, plan, info, group):
        """Fused partition + all-to-all: one pass over the shard, every bucket stored straight into its destination
        rank's symmetric receive buffer (NVLink peer stores).  Returns a uint8 view of this rank's received keys."""
        import ctypes
        from . import _lib
        from .cuarray import Workspace
        lib = _lib.load()
        G, r = _world(group)
        dev = shard.buf.device
        kbuf, khdl = _symm_buffer("keys", max(info["max_recv_total"], 1) * 4, dev, group)
        nb = ctypes.c_size_t(0)
        _lib.check(lib.b200_mgpu_scatter_workspace_bytes(shard.length, ctypes.byref(nb)))
        ws = Workspace(nb.value, dev)
        ptrs = (ctypes.c_void_p * G)(*[int(x) for x in khdl.buffer_ptrs])
        khdl.barrier()                                      # nobody still reads its receive buffer from a previous call
        _lib.check(lib.b200_mgpu_scatter(ws.ptr, ws.nbytes, shard.pointer if shard.length else None, shard.eltype,
                                         shard.length, 1 if rev else 0, plan.data_ptr(), ptrs, G,
                                         current_stream_handle(dev)))
        khdl.barrier()                                      # every rank's peer stores have landed
        return kbuf[:info["recv_total"] * 4]

    def sort_keys_from(self, src_u8, eltype, n, rev, hist_recv):
        """Local sort of the received keys: reads the receive buffer, writes a fresh result (no copy in between)."""
        import ctypes
        from . import _lib
        from .cuarray import Workspace
        lib = _lib.load()
        dev = src_u8.device
        out = CuArray.undef(eltype, (n,), dev)
        if n == 0:
            return out
        nb = ctypes.c_size_t(0)
        _lib.check(lib.b200_sort_workspace_bytes(eltype, -1, n, 1, ctypes.byref(nb)))
        ws = Workspace(nb.value, dev)
        _lib.check(lib.b200_sort_keys_from(ws.ptr, ws.nbytes, src_u8.data_ptr(), out.pointer, eltype, n, 1 if rev else 0,
                                           hist_recv.pointer if hist_recv is not None else None,
                                           current_stream_handle(dev)))
        return out

    def sort_pairs(self, keys, payload, rev):
        import ctypes
        from . import _lib
        from .cuarray import Workspace
        lib = _lib.load()
        if keys.length <= 1:
            return
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.b200_sort_workspace_bytes(keys.eltype, payload.eltype, keys.length, 1, ctypes.byref(nbytes)))
        ws = Workspace(nbytes.value, keys.buf.device)
        _lib.check(lib.b200_sort_pairs(ws.ptr, ws.nbytes, keys.pointer, payload.pointer, keys.eltype, payload.eltype,
                                       keys.length, 1 if rev else 0, current_stream_handle(keys.buf.device)))

    def partition(self, keys, payload, splitters, nbuckets, rev):
        """Stable partition of an unsorted shard into nbuckets (<= 16) destination buckets.
        Returns (keys_out, payload_out | None, counts CuArray{Int64}[nbuckets])."""
        import ctypes
        from . import _lib
        from .cuarray import Workspace
        lib = _lib.load()
        dev = keys.buf.device
        kout = keys.similar()
        pout = payload.similar() if payload is not None else None
        counts = CuArray.undef(dt.I64, (nbuckets,), dev)
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.b200_partition_workspace_bytes(keys.length, nbuckets, ctypes.byref(nbytes)))
        ws = Workspace(nbytes.value, dev)
        _lib.check(lib.b200_partition(ws.ptr, ws.nbytes, keys.pointer if keys.length else None,
                                      kout.pointer if keys.length else None,
                                      payload.pointer if (payload is not None and keys.length) else None,
                                      pout.pointer if (pout is not None and keys.length) else None,
                                      keys.eltype, keys.length, splitters.pointer if splitters.length else None,
                                      nbuckets, 1 if rev else 0, counts.pointer, current_stream_handle(dev)))
        return kout, pout, counts

    def partition_count(self, keys, splitters, nbuckets, rev):
        """Phase 1 of the fused partition + exchange: bucket totals (device Int64[nbuckets]); the returned
        workspace keeps the per-tile prefixes for partition_scatter."""
        import ctypes
        from . import _lib
        from .cuarray import Workspace
        lib = _lib.load()
        dev = keys.buf.device
        counts = CuArray.undef(dt.I64, (nbuckets,), dev)
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.b200_partition_workspace_bytes(keys.length, nbuckets, ctypes.byref(nbytes)))
        ws = Workspace(nbytes.value, dev)
        _lib.check(lib.b200_partition_count(ws.ptr, ws.nbytes, keys.pointer if keys.length else None, keys.eltype,
                                            keys.length, splitters.pointer if splitters.length else None, nbuckets,
                                            1 if rev else 0, counts.pointer, current_stream_handle(dev)))
        return counts, ws

    def partition_scatter(self, ws, keys, payload, splitters, nbuckets, rev, counts, dst_key_ptrs, dst_payload_ptrs,
                          dst_base):
        """Phase 2: bucket b is written straight to dst_key_ptrs[b][dst_base[b]:] — peer memory over NVLink."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        if keys.length == 0:
            return
        PtrArr = ctypes.c_void_p * nbuckets
        I64Arr = ctypes.c_int64 * nbuckets
        dk = PtrArr(*[int(x) for x in dst_key_ptrs])
        dv = PtrArr(*[int(x) for x in dst_payload_ptrs]) if payload is not None else None
        db = I64Arr(*[int(x) for x in dst_base])
        _lib.check(lib.b200_partition_scatter(ws.ptr, ws.nbytes, keys.pointer, payload.pointer if payload is not None else None,
                                              keys.eltype, keys.length, splitters.pointer if splitters.length else None,
                                              nbuckets, 1 if rev else 0, counts.pointer, dk, dv, db,
                                              current_stream_handle(keys.buf.device)))

    def lower_bound(self, sorted_keys, needles, rev):
        from . import _lib
        lib = _lib.load()
        out = CuArray.undef(dt.I64, (needles.length,), sorted_keys.buf.device)
        _lib.check(lib.b200_search_sorted(sorted_keys.pointer if sorted_keys.length else None, sorted_keys.length,
                                          needles.pointer if needles.length else None, needles.length,
                                          sorted_keys.eltype, 1 if rev else 0, out.pointer,
                                          current_stream_handle(sorted_keys.buf.device)))
        return out


def _world(group):
    return dist.get_world_size(group), dist.get_rank(group)


def _typed(c):
    """Storage viewed with a dtype every backend can transport (bit-preserving)."""
    if c.buf.numel() == 0:
        return torch.empty(0, dtype=_TRANSPORT[dt.SIZES[c.eltype]], device=c.buf.device)
    return c.buf.view(_TRANSPORT[dt.SIZES[c.eltype]])


def _acc_eltype(eltype):
    return dt.F32 if eltype in (dt.F16, dt.BF16) else eltype


# ----------------------------------------------------------------------------- reduce
def dist_mapreduce(f, op, shard, group=None, ops=None):
    """mapreduce(f, op, A) over the concatenation of all shards: shard-local single-pass
    reduction + allreduce of one partial per GPU.  Float16/BFloat16 partials are combined in
    Float32 and rounded once at the end.  Returns a host scalar (like `dims=:` does)."""
    from .mapreduce import (add, add_sum, mul, mul_prod, max_, min_, and_, or_, _out_eltype)
    ops = ops or DeviceOps()
    out_et = _out_eltype(f, op, shard.eltype)
    part_et = _acc_eltype(out_et)
    partial = ops.reduce(f, op, shard, part_et)
    red = {add: dist.ReduceOp.SUM, add_sum: dist.ReduceOp.SUM, mul: dist.ReduceOp.PRODUCT,
           mul_prod: dist.ReduceOp.PRODUCT, max_: dist.ReduceOp.MAX, min_: dist.ReduceOp.MIN,
           and_: dist.ReduceOp.BAND, or_: dist.ReduceOp.BOR}.get(op)
    if red is None:
        raise NotGraftedError(f"no collective for {op}")
    t = partial.typed()
    if t.dtype in (torch.uint32, torch.uint64):
        # NCCL/gloo have no unsigned reductions.  Wrap-around add / mul and the bitwise ops are bit-identical on the
        # signed view; max / min compare correctly once the sign bit is flipped (order-preserving map to signed).
        st = t.view(torch.int32 if t.dtype == torch.uint32 else torch.int64)
        if red in (dist.ReduceOp.MAX, dist.ReduceOp.MIN):
            sign = torch.iinfo(st.dtype).min
            st.bitwise_xor_(sign)
            dist.all_reduce(st, op=red, group=group)
            st.bitwise_xor_(sign)
        else:
            dist.all_reduce(st, op=red, group=group)
    elif t.dtype == torch.bool:
        # Bool partials travel as bytes: `|` and max are a byte MAX, `&`, min and `*` a byte MIN; `+` on Bool would
        # need a widened output (Base: reduce(+, bools) isa Int) and is not a sharded reduction we graft.
        if red in (dist.ReduceOp.BOR, dist.ReduceOp.MAX):
            bop = dist.ReduceOp.MAX
        elif red in (dist.ReduceOp.BAND, dist.ReduceOp.MIN, dist.ReduceOp.PRODUCT):
            bop = dist.ReduceOp.MIN
        else:
            raise NotGraftedError("reduce(+, ::Bool shards): use count / add_sum (Bool -> Int64)")
        dist.all_reduce(t.view(torch.uint8), op=bop, group=group)
    else:
        dist.all_reduce(t, op=red, group=group)
    host = partial.to_host().reshape(-1)[0]
    return host.astype(dt.np_dtype(out_et)) if part_et != out_et else host


def dist_extrema(shard, group=None, ops=None):
    """extrema over the concatenation of all shards: ONE fused pass per shard (B200_OP_EXTREMA writes (min, max)),
    one all-gather of the G pairs, and the library's own min / max over the gathered values — Julia's rules
    (NaN-propagating, -0.0 < +0.0) hold across ranks because the same kernels combine the partials."""
    from .mapreduce import identity, max_, min_, extrema_rf, mapreducedim_
    ops = ops or DeviceOps()
    G, r = _world(group)
    if not hasattr(ops, "extrema_pair"):
        return (dist_mapreduce(identity, min_, shard, group, ops), dist_mapreduce(identity, max_, shard, group, ops))
    pair = ops.extrema_pair(shard)                                   # CuArray{T}(2): (min, max), neutral for an empty shard
    gathered = CuArray.undef(shard.eltype, (2 * G,), shard.buf.device)
    if shard.buf.is_cuda and hasattr(dist, "all_gather_into_tensor"):
        dist.all_gather_into_tensor(gathered.buf, pair.buf, group=group)       # bytes: NCCL has no 16-bit integer type
    else:
        _all_gather_fallback(gathered.buf, pair.buf, group)
    g2 = gathered.reshape((2, G))                                    # column-major: column k = rank k's (min, max)
    R = CuArray.undef(shard.eltype, (2, 1), shard.buf.device)
    h = ops.extrema_combine(g2, R)
    return h[0], h[1]


# ----------------------------------------------------------------------------- scan
def dist_accumulate(op, shard, exclusive=False, group=None, ops=None, out_eltype=None):
    """accumulate(op, A) over the concatenation of all shards.  Each rank reduces its shard
    (N reads), the G shard totals are all-gathered, and every rank scans its shard once with the
    totals of the ranks before it folded into the seed on the device (N reads + N writes):
    3N traffic instead of the 4N of scan-then-fix-up, and no host synchronisation anywhere."""
    from .mapreduce import identity, neutral_element
    ops = ops or DeviceOps()
    G, r = _world(group)
    out_et = out_eltype if out_eltype is not None else shard.eltype
    part_et = _acc_eltype(out_et)
    total = ops.reduce(identity, op, shard, part_et)
    gathered = CuArray.undef(part_et, (G,), shard.buf.device)
    # bytes on the wire: NCCL has no 16-bit integer type, and a gather only moves bits
    dist.all_gather_into_tensor(gathered.buf, total.buf, group=group) if hasattr(dist, "all_gather_into_tensor") \
        and shard.buf.is_cuda else _all_gather_fallback(gathered.buf, total.buf, group)
    out = shard.similar(out_et)
    # the carry of rank r = op over the totals of ranks 0 .. r-1: folded into the scan's seed ON THE DEVICE
    # (b200_scan_carry reads gathered[0:r]) — no host read-back, no extra reduce launch, graph-capturable
    ops.scan(op, out, shard, None, exclusive, carry=gathered, carry_count=r)
    return out


def _all_gather_fallback(out_t, in_t, group):
    G = dist.get_world_size(group)
    parts = [torch.empty_like(in_t) for _ in range(G)]
    dist.all_gather(parts, in_t, group=group)
    out_t.copy_(torch.cat(parts))


# ----------------------------------------------------------------------------- sort
OVERSAMPLE = 32     # regular samples per rank and per destination
import os as _os
PADDED_EXCHANGE_MIN_RANKS = int(_os.environ.get("B200_PADDED_EXCHANGE_MIN_RANKS", "6"))   # from this many ranks on, equal-size slots beat the uneven all-to-all


def _exchange(c_sorted, bounds_host, group, extra=None):
    """All-to-all of the G segments [bounds[k], bounds[k+1]) of a bucketed shard.

    The segments differ in size, but torch/NCCL's UNEVEN all_to_all_single runs at a quarter of
    the speed of the even one on 8 B200s (measured, 1 GiB per rank: 5.8 ms vs 1.6 ms), so every
    segment travels in a slot of the global maximum size (samples keep the imbalance at a few
    percent) and the received slots are compacted afterwards (device-to-device copies)."""
    G, r = _world(group)
    dev = c_sorted.buf.device
    send_counts = np.diff(bounds_host).astype(np.int64)
    sc = torch.from_numpy(send_counts).to(dev)
    allc = torch.empty(G * G, dtype=torch.int64, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(allc, sc, group=group)
    else:
        _all_gather_fallback(allc, sc, group)
    counts = allc.view(G, G).cpu().numpy()               # counts[src, dst]
    recv_counts = counts[:, r]
    n_recv = int(recv_counts.sum())
    slot = int(counts.max())
    outs = []
    padded = G >= PADDED_EXCHANGE_MIN_RANKS
    for c in (c_sorted,) + ((extra,) if extra is not None else ()):
        t_in = _typed(c)
        if not padded:
            # few ranks: the uneven collective is fine (measured at 2 GPUs: 1.6 ms vs 4.2 ms padded)
            t_out = torch.empty(n_recv, dtype=t_in.dtype, device=dev)
            dist.all_to_all_single(t_out, t_in, output_split_sizes=[int(x) for x in recv_counts],
                                   input_split_sizes=[int(x) for x in send_counts], group=group)
            outs.append(CuArray(t_out.view(torch.uint8) if n_recv else torch.empty(0, dtype=torch.uint8, device=dev),
                                (n_recv,), c.eltype))
            continue
        if slot == 0:
            outs.append(CuArray(torch.empty(0, dtype=torch.uint8, device=dev), (0,), c.eltype))
            continue
        send = torch.empty(G * slot, dtype=t_in.dtype, device=dev)
        for k in range(G):
            if send_counts[k]:
                send[k * slot:k * slot + int(send_counts[k])].copy_(t_in[int(bounds_host[k]):int(bounds_host[k + 1])])
        recv = torch.empty(G * slot, dtype=t_in.dtype, device=dev)
        dist.all_to_all_single(recv, send, group=group)
        t_out = torch.cat([recv[k * slot:k * slot + int(recv_counts[k])] for k in range(G)]) if n_recv else \
            torch.empty(0, dtype=t_in.dtype, device=dev)
        outs.append(CuArray(t_out.view(torch.uint8) if n_recv else torch.empty(0, dtype=torch.uint8, device=dev),
                            (n_recv,), c.eltype))
    return outs, counts


def _splitters(c_sorted, rev, group, ops):
    """G-1 splitter keys chosen identically on every rank from regular samples of the
    locally sorted shards (samples are all-gathered and sorted with the library's sort)."""
    G, r = _world(group)
    n = c_sorted.length
    S = OVERSAMPLE * G
    dev = c_sorted.buf.device
    tk = _typed(c_sorted)
    if n > 0:
        pos = torch.clamp(((torch.arange(S, dtype=torch.int64) + 1) * n) // (S + 1), max=n - 1).to(dev)
        samples = tk[pos]
        have = torch.ones(1, dtype=torch.int64, device=dev) * S
    else:
        samples = torch.zeros(S, dtype=tk.dtype, device=dev)
        have = torch.zeros(1, dtype=torch.int64, device=dev)
    all_s = [torch.empty_like(samples) for _ in range(G)]
    all_h = [torch.empty_like(have) for _ in range(G)]
    dist.all_gather(all_s, samples, group=group)
    dist.all_gather(all_h, have, group=group)
    pool = torch.cat([s for s, h in zip(all_s, all_h) if int(h.item()) > 0]) if any(int(h.item()) for h in all_h) \
        else torch.zeros(0, dtype=tk.dtype, device=dev)
    if pool.numel() == 0:
        return None
    cpool = CuArray(pool.contiguous().view(torch.uint8), (pool.numel(),), c_sorted.eltype)
    ops.sort_keys(cpool, rev)
    tp = _typed(cpool)
    idx = torch.tensor([(k * pool.numel()) // G for k in range(1, G)], dtype=torch.int64, device=dev)
    spl = tp[idx].contiguous()
    return CuArray(spl.view(torch.uint8), (G - 1,), c_sorted.eltype)


MAX_PARTITION_BUCKETS = 16
FUSED_EXCHANGE = _os.environ.get("B200_FUSED_EXCHANGE", "1") != "0"
_SYMM = {}     # (group id, tag) -> (uint8 symmetric tensor, handle); grown on demand, reused across calls


def _symm_buffer(tag, nbytes, dev, group):
    """A symmetric (P2P-mapped on every rank) receive buffer of at least nbytes; collective."""
    import torch.distributed._symmetric_memory as symm_mem
    g = group if group is not None else dist.group.WORLD
    key = (id(g), tag)
    cur = _SYMM.get(key)
    if cur is None or cur[0].numel() < nbytes:
        size = int(nbytes * 1.25) + (1 << 20)
        t = symm_mem.empty(size, dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(t, g)
        _SYMM[key] = (t, hdl)
    return _SYMM[key]


def _fused_partition_exchange(shard, payload, spl, rev, group, ops):
    """Partition + all-to-all in ONE kernel: the ranked scatter of b200_partition_scatter stores every
    bucket directly into the destination rank's receive buffer (symmetric memory, NVLink peer stores,
    full 128-byte lines), instead of writing buckets locally and handing them to NCCL.
    Returns (keys CuArray, payload CuArray | None) or None when the path is unavailable."""
    G, r = _world(group)
    dev = shard.buf.device
    ksz = dt.SIZES[shard.eltype]
    counts, ws = ops.partition_count(shard, spl, G, rev)
    allc = torch.empty(G * G, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, counts.typed(), group=group)
    cm = allc.view(G, G).cpu().numpy()                      # cm[src, dst]
    n_recv_all = cm.sum(axis=0)
    cap = int(n_recv_all.max())                             # same on every rank: the buffers grow collectively
    kbuf, khdl = _symm_buffer("keys", cap * ksz, dev, group)
    pbuf = phdl = None
    if payload is not None:
        pbuf, phdl = _symm_buffer("payload", cap * 8, dev, group)
    dst_base = cm[:r, :].sum(axis=0) if r > 0 else np.zeros(G, dtype=np.int64)
    khdl.barrier()                                          # nobody still reads its receive buffer from a previous call
    ops.partition_scatter(ws, shard, payload, spl, G, rev, counts, list(khdl.buffer_ptrs),
                          list(phdl.buffer_ptrs) if phdl is not None else None, dst_base)
    khdl.barrier()                                          # every rank's peer stores have landed
    n_recv = int(n_recv_all[r])
    keys = CuArray(kbuf[:n_recv * ksz].clone(), (n_recv,), shard.eltype)
    pay = CuArray(pbuf[:n_recv * 8].clone(), (n_recv,), dt.I64) if payload is not None else None
    return keys, pay, cm


def _fused_available(shard, ops):
    if not (FUSED_EXCHANGE and shard.buf.is_cuda and hasattr(ops, "partition_count")):
        return False
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        return True
    except Exception:
        return False


def dist_sort(shard, rev=False, group=None, ops=None):
    """Distributed sample sort of the concatenation of all shards.  Returns this rank's slice of
    the globally sorted sequence (slices are in rank order; their sizes are data dependent) and
    the G x G matrix of exchanged counts.

    G <= 16: partition-first — sample the UNSORTED shard, agree on splitters, cut the shard into G
    buckets with one stable partition pass (b200_partition), all-to-all, sort ONCE.  More ranks:
    sort, cut at the splitters, exchange, sort again."""
    ops = ops or DeviceOps()
    G, r = _world(group)
    if G == 1:
        local = shard.copy()
        ops.sort_keys(local, rev)
        return local, np.array([[local.length]])
    if (G <= MAX_PARTITION_BUCKETS and hasattr(ops, "mgpu_hist16") and shard.eltype in ops.MGPU_KEY_TYPES
            and DEVICE_DRIVEN_SORT and _device_path_available(shard, ops)):
        return _dist_sort_device(shard, rev, group, ops)
    if G > MAX_PARTITION_BUCKETS or not hasattr(ops, "partition"):
        return _dist_sort_sorted_first(shard, rev, group, ops)
    spl = _splitters(shard, rev, group, ops)            # regular samples of the unsorted shard
    if spl is None:
        return shard.copy(), np.zeros((G, G), dtype=np.int64)
    if _fused_available(shard, ops):
        recv, _, cmat = _fused_partition_exchange(shard, None, spl, rev, group, ops)
    else:
        parted, _, counts = ops.partition(shard, None, spl, G, rev)
        bounds = np.concatenate([[0], np.cumsum(counts.to_host().astype(np.int64))])
        (recv,), cmat = _exchange(parted, bounds, group)
    ops.sort_keys(recv, rev)
    return recv, cmat


DEVICE_DRIVEN_SORT = _os.environ.get("B200_DEVICE_DRIVEN_SORT", "1") != "0"


def _device_path_available(shard, ops):
    if not shard.buf.is_cuda:
        return getattr(ops, "emulates_device_path", False)      # CPU tests: numpy stand-in for the kernels
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        return True
    except Exception:
        return False


def _dist_sort_device(shard, rev, group, ops):
    """Keys-only sample sort whose splitters, count matrix and destination offsets are computed ON THE DEVICE from one
    all-gather of the ranks' 16-bit key histograms (csrc/mgpu_sort.cu):
        hist16 (1 read of the shard) -> all-gather 256 KB per rank -> plan kernel -> [one 512-byte read-back: the slice
        length] -> fused partition + exchange (1 read of the shard, peer stores over NVLink) -> local hybrid MSD sort
        reading the receive buffer and reusing the plan's histogram.
    Buckets are ranges of 16-bit bins, so their sizes are within one bin's population of N / G and equal keys always
    share a rank; max_recv_total sizes the receive buffers before anything is sent."""
    G, r = _world(group)
    dev = shard.buf.device
    if shard.length > (1 << 30):
        raise NotGraftedError("more than 2^30 keys per rank")
    hist = ops.mgpu_hist16(shard, rev)
    ht = hist.typed().view(torch.int32) if hist.buf.numel() else torch.zeros(65536, dtype=torch.int32, device=dev)
    hist_all = torch.empty(G * 65536, dtype=torch.int32, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(hist_all, ht, group=group)
    else:
        _all_gather_fallback(hist_all, ht, group)
    plan, hist_recv, info = ops.mgpu_plan(hist_all, G, r)
    recv = ops.mgpu_exchange(shard, rev, plan, info, group)
    out = ops.sort_keys_from(recv, shard.eltype, info["recv_total"], rev, hist_recv)
    return out, info


def _dist_sortperm_device(shard, rev, group, ops):
    """Stable distributed sortperm with device-side splitters: the same histogram / plan as _dist_sort_device (with
    sortperm's key map: all NaNs are one key), a STABLE ranked partition of (key, global index) pairs written straight
    into the destination ranks' symmetric buffers, and a stable local pairs sort.  Equal keys never leave their bin's
    rank and keep ascending global index: ranks fill every destination in rank order, tiles in tile order, lanes in
    element order.  The global index offset of this rank is summed on the device from the gathered histograms."""
    G, r = _world(group)
    dev = shard.buf.device
    hist = ops.mgpu_hist16(shard, rev, nan_canonical=True)
    hist_all = torch.empty(G * 65536, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(hist_all, hist.typed().view(torch.int32), group=group)
    plan, _, info = ops.mgpu_plan(hist_all, G, r)
    tot = hist_all.view(G, 65536).sum(dim=1, dtype=torch.int64)         # shard lengths, on the device
    n = info["recv_total"]
    # the payload that travels and is sorted is 4 bytes: (source rank, position in its shard); it fits whenever
    # bits(G - 1) + bits(longest shard - 1) <= 32, i.e. up to 2^32 keys in total with even shards
    shift = max(1, (max(int(info.get("max_shard_len", 1 << 33)), 1) - 1).bit_length())
    if shift + max(G - 1, 1).bit_length() <= 32 and hasattr(ops, "mgpu_exchange_index"):
        offsets = torch.zeros(16, dtype=torch.int64, device=dev)
        offsets[1:G] = torch.cumsum(tot, 0)[:G - 1]
        kview, iview = ops.mgpu_exchange_index(shard, rev, plan, info, group, shift)
        rk = CuArray(kview, (n,), shard.eltype)
        ri = CuArray(iview, (n,), dt.U32)
        ops.sort_pairs(rk, ri, rev)                     # in place inside the symmetric buffers, 8 bytes per pair
        return ops.widen_index(iview, n, shift, offsets)
    offset = tot[:r].sum() if r > 0 else torch.zeros((), dtype=torch.int64, device=dev)
    idx = torch.arange(1, shard.length + 1, dtype=torch.int64, device=dev) + offset
    payload = CuArray(idx.view(torch.uint8), (shard.length,), dt.I64)
    kview, pview = ops.mgpu_exchange_pairs(shard, payload, rev, plan, info, group)
    rk = CuArray(kview, (n,), shard.eltype)
    rp = CuArray(pview, (n,), dt.I64)
    ops.sort_pairs(rk, rp, rev)                     # in place inside the symmetric buffers
    return CuArray(pview.clone(), (n,), dt.I64)     # the buffers are reused by the next call: hand out a copy (8 n bytes)


def _dist_sort_sorted_first(shard, rev, group, ops):
    G, r = _world(group)
    local = shard.copy()
    ops.sort_keys(local, rev)
    spl = _splitters(local, rev, group, ops)
    if spl is None:
        return local, np.zeros((G, G), dtype=np.int64)
    lb = ops.lower_bound(local, spl, rev).to_host().astype(np.int64)
    bounds = np.concatenate([[0], lb, [local.length]])
    (recv,), counts = _exchange(local, bounds, group)
    ops.sort_keys(recv, rev)            # G sorted runs -> one (stable radix sort; a G-way merge is "next")
    return recv, counts


def dist_sortperm(shard, rev=False, group=None, ops=None):
    """Distributed stable sortperm: returns this rank's slice of the global 1-based Int64
    permutation.  The payload is the global index; equal keys keep ascending global index
    because every stage (stable partition, rank-ordered exchange, stable local sort) is stable."""
    ops = ops or DeviceOps()
    G, r = _world(group)
    dev = shard.buf.device
    if (1 < G <= MAX_PARTITION_BUCKETS and hasattr(ops, "mgpu_exchange_pairs") and shard.eltype in ops.MGPU_KEY_TYPES
            and DEVICE_DRIVEN_SORT and shard.buf.is_cuda and _device_path_available(shard, ops) and shard.length <= (1 << 30)):
        return _dist_sortperm_device(shard, rev, group, ops)
    n_local = torch.tensor([shard.length], dtype=torch.int64, device=dev)
    sizes = [torch.empty_like(n_local) for _ in range(G)]
    dist.all_gather(sizes, n_local, group=group)
    offset = int(sum(int(s.item()) for s in sizes[:r]))
    idx = torch.arange(offset + 1, offset + 1 + shard.length, dtype=torch.int64, device=dev)
    payload = CuArray(idx.view(torch.uint8), (shard.length,), dt.I64)
    if G == 1:
        keys = shard.copy()
        ops.sort_pairs(keys, payload, rev)
        return payload
    if G <= MAX_PARTITION_BUCKETS and hasattr(ops, "partition"):
        spl = _splitters(shard, rev, group, ops)
        if spl is None:
            return payload
        if _fused_available(shard, ops):
            rk, rp, _ = _fused_partition_exchange(shard, payload, spl, rev, group, ops)
        else:
            pk, pp, counts = ops.partition(shard, payload, spl, G, rev)
            bounds = np.concatenate([[0], np.cumsum(counts.to_host().astype(np.int64))])
            (rk, rp), _ = _exchange(pk, bounds, group, extra=pp)
        ops.sort_pairs(rk, rp, rev)
        return rp
    keys = shard.copy()
    ops.sort_pairs(keys, payload, rev)
    spl = _splitters(keys, rev, group, ops)
    if spl is None:
        return payload
    lb = ops.lower_bound(keys, spl, rev).to_host().astype(np.int64)
    bounds = np.concatenate([[0], lb, [keys.length]])
    (rk, rp), _ = _exchange(keys, bounds, group, extra=payload)
    ops.sort_pairs(rk, rp, rev)
    return rp
