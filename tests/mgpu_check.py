"""Multi-GPU parity check of cuda.jl_b200/multigpu.py on real GPUs (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/mgpu_check.py [log2n]

Every rank holds one shard; results are compared on the host with the CPU oracle
(bit-exact for ints / sort / sortperm, tolerance for float sums and scans)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import b200arrays as b  # noqa: E402
import oracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    timing = "--no-timing" not in sys.argv
    lg = int(args[0]) if args else 22
    n = 1 << lg
    rng = np.random.default_rng(99)
    full_f = rng.standard_normal(n).astype(np.float32)
    full_f[::1001] = 0.5
    full_f[17] = np.nan
    full_i = rng.integers(-2**40, 2**40, size=n, dtype=np.int64)
    full_u = rng.integers(0, 2**32, size=n, dtype=np.uint32)
    cuts = np.linspace(0, n, world + 1).astype(np.int64)
    lo, hi = cuts[rank], cuts[rank + 1]
    mg = b.multigpu

    def gather(c):
        """Concatenate the shards' host copies on every rank (sizes differ)."""
        h = c.to_host()
        objs = [None] * world
        dist.all_gather_object(objs, h)
        return np.concatenate(objs)

    ok = True

    def check(name, cond):
        nonlocal ok
        ok = ok and bool(cond)
        if rank == 0:
            print(f"{name:40s} {'ok' if cond else 'FAIL'}", flush=True)

    si, sf, su = b.CuArray(full_i[lo:hi]), b.CuArray(full_f[lo:hi]), b.CuArray(full_u[lo:hi])
    check("dist sum Int64", int(mg.dist_mapreduce(b.identity, b.add_sum, si)) == int(full_i.sum()))
    check("dist maximum UInt32->max via Int64", int(mg.dist_mapreduce(b.identity, b.max_, si)) == int(full_i.max()))
    fz = np.nan_to_num(full_f, nan=0.0)
    sfz = b.CuArray(fz[lo:hi])
    s = float(mg.dist_mapreduce(b.identity, b.add_sum, sfz))
    check("dist sum Float32 (tolerance)", abs(s - fz.astype(np.float64).sum()) <= float(oracle.sum_tolerance("identity", fz, (1,), np.float32)[0]))
    a2 = float(mg.dist_mapreduce(b.abs2, b.add, b.CuArray(fz[lo:hi].astype(oracle.bfloat16))))
    ref = (fz.astype(oracle.bfloat16).astype(np.float64) ** 2).sum()
    check("dist mapreduce(abs2,+) BFloat16", abs(a2 - ref) <= 2.0 ** -7 * ref)
    mn, mx = mg.dist_extrema(si)
    check("dist extrema Int64", (int(mn), int(mx)) == (int(full_i.min()), int(full_i.max())))
    mn, mx = mg.dist_extrema(sfz)
    check("dist extrema Float32", (float(mn), float(mx)) == (float(fz.min()), float(fz.max())))
    fb = np.zeros(n, dtype=np.bool_); fb[n - 3] = True                      # only the last rank holds a True
    sb = b.CuArray(fb[lo:hi])
    check("dist any / all / max / min over Bool shards",
          (bool(mg.dist_mapreduce(b.identity, b.or_, sb)), bool(mg.dist_mapreduce(b.identity, b.and_, sb)),
           bool(mg.dist_mapreduce(b.identity, b.max_, sb)), bool(mg.dist_mapreduce(b.identity, b.min_, sb))) == (True, False, True, False))
    check("dist maximum / minimum UInt32", (int(mg.dist_mapreduce(b.identity, b.max_, su)), int(mg.dist_mapreduce(b.identity, b.min_, su)))
          == (int(full_u.max()), int(full_u.min())))
    check("dist cumsum Int64", np.array_equal(gather(mg.dist_accumulate(b.add, si)), np.cumsum(full_i)))
    check("dist exclusive cumsum Int64", np.array_equal(gather(mg.dist_accumulate(b.add, si, exclusive=True)),
                                                         np.concatenate([[0], np.cumsum(full_i)[:-1]])))
    cs = gather(mg.dist_accumulate(b.add, sfz)).astype(np.float64)
    check("dist cumsum Float32 (tolerance)", np.all(np.abs(cs - np.cumsum(fz.astype(np.float64))) <= oracle.scan_tolerance(fz, 1, np.float32) + 1e-6))
    for rev in (False, True):
        srt, counts = mg.dist_sort(sf, rev=rev)
        got = gather(srt)
        check(f"dist sort! Float32 rev={rev}", np.array_equal(got.view(np.uint32)[~np.isnan(got)], oracle.sort(full_f, rev=rev).view(np.uint32)[~np.isnan(got)]) and got.size == n)
        perm = gather(mg.dist_sortperm(sf, rev=rev))
        check(f"dist sortperm Float32 rev={rev}", np.array_equal(perm, oracle.sortperm(full_f, rev=rev)))
    check("dist sort! UInt32", np.array_equal(gather(mg.dist_sort(su)[0]), np.sort(full_u)))
    check("dist sort! UInt32 rev", np.array_equal(gather(mg.dist_sort(su, rev=True)[0]), np.sort(full_u)[::-1]))
    dupk = (full_u % np.uint32(7)) * np.uint32(600_000_000)                  # 7 distinct keys: whole bins move together
    check("dist sort! UInt32, 7 distinct keys", np.array_equal(gather(mg.dist_sort(b.CuArray(dupk[lo:hi]))[0]), np.sort(dupk)))
    si32 = full_i.astype(np.int32)
    check("dist sort! Int32", np.array_equal(gather(mg.dist_sort(b.CuArray(si32[lo:hi]))[0]), np.sort(si32)))
    check("dist sortperm UInt32", np.array_equal(gather(mg.dist_sortperm(su)), oracle.sortperm(full_u)))
    if not timing:
        if rank == 0:
            print("ALL OK" if ok else "SOME FAILED", flush=True)
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    # timing of the sharded sort on device-resident shards (max over ranks)
    big = 1 << 28
    g = torch.Generator(device="cuda"); g.manual_seed(1 + rank)
    K = b.CuArray.from_torch(torch.randint(-2**31, 2**31 - 1, (big,), device="cuda", generator=g, dtype=torch.int32), (big,), b.dtypes.U32)
    mg.dist_sort(K); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); mg.dist_sort(K); torch.cuda.synchronize(); dist.barrier()
    dt_ = time.perf_counter() - t0
    if rank == 0:
        print(f"dist_sort UInt32 2^28 per rank x {world} ranks: {dt_*1e3:.2f} ms  {big * world / dt_ / 1e9:.1f} Gkeys/s (wall, incl. host logic)", flush=True)
        print("ALL OK" if ok else "SOME FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
