"""Stage timing of the device-driven multi-GPU sort / sortperm / cumsum (torchrun, N ranks): CUDA events between the
stages on every rank, max over ranks.  usage: torchrun ... tools/mgpu_profile.py [log2 of the TOTAL key count]"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200arrays as b
from b200arrays import multigpu as mg, dtypes as dt

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = (1 << lg) // world
g = torch.Generator(device=dev); g.manual_seed(1 + rank)
K = b.CuArray.from_torch(torch.randint(-2**31, 2**31 - 1, (n,), device=dev, generator=g, dtype=torch.int32), (n,), dt.U32)
ops = mg.DeviceOps()


def staged(fn_list, reps=4):
    """fn_list: [(name, fn)], run in order; returns {name: max-over-ranks ms} of the last rep."""
    res = None
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(fn_list) + 1)]
        t0 = time.perf_counter()
        evs[0].record()
        for i, (_, fn) in enumerate(fn_list):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(fn_list))] + [evs[0].elapsed_time(evs[-1]), wall]
        t = torch.tensor(ms, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res = t.tolist()
    return dict(zip([nm for nm, _ in fn_list] + ["device total", "wall total"], res))


st = {}
def s_hist(): st["hist"] = ops.mgpu_hist16(K, False)
def s_gather():
    st["all"] = torch.empty(world * 65536, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(st["all"], st["hist"].typed().view(torch.int32))
def s_plan(): st["plan"], st["hr"], st["info"] = ops.mgpu_plan(st["all"], world, rank)
def s_exchange(): st["recv"] = ops.mgpu_exchange(K, False, st["plan"], st["info"], None)
def s_sort(): st["out"] = ops.sort_keys_from(st["recv"], K.eltype, st["info"]["recv_total"], False, st["hr"])

r = staged([("hist16", s_hist), ("all-gather 256 KB/rank", s_gather), ("plan + 512 B read-back", s_plan),
            ("barrier + scatter (peer stores) + barrier", s_exchange), ("local sort (hybrid, plan's histogram)", s_sort)])
if rank == 0:
    print(f"dist sort! UInt32 2^{lg} total on {world} GPUs ({n} keys per rank):")
    for k, v in r.items():
        print(f"  {k:45s} {v:8.3f} ms")
    print(f"  -> {((1 << lg) / r['device total'] / 1e6):.1f} Gkeys/s (device), recv_total {st['info']['recv_total']}", flush=True)
r = staged([("dist_sort (one call)", lambda: mg.dist_sort(K))])
if rank == 0:
    print(f"  dist_sort as one call: device {r['device total']:.3f} ms, wall {r['wall total']:.3f} ms")
F = b.CuArray.from_torch(torch.rand(n, device=dev, generator=g, dtype=torch.float32))
r = staged([("dist_sortperm Float32", lambda: mg.dist_sortperm(F))], reps=3)
if rank == 0:
    print(f"  dist_sortperm Float32 2^{lg} total: device {r['device total']:.3f} ms = {((1 << lg) / r['device total'] / 1e6):.1f} Gkeys/s")
r = staged([("dist_accumulate Float32", lambda: mg.dist_accumulate(b.add, F))], reps=4)
if rank == 0:
    print(f"  dist cumsum Float32 2^{lg} total: device {r['device total']:.3f} ms = {8 * (1 << lg) / r['device total'] / 1e6:.0f} GB/s")
L = K.copy()
for _ in range(2):
    L.buf.copy_(K.buf); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.sort_keys(L, False); e1.record(); e1.synchronize()
if rank == 0:
    print(f"  single-GPU sort! of one shard ({n} keys, full key range): {e0.elapsed_time(e1):.3f} ms", flush=True)
dist.destroy_process_group()
