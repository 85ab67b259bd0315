"""2-GPU debug: dist_sort / dist_accumulate at bench sizes with full tracebacks (torchrun)."""
import os, sys, traceback
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200arrays as b
from b200arrays import dtypes as dt, multigpu as mg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = (1 << lg) // world
g = torch.Generator(device=dev); g.manual_seed(100 + rank)

def say(*a):
    print(f"[rank {rank}]", *a, flush=True)

try:
    I = torch.randint(-2**20, 2**20, (n,), device=dev, generator=g, dtype=torch.int64)
    r = mg.dist_accumulate(b.add, b.CuArray.from_torch(I)).typed()
    part = I.sum()
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    carry = sum(int(p.item()) for p in parts[:rank])
    ref = torch.cumsum(I, 0) + carry
    say("cumsum i64 equal:", bool(torch.equal(r, ref)), "last", int(r[-1].item()), int(ref[-1].item()), "first", int(r[0].item()), int(ref[0].item()),
        "mismatches", int((r != ref).sum().item()))
    del I, r, ref
except Exception:
    say(traceback.format_exc())
torch.cuda.empty_cache()
try:
    U = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, generator=g, dtype=torch.int32)
    free, total = torch.cuda.mem_get_info()
    say("mem free/total GB", free / 2**30, total / 2**30)
    out, info = mg.dist_sort(b.CuArray.from_torch(U, (n,), dt.U32))
    t = out.typed().view(torch.int32) ^ -2**31
    say("sort ok", bool((t[1:] >= t[:-1]).all()), "recv_total", info["recv_total"] if isinstance(info, dict) else info)
    torch.cuda.synchronize()
    import time
    for _ in range(3):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        mg.dist_sort(b.CuArray.from_torch(U, (n,), dt.U32))
        torch.cuda.synchronize(); dist.barrier()
        say("dist_sort wall ms", (time.perf_counter() - t0) * 1e3)
except Exception:
    say(traceback.format_exc())
dist.destroy_process_group()
