"""Does dist._coalescing_manager with mixed reduce ops disturb later collectives?  (2-GPU probe)"""
import os, sys
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
def probe(tag):
    part = torch.full((), 1000.0 + rank, device=dev, dtype=torch.float64)
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    big = torch.full((1 << 20,), rank + 1, device=dev, dtype=torch.int32)
    out = torch.empty((world << 20,), device=dev, dtype=torch.int32)
    dist.all_gather_into_tensor(out, big)
    print(f"[{tag} rank {rank}] all_gather 0-dim: {[p.item() for p in parts]}  into_tensor sums: {[int(out[k << 20:(k + 1) << 20].sum()) for k in range(world)]}", flush=True)
probe("before")
a = torch.ones(16384, device=dev); b2 = torch.ones(16384, device=dev, dtype=torch.bfloat16); c = torch.ones(1, device=dev, dtype=torch.int64)
m1 = torch.full((1,), float(rank), device=dev); m2 = torch.full((1,), float(rank), device=dev, dtype=torch.bfloat16)
for _ in range(3):
    with dist._coalescing_manager(device=dev):
        dist.all_reduce(a); dist.all_reduce(b2)
        dist.all_reduce(m1, op=dist.ReduceOp.MAX); dist.all_reduce(m2, op=dist.ReduceOp.MAX)
        dist.all_reduce(c)
torch.cuda.synchronize()
print(f"[rank {rank}] after coalesced: a[0]={a[0].item()} (3 rounds of SUM over {world}: {world**3}), m1={m1.item()} (MAX: {world-1}), c={c.item()}", flush=True)
probe("after")
dist.destroy_process_group()
