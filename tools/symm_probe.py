"""Probe: does torch symmetric memory (P2P-mapped buffers across processes) work on this box?"""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm_mem
    n = 1 << 28
    t = symm_mem.empty(n, dtype=torch.int32, device=f"cuda:{local}")
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok; handle attrs:", [a for a in dir(hdl) if not a.startswith("_")][:40], flush=True)
    t.fill_(rank)
    hdl.barrier()
    peer = (rank + 1) % world
    pb = hdl.get_buffer(peer, (n,), torch.int32, 0)
    print(rank, "peer ptr", hex(pb.data_ptr()), "own ptr", hex(t.data_ptr()), "peer[0]=", int(pb[0].item()), flush=True)
    src = torch.full((n,), 100 + rank, dtype=torch.int32, device=f"cuda:{local}")
    torch.cuda.synchronize(); hdl.barrier()
    t0 = time.perf_counter()
    pb.copy_(src)                    # store into the peer's memory over NVLink
    torch.cuda.synchronize()
    dt_ = time.perf_counter() - t0
    hdl.barrier()
    print(rank, f"peer write {n*4/1e9:.2f} GB in {dt_*1e3:.2f} ms = {n*4/dt_/1e9:.0f} GB/s; my buffer now holds", int(t[0].item()), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "SYMM FAILED:", repr(e), flush=True)
dist.destroy_process_group()
