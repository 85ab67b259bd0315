"""Stage timing of the partition-first sample sort (torchrun, N ranks)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200arrays as b
from b200arrays import multigpu as mg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
big = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
g = torch.Generator(device="cuda"); g.manual_seed(1 + rank)
K = b.CuArray.from_torch(torch.randint(-2**31, 2**31 - 1, (big,), device="cuda", generator=g, dtype=torch.int32), (big,), b.dtypes.U32)
ops = mg.DeviceOps()

def T():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(3):
    dist.barrier(); t0 = T()
    spl = mg._splitters(K, False, None, ops); t1 = T()
    parted, _, counts = ops.partition(K, None, spl, world, False); t2 = T()
    bounds = np.concatenate([[0], np.cumsum(counts.to_host().astype(np.int64))]); t3 = T()
    (recv,), cmat = mg._exchange(parted, bounds, None); t4 = T()
    ops.sort_keys(recv, False); t5 = T()
    dist.barrier(); t6 = T()
    if rank == 0 and rep == 2:
        print(f"n/rank 2^{int(np.log2(big))} x {world}: splitters {1e3*(t1-t0):.2f} | partition {1e3*(t2-t1):.2f} | counts->host {1e3*(t3-t2):.2f} | "
              f"exchange {1e3*(t4-t3):.2f} | local sort ({recv.length} keys) {1e3*(t5-t4):.2f} | total {1e3*(t6-t0):.2f} ms  "
              f"{big*world/(t6-t0)/1e9:.1f} Gkeys/s", flush=True)
L = K.copy(); t0 = T(); ops.sort_keys(L, False); t1 = T()
if rank == 0:
    print(f"single-GPU sort of 2^{int(np.log2(big))}: {1e3*(t1-t0):.2f} ms  {big/(t1-t0)/1e9:.1f} Gkeys/s", flush=True)
dist.destroy_process_group()
