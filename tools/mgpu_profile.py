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
    srt, _ = mg.dist_sort(K); t1 = T()
    if rank == 0 and rep == 2:
        print(f"dist_sort (default path, fused={mg.FUSED_EXCHANGE}) n/rank 2^{int(np.log2(big))} x {world}: {1e3*(t1-t0):.2f} ms  {big*world/(t1-t0)/1e9:.1f} Gkeys/s", flush=True)
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
# pieces of _exchange
dev = parted.buf.device
G = world
send_counts = torch.tensor(np.diff(bounds), dtype=torch.int64)
dist.barrier(); t0 = T()
sc = send_counts.to(dev); ac = [torch.empty(G, dtype=torch.int64, device=dev) for _ in range(G)]
dist.all_gather(ac, sc); cm = torch.stack(ac).cpu().numpy(); t1 = T()
recv_counts = cm[:, rank]; n_recv = int(recv_counts.sum())
t_in = parted.buf.view(torch.int32); t_out = torch.empty(n_recv, dtype=torch.int32, device=dev); t2 = T()
dist.all_to_all_single(t_out, t_in, output_split_sizes=[int(x) for x in recv_counts], input_split_sizes=[int(x) for x in send_counts]); t3 = T()
dist.all_to_all_single(t_out, t_in, output_split_sizes=[int(x) for x in recv_counts], input_split_sizes=[int(x) for x in send_counts]); t4 = T()
if rank == 0:
    print(f"exchange pieces: counts all_gather+cpu {1e3*(t1-t0):.2f} | alloc {1e3*(t2-t1):.2f} | uneven a2a first {1e3*(t3-t2):.2f} | uneven a2a again {1e3*(t4-t3):.2f} ms", flush=True)
# all_to_all_single alone, even splits of the same volume
per = (big // world)
src = torch.empty(per * world, dtype=torch.int32, device="cuda"); dst = torch.empty_like(src)
for _ in range(2):
    dist.all_to_all_single(dst, src)
dist.barrier(); t0 = T()
for _ in range(5):
    dist.all_to_all_single(dst, src)
t1 = T()
if rank == 0:
    vol = src.numel() * 4 * (world - 1) / world
    print(f"all_to_all_single even, {src.numel()*4/1e9:.2f} GB per rank: {1e3*(t1-t0)/5:.2f} ms  {vol/((t1-t0)/5)/1e9:.0f} GB/s out per GPU "
          f"(NCCL_MIN_P2P_NCHANNELS={os.environ.get('NCCL_MIN_P2P_NCHANNELS')}, NCCL_NCHANNELS_PER_NET_PEER={os.environ.get('NCCL_NCHANNELS_PER_NET_PEER')})", flush=True)
L = K.copy(); t0 = T(); ops.sort_keys(L, False); t1 = T()
if rank == 0:
    print(f"single-GPU sort of 2^{int(np.log2(big))}: {1e3*(t1-t0):.2f} ms  {big/(t1-t0)/1e9:.1f} Gkeys/s", flush=True)
dist.destroy_process_group()
