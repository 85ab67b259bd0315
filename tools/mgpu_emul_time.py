"""Time b200_mgpu_scatter on ONE GPU with G emulated ranks (all receive buffers local): separates the kernel's own
cost from the NVLink store path.  usage: python tools/mgpu_emul_time.py [log2 keys per rank] [G]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200arrays as b
from b200arrays import _lib, dtypes as dt
lib = _lib.load()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = 1 << lg
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda"); g.manual_seed(3)
keys = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, dtype=torch.int32, generator=g)
hist_all = torch.empty(G * 65536, dtype=torch.int32, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(3):
    e0, e1 = ev(), ev(); e0.record()
    _lib.check(lib.b200_mgpu_hist16(keys.data_ptr(), dt.U32, n, 0, 0, hist_all.data_ptr(), stream))
    e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"hist16 {min(ts):.3f} ms")
for r in range(1, G):
    hist_all[r * 65536:(r + 1) * 65536] = hist_all[:65536]
nb = ctypes.c_size_t(0); _lib.check(lib.b200_mgpu_plan_bytes(ctypes.byref(nb)))
plan = torch.empty(nb.value, dtype=torch.uint8, device=dev); hr = torch.empty(65536, dtype=torch.int32, device=dev)
ts = []
for _ in range(3):
    e0, e1 = ev(), ev(); e0.record()
    _lib.check(lib.b200_mgpu_plan(hist_all.data_ptr(), G, 0, plan.data_ptr(), hr.data_ptr(), stream))
    e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"plan {min(ts):.3f} ms")
cap = int(1.1 * n) + 65536
recv = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(G)]
ptrs = (ctypes.c_void_p * G)(*[t.data_ptr() for t in recv])
w = ctypes.c_size_t(0); _lib.check(lib.b200_mgpu_scatter_workspace_bytes(n, ctypes.byref(w)))
ws = torch.empty(w.value, dtype=torch.uint8, device=dev)
ts = []
for _ in range(4):
    _lib.check(lib.b200_mgpu_plan(hist_all.data_ptr(), G, 0, plan.data_ptr(), hr.data_ptr(), stream))   # resets the ticket
    torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record()
    _lib.check(lib.b200_mgpu_scatter(ws.data_ptr(), w.value, keys.data_ptr(), dt.U32, n, 0, plan.data_ptr(), ptrs, G, stream))
    e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"scatter 2^{lg} keys into {G} local buffers: {min(ts):.3f} ms = {n / min(ts) / 1e6:.1f} Gkeys/s")
