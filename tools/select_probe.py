"""partialsort(c, k) by radix select vs full sort at 2^30 Float32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200arrays as b
from b200arrays import dtypes as dt
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 30)
torch.cuda.set_device(0)
x = torch.rand(n, device="cuda")
c = b.CuArray.from_torch(x, (n,))
for k in (1, n // 2, n):
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); v = b.partialsort(c, k); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"partialsort(c, {k}) n=2^{n.bit_length()-1} F32: {min(ts):.3f} ms  value {v}  ({n * 4 * 4 / min(ts) / 1e6:.0f} GB/s over 4 reads)", flush=True)
kv = torch.kthvalue(x, n // 2)
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); kv = torch.kthvalue(x, n // 2); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"torch.kthvalue: {min(ts):.3f} ms value {float(kv.values)}")
