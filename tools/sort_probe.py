"""Time keys-only sort! of n UInt32 / Float32 keys (B200_SORT_CTR etc. select the pass kernel)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200arrays as b
from b200arrays import dtypes as dt

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 1 << logn
torch.cuda.set_device(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
src = torch.randint(-2**31, 2**31 - 1, (n,), device="cuda", dtype=torch.int32, generator=g)
for e, name in ((dt.U32, "U32"), (dt.F32, "F32")):
    K = b.CuArray(src.clone().view(torch.uint8), (n,), e)
    ts = []
    for rep in range(5):
        K.buf.copy_(src.view(torch.uint8))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.sort_(K); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = K.buf.view(torch.int32)
    if e == dt.U32:
        # unsigned order == signed order of (x ^ 0x80000000)
        r = (res ^ -2**31)
        ok = bool((r[1:] >= r[:-1]).all()) and int(res.sum(dtype=torch.int64)) == int(src.sum(dtype=torch.int64))
    else:
        f = res.view(torch.float32)
        fin = ~torch.isnan(f)
        nn = int(fin.sum())
        ok = bool((f[1:nn] >= f[:nn - 1]).all()) and bool(torch.isnan(f[nn:]).all())
    print(f"sort! 2^{logn} {name} CTR={os.environ.get('B200_SORT_CTR','0')}: min {min(ts):.3f} ms  med {sorted(ts)[2]:.3f}  {n / min(ts) / 1e6:.1f} Gkeys/s  sorted_ok={ok}", flush=True)
