import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import b200arrays as b
from b200arrays import dtypes as dt
from microbench import rand_array
n = 1 << 30
V = rand_array(dt.F32, (n,)); O = b.CuArray.undef(dt.F32, (n,))
dbg = torch.zeros(1024 * 8, dtype=torch.int64, device="cuda")
b.accumulate_(b.add, O, V); torch.cuda.synchronize()
os.environ["B200_SCAN_DBGPTR"] = hex(dbg.data_ptr())
b.accumulate_(b.add, O, V); torch.cuda.synchronize()
r = dbg.view(-1, 8).cpu().numpy()
r = r[r[:, 6] > 0]
its = r[:, 6]
names = ["wait_full+snapshot", "phaseA(s)", "pass1+warp scans", "lookback", "pass2+store", "acquire"]
print("CTAs", len(r), "iters/CTA mean", its.mean())
for i, nm in enumerate(names):
    print(f"{nm:22s} per-iter mean {np.mean(r[:, i] / its):8.1f} ns")
print("total per iter", np.mean(r[:, :6].sum(1) / its))
