"""A few launches of the kernels VERDICT r1 asked ncu rows for (run under ncu): strided look-back scan, argreduce,
fused findall, pairs LSD pass.  python tools/prof_misc.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200arrays as b
from b200arrays import dtypes as dt
torch.cuda.set_device(0)
g = torch.Generator(device="cuda"); g.manual_seed(5)
M = 16384
A = b.CuArray.from_torch(torch.rand(M * M, device="cuda", generator=g), (M, M))
O = b.CuArray.undef(dt.F32, (M, M))
for _ in range(2):
    b.accumulate_(b.add, O, A, dims=2)          # scan_strided_lb_kernel
del O
n = 1 << 28
V = b.CuArray.from_torch(torch.rand(n, device="cuda", generator=g), (n,))
for _ in range(2):
    b.findmax(V)                                # argreduce kernels
F = b.CuArray.from_torch(torch.rand(n, device="cuda", generator=g) < 0.3, (n,))
for _ in range(2):
    b.findall(F)                                # select_flagged_kernel
for _ in range(1):
    b.sortperm(V)                               # radix_onesweep_kernel, pairs
torch.cuda.synchronize()
print("done")
