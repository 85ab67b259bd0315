"""Ten-second sanity check of the contiguous scan host paths (chain2, look-back tiling, unaligned fallback)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200arrays as b
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
ok = True
for n in (5, 13441, (1 << 22) + 77):
    I = torch.randint(-2**20, 2**20, (n,), device=dev, generator=g, dtype=torch.int64)
    ok &= bool(torch.equal(b.accumulate(b.add, b.CuArray.from_torch(I)).buf.view(torch.int64)[:n], torch.cumsum(I, 0)))
    J = torch.randint(-2**30, 2**30, (n,), device=dev, generator=g, dtype=torch.int32)
    ok &= bool(torch.equal(b.accumulate(b.add, b.CuArray.from_torch(J)).buf.view(torch.int32)[:n], torch.cumsum(J, 0, dtype=torch.int32)))
for rows, cols in ((16384, 64), (50_001, 37), (13440 * 2, 5)):
    M = torch.randint(-1000, 1000, (cols, rows), device=dev, generator=g, dtype=torch.int64)
    got = b.accumulate(b.add, b.CuArray.from_torch(M, (rows, cols)), dims=1, init=7).buf.view(torch.int64).reshape(cols, rows)
    ok &= bool(torch.equal(got, torch.cumsum(M, 1) + 7))
base = torch.randint(-1000, 1000, (100_003,), device=dev, generator=g, dtype=torch.int32)
d = b.CuArray.from_torch(base)
v = b.CuArray(d.buf[4:], (base.numel() - 1,), d.eltype)
ok &= bool(torch.equal(b.accumulate(b.add, v).buf.view(torch.int32)[:base.numel() - 1], torch.cumsum(base[1:], 0, dtype=torch.int32)))
print("QUICK SCAN CHECK", "OK" if ok else "FAILED", flush=True)
