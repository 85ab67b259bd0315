"""The hybrid MSD path of keys-only `sort!` (csrc/sort_msd.cuh) on the input classes that drive its branches, at the
smallest size it is attempted for (2^27 keys): dense uniform slices, Float32 `rand` / `randn` (heavy digits in the
partition passes, segments of 10^5..10^6 keys in the 16-bit leaf), a two-population mix, a few values with more than
65535 copies each (the 16-bit leaf's spill table), 65536 values x 2048 copies (8-bit bins overflow), presorted input
(row-uniform fast paths) and inputs the plan kernel must decline (small integers, constant).  The result is compared
bit for bit with a sort of the same data by torch on the device (a library used as the test's checker only; the oracle
cannot sort 2^27 keys in seconds) — exact equality is the reference's own bar for keys-only sorts
(test/core/sorting.jl:285-296: `sort!(c) == sort!(Array(c))`)."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu

N = 1 << 27


def _gen(torch, n, dist, seed=11):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    ri = lambda lo, hi: torch.randint(lo, hi, (n,), device="cuda", dtype=torch.int32, generator=g)
    if dist == "narrow8":
        return ri(0, 2**29) + (3 << 29)
    if dist == "randf":
        return torch.rand(n, device="cuda", generator=g).view(torch.int32)
    if dist == "randnf":
        return torch.randn(n, device="cuda", generator=g).view(torch.int32)
    if dist == "skew2":
        x = ri(-2**31, 2**31 - 1)
        return torch.where(x > 0, x, ri(0, 2**24) + (0x5A << 24))
    if dist == "bigval":
        return (ri(0, 1024) * 4194301) ^ -2**31
    if dist == "dup16":
        return ri(0, 2**16) * 65537
    if dist == "sorted":
        return torch.arange(n, device="cuda", dtype=torch.int64).mul_(3).to(torch.int32)
    if dist == "lowbits":
        return ri(0, 2**20)
    if dist == "const":
        return torch.full((n,), 123456789, device="cuda", dtype=torch.int32)
    raise ValueError(dist)


def _expected(torch, src, dtype, rev):
    if dtype == "uint32":
        return torch.sort(src ^ -2**31, descending=rev).values ^ -2**31
    if dtype == "int32":
        return torch.sort(src, descending=rev).values
    key = torch.where(src < 0, ~src, src | -2**31) ^ -2**31          # Float32 bit patterns -> signed-comparable total order (no NaNs here)
    return src[torch.sort(key, descending=rev, stable=True).indices]


@pytest.mark.parametrize("dist,dtype,rev,hybrid", [
    ("narrow8", "uint32", False, True),
    ("randf", "float32", False, True),
    ("randf", "float32", True, True),
    ("randnf", "float32", False, True),
    ("skew2", "uint32", False, True),
    ("skew2", "int32", True, True),
    ("bigval", "uint32", False, True),
    ("dup16", "uint32", False, None),        # admitted from 2^28 (2048 keys per segment here): either path must be exact
    ("sorted", "uint32", False, True),
    ("lowbits", "uint32", False, False),
    ("const", "uint32", False, False),
])
def test_hybrid_path_input_classes(b, dist, dtype, rev, hybrid):
    import torch
    lib = b._lib.load()
    dt = b.dtypes
    et = {"uint32": dt.U32, "int32": dt.I32, "float32": dt.F32}[dtype]
    src = _gen(torch, N, dist)
    keys = src.clone()
    nb = ctypes.c_size_t(0)
    b._lib.check(lib.b200_sort_workspace_bytes(et, -1, N, 1, ctypes.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    ws.fill_(0xA5)                                                     # the library must not rely on zeroed workspace
    stream = torch.cuda.current_stream().cuda_stream
    b._lib.check(lib.b200_sort_keys(ws.data_ptr(), nb.value, keys.data_ptr(), et, N, 1, int(rev), stream))
    torch.cuda.synchronize()
    mode = int(ws[:4].view(torch.int32).item())                        # SortCtl.mode: 1 = the hybrid path took the call
    if hybrid is not None:
        assert (mode != 0) == hybrid, f"plan kernel decision for {dist}: mode={mode}"
    want = _expected(torch, src, dtype, rev)
    assert torch.equal(keys, want), f"{dist}/{dtype}: {int((keys != want).sum())} keys differ"
