"""Boundary behaviour the C ABI promises (include/b200arrays.h): stream-ordered,
asynchronous, CUDA-graph capturable (no allocation / synchronisation inside the library; control
words cleared by memset nodes), safe with several streams in flight, status codes for bad input."""
import ctypes

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(0xAB1)


def test_cuda_graph_capture_and_replay(b):
    import torch
    n = 1 << 20
    a = RNG.integers(-1000, 1000, size=n).astype(np.int64)
    f = RNG.random(n).astype(np.float32)
    dA, dF = b.CuArray(a), b.CuArray(f)
    R = b.CuArray.undef(b.dtypes.I64, (1,))
    O = b.CuArray.undef(b.dtypes.I64, (n,))
    K = b.CuArray.undef(b.dtypes.F32, (n,))
    P = b.CuArray.undef(b.dtypes.I64, (n,))

    def work():
        b.mapreducedim_(b.identity, b.add_sum, R, dA, init=0)
        b.accumulate_(b.add, O, dA)
        K.buf.copy_(dF.buf)
        b.sort_(K)
        b.sortperm_(P, dF)

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        work()                                  # warm-up outside capture (function attributes, allocator pools)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        work()
    # new inputs, same buffers: replay must recompute
    a2 = RNG.integers(-1000, 1000, size=n).astype(np.int64)
    f2 = RNG.random(n).astype(np.float32)
    dA.buf.copy_(b.CuArray(a2).buf)
    dF.buf.copy_(b.CuArray(f2).buf)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert b.Array(R)[0] == a2.sum()
    np.testing.assert_array_equal(b.Array(O), np.cumsum(a2))
    np.testing.assert_array_equal(b.Array(K), np.sort(f2))
    np.testing.assert_array_equal(b.Array(P), oracle.sortperm(f2))


def test_concurrent_streams(b):
    """Many tasks / streams in flight at once (docs/src/usage/multitasking.md:10): every call owns its workspace."""
    import torch
    streams = [torch.cuda.Stream() for _ in range(4)]
    n = (1 << 21) + 17
    datas = [RNG.integers(-2**40, 2**40, size=n, dtype=np.int64) for _ in streams]
    keys = [RNG.random(n).astype(np.float32) for _ in streams]
    outs = []
    dev = [(b.CuArray(d), b.CuArray(k)) for d, k in zip(datas, keys)]
    torch.cuda.synchronize()
    for st, (dd, kk) in zip(streams, dev):
        with torch.cuda.stream(st):
            for _ in range(3):
                c = b.cumsum(dd)
                sm = b.sum(dd, dims=1)
                p = b.sortperm(kk)
            outs.append((c, sm, p))
    torch.cuda.synchronize()
    for (c, sm, p), d, k in zip(outs, datas, keys):
        np.testing.assert_array_equal(b.Array(c), np.cumsum(d))
        assert b.Array(sm)[0] == d.sum()
        np.testing.assert_array_equal(b.Array(p), oracle.sortperm(k))


def test_status_codes(b):
    import torch
    lib = b._lib.load()
    x = b.CuArray(np.ones(1 << 22, dtype=np.float32))
    r = b.CuArray.undef(b.dtypes.F32, (1,))
    need = ctypes.c_size_t(0)
    assert lib.b200_reduce_workspace_bytes(2, 2, 0, 0, 1, x.length, 1, ctypes.byref(need)) == 0 and need.value > 0
    small = torch.empty(16, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert lib.b200_reduce(small.data_ptr(), 16, x.pointer, r.pointer, 2, 2, 0, 0, 1, x.length, 1, 0, st) == 3   # WORKSPACE_TOO_SMALL
    assert lib.b200_reduce(None, 0, x.pointer, r.pointer, 2, 2, 0, 0, 1, x.length, 1, 0, st) in (1, 3)
    assert lib.b200_reduce(None, 0, None, r.pointer, 2, 2, 0, 0, 1, 8, 1, 0, st) == 1                           # null input
    assert lib.b200_reduce(None, 0, x.pointer, r.pointer, 2, 2, 4, 0, 1, 8, 1, 0, st) == 2                      # & on floats
    with pytest.raises(b.B200ArraysError):
        b._lib.check(2)
    assert lib.b200_check_device() == 0
