"""CPU oracle for the grafted reduce / scan / sort path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module; the product (cuda.jl_b200/) never does.

What it restates (paths relative to the JuliaGPU/CUDA.jl v6.3.0 tree):
  * GPUArrays.mapreducedim! on CuArray      CUDACore/src/mapreduce.jl:168-301
  * CUDACore.scan! / Base._accumulate!       CUDACore/src/accumulate.jl:135-234
  * Base.sort! / sortperm! / partialsort!    CUDACore/src/sorting.jl:993-1070, :574-592
  * device max/min (IEEE-754-2019)           CUDACore/src/device/intrinsics/math.jl:458-487
plus the Julia Base / GPUArrays.jl (>= 11.5.4, CUDACore/Project.toml:56; external, not
under /root/reference) semantics those call sites rely on: add_sum / mul_prod
widening, neutral elements, `isless` total order, stable sortperm, 1-based Int64
indices (SURVEY §8 a1').

PARITY UNPINNED for random-data cases: the reference is Julia; no julia binary exists
in this image, and its tests hold no stored numeric files — every pin is "equal to
Base Julia on rand() data" (test/setup.jl:112).  What IS pinned, and checked in
tests/test_oracle_golden.py, are the reference tests with literal expected values
(tests/golden/reference_literals.json, each with its file:line) and the
equality-class of every other test (bit-exact vs tolerance, shapes, edge sizes).
"""
import math

import numpy as np

try:
    import ml_dtypes
    bfloat16 = np.dtype(ml_dtypes.bfloat16)
except Exception:  # pragma: no cover
    bfloat16 = None


# ----------------------------------------------------------------------------- dtypes
def is_float(dtype):
    dtype = np.dtype(dtype)
    return dtype.kind == "f" or (bfloat16 is not None and dtype == bfloat16)


def eps(dtype):
    dtype = np.dtype(dtype)
    if bfloat16 is not None and dtype == bfloat16:
        return 2.0 ** -7
    return float(np.finfo(dtype).eps)


def widen_add(dtype):
    """Base.add_sum / mul_prod result type: BitSignedSmall -> Int, BitUnsignedSmall -> UInt, Bool -> Int."""
    dtype = np.dtype(dtype)
    if dtype == np.bool_ or (dtype.kind == "i" and dtype.itemsize < 8):
        return np.dtype(np.int64)
    if dtype.kind == "u" and dtype.itemsize < 8:
        return np.dtype(np.uint64)
    return dtype


# ----------------------------------------------------------------------------- reduce
def _factor(adims, rdims):
    nd = len(adims)
    rdims = tuple(rdims) + (1,) * (nd - len(rdims))          # mapreduce.jl:187-190
    axes = tuple(d for d in range(nd) if rdims[d] == 1 and adims[d] != 1)   # mapreduce.jl:196
    return axes, rdims


def ieee_max(a, b):
    """IEEE-754-2019 maximum: NaN-propagating, -0.0 < +0.0 (math.jl:458-487)."""
    r = np.maximum(a, b)                                      # np.maximum propagates NaN
    both_zero = (a == 0) & (b == 0)
    if np.any(both_zero) and is_float(np.asarray(a).dtype):
        r = np.where(both_zero, np.where(np.signbit(a) & np.signbit(b), a, np.abs(a)), r)
    return r


def ieee_min(a, b):
    r = np.minimum(a, b)
    both_zero = (a == 0) & (b == 0)
    if np.any(both_zero) and is_float(np.asarray(a).dtype):
        r = np.where(both_zero, np.where(np.signbit(a) | np.signbit(b), -np.abs(a), a), r)
    return r


def apply_map(f, A):
    """f in {'identity', 'abs2'}; abs2 on integers wraps in the input type."""
    if f == "identity":
        return A
    if f == "abs2":
        with np.errstate(over="ignore"):
            return A * A
    raise ValueError(f)


def mapreducedim(f, op, A, rdims, out_dtype, exact_order=False):
    """R = mapreducedim!(f, op, R, A; init=neutral) for a dense column-major A.

    Integers / Bool / max / min: the exact value (wrap-around two's complement, order
    independent).  Float add / mul: evaluated in Float64 with numpy's pairwise summation
    (the "Float64 pairwise reference" of the north star) and NOT rounded, so callers
    compare with `sum_tolerance`.  exact_order=True instead accumulates sequentially in
    index order in out_dtype from op(neutral, neutral), which is what the reference's
    serial kernel does (mapreduce.jl:136-157) and its test pins with `==`
    (test/core/array.jl:771-773).
    """
    A = np.asarray(A)
    out_dtype = np.dtype(out_dtype)
    axes, rdims = _factor(A.shape, rdims)
    M = apply_map(f, A)
    if op == "extrema":
        mn = mapreducedim("identity", "min", M, rdims, out_dtype)
        mx = mapreducedim("identity", "max", M, rdims, out_dtype)
        return np.stack([mn, mx], axis=0)
    if not axes:
        return M.astype(out_dtype).reshape(rdims)
    flt = is_float(out_dtype)
    if op in ("add", "mul"):
        if flt and not exact_order:
            W = M.astype(np.float64)
            R = W.sum(axis=axes, keepdims=True) if op == "add" else W.prod(axis=axes, keepdims=True)
            return R.reshape(rdims)
        if flt and exact_order:
            assert len(axes) == 1
            ax = axes[0]
            W = np.moveaxis(M.astype(out_dtype), ax, 0)
            ne = out_dtype.type(0 if op == "add" else 1)
            acc = np.full(W.shape[1:], ne + ne if op == "add" else ne * ne, dtype=out_dtype)
            for r in range(W.shape[0]):
                acc = (acc + W[r]).astype(out_dtype) if op == "add" else (acc * W[r]).astype(out_dtype)
            return np.expand_dims(acc, ax).reshape(rdims)
        with np.errstate(over="ignore"):
            W = M.astype(out_dtype)
            ufunc = np.add if op == "add" else np.multiply
            return ufunc.reduce(W, axis=axes, dtype=out_dtype, keepdims=True).reshape(rdims)
    W = M.astype(out_dtype) if not (bfloat16 is not None and out_dtype == bfloat16) else M.astype(np.float32)
    if op in ("max", "min"):
        if flt:
            # NaN-propagating; signed zeros resolved through the total-order key
            Wf = W.astype(np.float64)
            fn = np.max if op == "max" else np.min
            R = fn(Wf, axis=axes, keepdims=True)              # propagates NaN
            zero = (R == 0)
            if np.any(zero):
                neg = np.signbit(Wf) & (Wf == 0)
                pos = (~np.signbit(Wf)) & (Wf == 0)
                if op == "max":
                    anypos = pos.any(axis=axes, keepdims=True)
                    R = np.where(zero, np.where(anypos, 0.0, -0.0), R)
                else:
                    anyneg = neg.any(axis=axes, keepdims=True)
                    R = np.where(zero, np.where(anyneg, -0.0, 0.0), R)
            return R.astype(out_dtype).reshape(rdims)
        fn = np.max if op == "max" else np.min
        return fn(W, axis=axes, keepdims=True).reshape(rdims)
    if op in ("and", "or"):
        if out_dtype == np.bool_:
            fn = np.all if op == "and" else np.any
            return fn(W, axis=axes, keepdims=True).reshape(rdims)
        ufunc = np.bitwise_and if op == "and" else np.bitwise_or
        return ufunc.reduce(W, axis=axes, keepdims=True).reshape(rdims)
    raise ValueError(op)


def sum_tolerance(f, A, rdims, T):
    """Absolute bound per output: 4*log2(n)*eps(T) * sum(|f(x)|) over the reduced slice,
    i.e. relative error <= 4*log2(n)*eps(T) w.r.t. the Float64 pairwise reference for
    same-sign data (BASELINE north_star), plus one rounding of the result to T."""
    A = np.asarray(A)
    axes, rdims = _factor(A.shape, rdims)
    n = int(np.prod([A.shape[a] for a in axes])) if axes else 1
    mag = np.abs(apply_map(f, A.astype(np.float64))).sum(axis=axes, keepdims=True).reshape(rdims)
    return (4.0 * max(1.0, math.log2(max(n, 2))) * eps(T)) * mag + np.finfo(np.float64).tiny


# ----------------------------------------------------------------------------- scan
def accumulate(op, A, dims=1, init=None, out_dtype=None, exclusive=False):
    """scan!(op, out, A; dims, init) (accumulate.jl:135-199); dims is 1-based.

    init (the x of Some(x)) is folded into every output exactly once
    (accumulate.jl:90-92,123-125).  Floats under add/mul are evaluated in Float64
    (compare with scan_tolerance); everything else is exact in out_dtype.
    exclusive=True is the variant BASELINE config 3 names (not reachable from the
    reference's public API, accumulate.jl:152,177): out[k] = op(init|neutral, A[1..k-1]).
    """
    A = np.asarray(A)
    out_dtype = np.dtype(out_dtype if out_dtype is not None else A.dtype)
    if dims < 1:
        raise ValueError("dims must be a positive integer")     # accumulate.jl:137
    if dims > A.ndim:
        return A.astype(out_dtype)                               # accumulate.jl:140 (copy)
    ax = dims - 1
    if A.shape[ax] == 0:
        return A.astype(out_dtype)
    flt = is_float(out_dtype)
    work = np.float64 if (flt and op in ("add", "mul")) else (np.float32 if (bfloat16 is not None and out_dtype == bfloat16) else out_dtype)
    W = A.astype(work)
    with np.errstate(over="ignore"):
        if op == "add":
            R = np.add.accumulate(W, axis=ax, dtype=work)
        elif op == "mul":
            R = np.multiply.accumulate(W, axis=ax, dtype=work)
        elif op == "max":
            R = np.maximum.accumulate(W, axis=ax)
        elif op == "min":
            R = np.minimum.accumulate(W, axis=ax)
        elif op == "and":
            R = np.bitwise_and.accumulate(W, axis=ax) if out_dtype != np.bool_ else np.logical_and.accumulate(W, axis=ax)
        elif op == "or":
            R = np.bitwise_or.accumulate(W, axis=ax) if out_dtype != np.bool_ else np.logical_or.accumulate(W, axis=ax)
        else:
            raise ValueError(op)
        if exclusive:
            ne = {"add": 0, "mul": 1, "or": 0}.get(op)
            if ne is None:
                if op == "and":
                    ne = True if out_dtype == np.bool_ else np.array(-1).astype(work)
                elif flt:
                    ne = -np.inf if op == "max" else np.inf
                else:
                    ii = np.iinfo(out_dtype)
                    ne = ii.min if op == "max" else ii.max
            first = np.full_like(np.take(R, [0], axis=ax), ne)
            R = np.concatenate([first, np.take(R, range(0, A.shape[ax] - 1), axis=ax)], axis=ax)
        if init is not None:
            iv = np.asarray(init).astype(work)
            if op == "add":
                R = iv + R
            elif op == "mul":
                R = iv * R
            elif op == "max":
                R = np.maximum(iv, R)
            elif op == "min":
                R = np.minimum(iv, R)
            elif op == "and":
                R = iv & R
            elif op == "or":
                R = iv | R
    if flt and op in ("add", "mul"):
        return R                                                  # Float64 reference, not rounded
    return R.astype(out_dtype)


def scan_tolerance(A, dims, T, init=None):
    """4*log2(n)*eps(T) * running sum of |x| (+|init|): the same bound as sum_tolerance per prefix."""
    A = np.asarray(A).astype(np.float64)
    ax = dims - 1
    n = A.shape[ax]
    mag = np.add.accumulate(np.abs(A), axis=ax)
    if init is not None:
        mag = mag + abs(float(init))
    return (4.0 * max(1.0, math.log2(max(n, 2))) * eps(T)) * mag + np.finfo(np.float64).tiny


# ----------------------------------------------------------------------------- sort
def total_order_key(a):
    """Unsigned integer whose order is Julia's `isless` on the element type:
    -Inf < ... < -0.0 < +0.0 < ... < +Inf < NaN, every NaN equal and last
    (Base.isless(::Float, ::Float)); two's-complement order for signed ints; false < true."""
    a = np.asarray(a)
    dt = a.dtype
    if dt == np.bool_:
        return a.astype(np.uint8)
    if dt.kind == "u":
        return a
    if dt.kind == "i":
        u = np.dtype(f"u{dt.itemsize}")
        return a.view(u) ^ u.type(1 << (8 * dt.itemsize - 1))
    # floats (incl. bfloat16 / float16)
    u = np.dtype(f"u{dt.itemsize}")
    bits = a.view(u)
    sign = u.type(1 << (8 * dt.itemsize - 1))
    allones = u.type(~u.type(0))
    key = np.where(bits & sign, ~bits, bits | sign)
    nan = np.isnan(a.astype(np.float64)) if dt.itemsize < 4 else np.isnan(a)
    return np.where(nan, allones, key).astype(u)


def sortperm(a, rev=False):
    """Base.sortperm(a; rev): stable, 1-based Int64; rev=true reverses the key order but
    keeps ascending index among equal keys (sorting.jl:582-592, Base's stable rev)."""
    key = total_order_key(np.asarray(a).reshape(-1))
    if rev:
        key = ~key if key.dtype != np.uint8 or np.asarray(a).dtype != np.bool_ else (1 - key).astype(np.uint8)
    return np.argsort(key, kind="stable").astype(np.int64) + 1


def sort(a, rev=False):
    """Base.sort(a; rev) under isless (values; NaNs last, -0.0 before +0.0)."""
    a = np.asarray(a).reshape(-1)
    return a[sortperm(a, rev=rev) - 1]


def sort_dims(A, dims, rev=False):
    A = np.asarray(A)
    ax = dims - 1
    M = np.moveaxis(A, ax, -1)
    out = np.empty_like(M)
    for idx in np.ndindex(M.shape[:-1]):
        out[idx] = sort(M[idx], rev=rev)
    return np.moveaxis(out, -1, ax)


def sortperm_dims(A, dims, rev=False):
    """sortperm(A; dims): LINEAR 1-based indices into A (sorting.jl:1067-1070)."""
    A = np.asarray(A)
    ax = dims - 1
    lin = (np.arange(A.size, dtype=np.int64) + 1).reshape(A.shape, order="F")
    M = np.moveaxis(A, ax, -1)
    L = np.moveaxis(lin, ax, -1)
    out = np.empty(M.shape, dtype=np.int64)
    for idx in np.ndindex(M.shape[:-1]):
        p = sortperm(M[idx], rev=rev) - 1
        out[idx] = L[idx][p]
    return np.moveaxis(out, -1, ax)


def is_valid_sort(orig, result, rev=False):
    """check_equivalence of the reference's tests (test/core/sorting.jl:169-171):
    same multiset and issorted under isless."""
    ko = np.sort(total_order_key(np.asarray(orig).reshape(-1)))
    kr = total_order_key(np.asarray(result).reshape(-1))
    srt = np.all(kr[:-1] >= kr[1:]) if rev else np.all(kr[:-1] <= kr[1:])
    return bool(srt) and np.array_equal(ko, np.sort(kr))


def findall(mask):
    """findall(::Array{Bool}) -> 1-based linear indices (CUDACore/src/indexing.jl:26-66)."""
    return np.flatnonzero(np.asarray(mask).reshape(-1, order="F")).astype(np.int64) + 1


def findminmax(A, dims=None, is_max=True):
    """Base.findmax / findmin: first extremal element, NaN is the maximum (findmax) or the minimum
    (findmin), -0.0 < +0.0.  Returns (values, 1-based linear indices); dims=None -> over everything."""
    A = np.asarray(A)
    key = total_order_key(A)
    kf = key.astype(np.float64) if key.dtype.itemsize < 8 else key
    if is_float(A.dtype) and not is_max:
        nan = np.isnan(A.astype(np.float64))
        key = np.where(nan, np.zeros_like(key), key)
    lin = (np.arange(A.size, dtype=np.int64) + 1).reshape(A.shape, order="F")
    if dims is None:
        flat_k, flat_l = key.reshape(-1, order="F"), lin.reshape(-1, order="F")
        j = int(np.argmax(flat_k) if is_max else np.argmin(flat_k))     # first occurrence
        return A.reshape(-1, order="F")[j], int(flat_l[j])
    ax = tuple(d - 1 for d in ((dims,) if isinstance(dims, int) else dims))
    assert len(ax) == 1, "oracle: one reduced dim"
    a = ax[0]
    j = np.argmax(key, axis=a) if is_max else np.argmin(key, axis=a)
    jj = np.expand_dims(j, a)
    return np.take_along_axis(A, jj, axis=a), np.take_along_axis(lin, jj, axis=a)
