// argreduce.cu — findmax / findmin / argmax / argmin: (value, index) reductions
// with Julia's rules: the FIRST extremal element wins, NaN is the maximum for findmax and the
// minimum for findmin, -0.0 < +0.0 (Base.findmax / findmin; the reference reaches them through
// GPUArrays' findminmax -> mapreducedim! with a tuple-valued op, tests test/core/array.jl:767-769,
// 782-784).  Values are compared through the same order-preserving unsigned keys as the sort
// (sort.cuh key_encode), so no float special cases remain in the kernels; the winning value is
// re-read from the input at the winning index.
//
// (pre, red, post) view as everywhere.  Two shapes: pre == 1 -> S CTAs per contiguous segment;
// pre > 1 -> lanes along pre, TY warps x S CTAs split red.  Multi-CTA outputs finish in the same
// launch through a ticket, like the plain reductions.
#include "sort.cuh"

namespace b200 {

struct ArgAcc { unsigned long long key; long long idx; };   // idx = 0-based position along red

template <bool IS_MAX> __device__ __forceinline__ ArgAcc arg_identity() {
    return {IS_MAX ? 0ull : ~0ull, 0x7fffffffffffffffLL};
}
template <bool IS_MAX> __device__ __forceinline__ ArgAcc arg_better(ArgAcc a, ArgAcc b) {
    const bool a_wins = IS_MAX ? (a.key > b.key || (a.key == b.key && a.idx < b.idx))
                               : (a.key < b.key || (a.key == b.key && a.idx < b.idx));
    return a_wins ? a : b;
}
__device__ __forceinline__ ArgAcc shfl_xor_arg(ArgAcc v, int m) {
    ArgAcc r;
    r.key = __shfl_xor_sync(0xffffffffu, v.key, m);
    r.idx = __shfl_xor_sync(0xffffffffu, v.idx, m);
    return r;
}

struct ArgParams {
    const void *in;
    void *out_vals;
    long long *out_idx;
    ArgAcc *partials;
    unsigned int *tickets;
    int64_t pre, red, post;
    int32_t S, TY, BX;
    SortXform xf;
};

// Ascending order-preserving key; every NaN encodes to the all-ones key (nan_or = all-ones), which no
// other float produces.  findmax: NaN therefore wins.  findmin (smallest key wins): NaN is remapped to
// key 0, which no non-NaN float produces either (-Inf encodes to 0x007f...ff).
template <class K, bool IS_MAX>
__device__ __forceinline__ unsigned long long arg_key(K raw, const SortXform &xf) {
    K k = key_encode<K>(raw, xf);
    if (!IS_MAX && xf.mag_mask != 0 && k == (K)xf.base_dec) k = 0;
    return (unsigned long long)k;
}

// pre == 1: blockIdx.x = s + S*seg
template <class K, bool IS_MAX>
__global__ void __launch_bounds__(256) argreduce_contig_kernel(ArgParams p) {
    __shared__ ArgAcc s_warp[8];
    __shared__ bool s_last;
    const int64_t seg = blockIdx.x / p.S;
    const int s = (int)(blockIdx.x - seg * p.S);
    const K *in = static_cast<const K *>(p.in) + seg * p.red;
    ArgAcc acc = arg_identity<IS_MAX>();
    // A thread visits its elements in increasing position, so "strictly better" keeps the first extremal one: the hot
    // loop compares keys of the element width and moves the position only on an improvement (round 1 compared 64-bit
    // (key, position) pairs per element behind scalar loads: 1.9 TB/s at 2^28 Float32, ALU-bound).  16-byte loads for
    // the aligned interior, scalar head and tail.
    {
        constexpr int VN = 16 / (int)sizeof(K);
        K bestk = IS_MAX ? K(0) : (K)~K(0);
        int64_t besti = 0x7fffffffffffffffLL;
        auto see = [&](K raw, int64_t r) {
            const K k = (K)arg_key<K, IS_MAX>(raw, p.xf);
            const bool better = IS_MAX ? (k > bestk) : (k < bestk);
            // the identity key itself (all-zero for max, all-ones for min) must still be recorded once
            if (better || besti == 0x7fffffffffffffffLL) { bestk = k; besti = r; }
        };
        int64_t head = (int64_t)(((16 - (reinterpret_cast<uintptr_t>(in) & 15)) & 15) / sizeof(K));
        if (head > p.red) head = p.red;
        const int64_t nvec = (p.red - head) / VN;
        const int64_t tail0 = head + nvec * VN;
        const int64_t t0 = (int64_t)s * 256 + threadIdx.x, stride = (int64_t)p.S * 256;
        if (t0 < head) see(ldg_stream(in + t0), t0);                       // head < 16 elements: CTA 0's first threads
        for (int64_t v = t0; v < nvec; v += stride) {
            const Vec16 raw = ldg_stream16(reinterpret_cast<const char *>(in + head) + v * 16);
            K e[VN];
            memcpy(e, &raw, 16);
#pragma unroll
            for (int q = 0; q < VN; ++q) see(e[q], head + v * VN + q);
        }
        if (tail0 + t0 < p.red && t0 < VN) see(ldg_stream(in + tail0 + t0), tail0 + t0);
        if (besti != 0x7fffffffffffffffLL) acc = ArgAcc{(unsigned long long)bestk, besti};
    }
    auto block_best = [&](ArgAcc v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = arg_better<IS_MAX>(v, shfl_xor_arg(v, o));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
        __syncthreads();
        ArgAcc w = s_warp[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) w = arg_better<IS_MAX>(w, s_warp[k]);
        return w;
    };
    acc = block_best(acc);
    auto finish = [&](ArgAcc best) {
        static_cast<K *>(p.out_vals)[seg] = in[best.idx];
        p.out_idx[seg] = 1 + seg * p.red + best.idx;           // 1-based linear index (pre == 1)
    };
    if (p.S == 1) { if (threadIdx.x == 0) finish(acc); return; }
    ArgAcc *partials = p.partials + seg * p.S;
    if (threadIdx.x == 0) {
        partials[s] = acc;
        __threadfence();
        s_last = atomicAdd(p.tickets + seg, 1u) == (unsigned int)p.S - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    ArgAcc f = arg_identity<IS_MAX>();
    for (int k = threadIdx.x; k < p.S; k += 256) {
        ArgAcc q;
        q.key = __ldcg(&partials[k].key);
        q.idx = __ldcg(&partials[k].idx);
        f = arg_better<IS_MAX>(f, q);
    }
    f = block_best(f);
    if (threadIdx.x == 0) { finish(f); p.tickets[seg] = 0; }
}

// pre > 1: thread (x, y): output o = xtile*BX + x, rows r = s*TY + y + k*S*TY
template <class K, bool IS_MAX>
__global__ void __launch_bounds__(256) argreduce_strided_kernel(ArgParams p) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    ArgAcc *s_acc = reinterpret_cast<ArgAcc *>(s_raw);   // [TY][BX]
    __shared__ bool s_last;
    const int BX = p.BX, TY = p.TY, S = p.S;
    const int x = threadIdx.x % BX, y = threadIdx.x / BX;
    const int64_t xtile = blockIdx.x / S;
    const int s = (int)(blockIdx.x - xtile * S);
    const int64_t outputs = p.pre * p.post;
    const int64_t o = xtile * BX + x;
    const bool active = o < outputs;
    const int64_t j = active ? o / p.pre : 0, i = active ? o - j * p.pre : 0;
    const K *base = static_cast<const K *>(p.in) + i + p.pre * p.red * j;
    ArgAcc acc = arg_identity<IS_MAX>();
    if (active) {
        const int64_t rstride = (int64_t)S * TY;
        for (int64_t r = (int64_t)s * TY + y; r < p.red; r += rstride)
            acc = arg_better<IS_MAX>(acc, ArgAcc{arg_key<K, IS_MAX>(ldg_stream(base + p.pre * r), p.xf), r});
    }
    if (TY > 1) {
        s_acc[y * BX + x] = acc;
        __syncthreads();
        if (y == 0) for (int yy = 1; yy < TY; ++yy) acc = arg_better<IS_MAX>(acc, s_acc[yy * BX + x]);
    }
    auto finish = [&](ArgAcc best) {
        static_cast<K *>(p.out_vals)[o] = base[p.pre * best.idx];
        p.out_idx[o] = 1 + i + p.pre * (best.idx + p.red * j);
    };
    if (S == 1) { if (active && y == 0) finish(acc); return; }
    ArgAcc *partials = p.partials + (xtile * S) * (int64_t)BX;
    if (y == 0) partials[(int64_t)s * BX + x] = acc;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(p.tickets + xtile, 1u) == (unsigned int)S - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (active && y == 0) {
        ArgAcc f = arg_identity<IS_MAX>();
        for (int ss = 0; ss < S; ++ss) {
            ArgAcc q;
            q.key = __ldcg(&partials[(int64_t)ss * BX + x].key);
            q.idx = __ldcg(&partials[(int64_t)ss * BX + x].idx);
            f = arg_better<IS_MAX>(f, q);
        }
        finish(f);
    }
    if (threadIdx.x == 0) p.tickets[xtile] = 0;
}

struct ArgPlan { int S, TY, BX; int64_t grid, ntickets, npartials; size_t smem; };

static ArgPlan plan_arg(int64_t pre, int64_t red, int64_t post) {
    const int sms = devinfo().sm_count > 0 ? devinfo().sm_count : 148;
    ArgPlan pl{};
    pl.S = 1; pl.TY = 1; pl.BX = 256;
    const int64_t target = 8LL * sms;
    if (pre == 1) {
        int64_t S = cdiv(target, post);
        const int64_t maxS = red / 4096;
        if (S > maxS) S = maxS;
        if (S < 1) S = 1;
        pl.S = (int)S;
        pl.grid = post * S;
        if (S > 1) { pl.ntickets = post; pl.npartials = post * S; }
        return pl;
    }
    const int64_t outputs = pre * post;
    if (outputs < 1024LL * sms) {
        const int bx = outputs >= 4096 ? 64 : 32;
        int ty = 256 / bx;
        while (ty > 1 && red / ty < 16) ty >>= 1;
        const int64_t xtiles = cdiv(outputs, bx);
        int64_t S = cdiv((13LL * sms) / 3, xtiles);
        const int64_t maxS = red / ((int64_t)ty * 16);
        if (S > maxS) S = maxS;
        if (S < 1) S = 1;
        pl.BX = bx; pl.TY = ty; pl.S = (int)S;
    }
    const int64_t xtiles = cdiv(outputs, pl.BX);
    pl.grid = xtiles * pl.S;
    if (pl.S > 1) { pl.ntickets = xtiles; pl.npartials = xtiles * pl.S * pl.BX; }
    pl.smem = pl.TY > 1 ? (size_t)pl.TY * pl.BX * sizeof(ArgAcc) : 0;
    return pl;
}

template <class K>
static int32_t launch_arg(void *ws, size_t wsb, const void *in, void *out_vals, void *out_idx, bool is_max,
                          int64_t pre, int64_t red, int64_t post, const SortXform &xf, cudaStream_t st, size_t *query) {
    const ArgPlan pl = plan_arg(pre, red, post);
    Carver c(ws);
    unsigned int *tickets = c.take<unsigned int>(pl.ntickets);
    ArgAcc *partials = c.take<ArgAcc>(pl.npartials);
    if (query) { *query = c.bytes(); return B200_OK; }
    if (c.bytes() > wsb) return B200_WORKSPACE_TOO_SMALL;
    if (c.bytes() > 0 && !ws) return B200_INVALID_VALUE;
    if (pl.grid > 2147483647LL) return B200_INVALID_VALUE;
    ArgParams p{};
    p.in = in; p.out_vals = out_vals; p.out_idx = static_cast<long long *>(out_idx);
    p.partials = partials; p.tickets = tickets; p.pre = pre; p.red = red; p.post = post;
    p.S = pl.S; p.TY = pl.TY; p.BX = pl.BX; p.xf = xf;
    if (pl.ntickets > 0) B200_CUDA_TRY(cudaMemsetAsync(tickets, 0, pl.ntickets * sizeof(unsigned int), st));
    const unsigned int grid = (unsigned int)pl.grid;
    if (pre == 1) {
        if (is_max) argreduce_contig_kernel<K, true><<<grid, 256, 0, st>>>(p);
        else argreduce_contig_kernel<K, false><<<grid, 256, 0, st>>>(p);
    } else {
        const unsigned int threads = (unsigned int)(pl.BX * pl.TY);
        if (is_max) argreduce_strided_kernel<K, true><<<grid, threads, pl.smem, st>>>(p);
        else argreduce_strided_kernel<K, false><<<grid, threads, pl.smem, st>>>(p);
    }
    B200_CHECK_LAUNCH();
    return B200_OK;
}

// order-preserving key transform with NaN as the maximum (findmax) or the minimum (findmin)
static bool arg_xform(int dtype, bool is_max, SortXform &x, int &kb) {
    x = SortXform{};
    kb = dtype_size(dtype);
    if (kb == 0) return false;
    const uint64_t all = kb == 8 ? ~0ull : ((1ull << (8 * kb)) - 1);
    const uint64_t sign = 1ull << (8 * kb - 1);
    x.sign = sign; x.inf_bits = all;
    bool is_float = false;
    switch (dtype) {
        case B200_F16: is_float = true; x.inf_bits = 0x7C00ull; break;
        case B200_BF16: is_float = true; x.inf_bits = 0x7F80ull; break;
        case B200_F32: is_float = true; x.inf_bits = 0x7F800000ull; break;
        case B200_F64: is_float = true; x.inf_bits = 0x7FF0000000000000ull; break;
        case B200_I32: case B200_I64: x.base_enc = sign; x.base_dec = sign; break;
        case B200_U32: case B200_U64: case B200_BOOL: break;
        default: return false;
    }
    if (is_float) {
        x.base_enc = sign; x.base_dec = all; x.fmn = all & ~sign; x.mag_mask = all & ~sign;
        x.nan_or = all;   // every NaN -> the all-ones key (see arg_key)
    }
    (void)is_max;
    return true;
}

}  // namespace b200

extern "C" {

int32_t b200_findminmax_workspace_bytes(int32_t dtype, int64_t pre, int64_t red, int64_t post, size_t *bytes) {
    using namespace b200;
    if (!bytes || pre < 0 || red < 0 || post < 0) return B200_INVALID_VALUE;
    *bytes = 0;
    SortXform xf; int kb;
    if (!arg_xform(dtype, true, xf, kb)) return B200_UNSUPPORTED;
    if (pre == 0 || red == 0 || post == 0) return B200_OK;
    return launch_arg<uint32_t>(nullptr, 0, nullptr, nullptr, nullptr, true, pre, red, post, xf, nullptr, bytes);
}

int32_t b200_findminmax(void *workspace, size_t workspace_bytes, const void *in, void *out_vals, void *out_idx,
                        int32_t dtype, int32_t is_max, int64_t pre, int64_t red, int64_t post, b200_stream_t stream) {
    using namespace b200;
    if (pre < 0 || red < 0 || post < 0) return B200_INVALID_VALUE;
    SortXform xf; int kb;
    if (!arg_xform(dtype, is_max != 0, xf, kb)) return B200_UNSUPPORTED;
    if (pre == 0 || red == 0 || post == 0) return B200_OK;
    if (!in || !out_vals || !out_idx) return B200_INVALID_VALUE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool mx = is_max != 0;
    switch (kb) {
        case 1: return launch_arg<uint8_t>(workspace, workspace_bytes, in, out_vals, out_idx, mx, pre, red, post, xf, st, nullptr);
        case 2: return launch_arg<uint16_t>(workspace, workspace_bytes, in, out_vals, out_idx, mx, pre, red, post, xf, st, nullptr);
        case 4: return launch_arg<uint32_t>(workspace, workspace_bytes, in, out_vals, out_idx, mx, pre, red, post, xf, st, nullptr);
        case 8: return launch_arg<uint64_t>(workspace, workspace_bytes, in, out_vals, out_idx, mx, pre, red, post, xf, st, nullptr);
        default: return B200_UNSUPPORTED;
    }
}

}  // extern "C"
