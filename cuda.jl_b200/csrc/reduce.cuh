// reduce.cuh — single-pass mapreduce kernels for sm_100a.
//
// Replaces partial_mapreduce_grid / serial_mapreduce_kernel and the recursive
// second launch of CUDACore/src/mapreduce.jl:90-157,259-298.
//
// A dense column-major array is viewed as (pre, red, post); element (i, r, j) is
// at i + pre*(r + red*j).  Two kernel families:
//
//   contig  (pre == 1): every output is the reduction of `red` CONTIGUOUS
//           elements.  Either a CTA (or S CTAs, grid-strided) per segment, or G
//           lanes of a warp per segment for short segments.  Streaming 16/32-byte
//           LDG.E.128/.256 loads, UNROLL independent loads in flight per thread.
//   strided (pre > 1): lanes run along `pre` (coalesced), each thread marches
//           down `red` for its column(s); TY warps and S CTAs split `red` when
//           there are too few columns to fill 148 SMs.
//
// Multi-CTA segments finish in the SAME launch: each CTA publishes its partial,
// takes a ticket, and the last arriver folds the partials in a fixed order
// (deterministic for a given shape).  Tickets live in the caller's workspace and
// are zeroed by a cudaMemsetAsync node on the stream (no second kernel).
#pragma once
#include "ops.cuh"

namespace b200 {

constexpr int RC_THREADS = 256;   // contig kernels
constexpr int RC_UNROLL = 4;      // independent vector loads in flight per thread
#ifndef B200_REDUCE_VB
#define B200_REDUCE_VB 32         // bytes per streaming load: 16 (LDG.E.128) or 32 (LDG.E.256)
#endif
constexpr int RS_THREADS = 256;   // strided kernels
#ifndef B200_RS_UNROLL
#define B200_RS_UNROLL 8
#endif
constexpr int RS_UNROLL = B200_RS_UNROLL;

template <int VB> struct VecOf;
template <> struct VecOf<16> { using type = Vec16; __device__ static Vec16 load(const void *p) { return ldg_stream16(p); } };
template <> struct VecOf<32> { using type = Vec32; __device__ static Vec32 load(const void *p) { return ldg_stream32(p); } };

struct ReduceParams {
    const void *in;
    void *out;
    void *partials;       // acc_t[nseg_tiles * S * tile_width]
    unsigned int *tickets;
    int64_t pre, red, post;
    int32_t S;            // CTAs cooperating on one output tile
    int32_t G;            // contig group mode: lanes per segment (1..32), 0 = block mode
    int32_t TY;           // strided: warps-rows per CTA along red
    int32_t BX;           // strided: threads along pre per CTA
    int32_t init_mode;
};

// L2-coherent load of a partial published by another CTA.
template <class T> __device__ __forceinline__ T ld_cg_any(const T *p) {
    T r;
    if constexpr (sizeof(T) == 1) {
        unsigned char b = __ldcg(reinterpret_cast<const unsigned char *>(p)); memcpy(&r, &b, 1);
    } else if constexpr (sizeof(T) == 2) {
        unsigned short b = __ldcg(reinterpret_cast<const unsigned short *>(p)); memcpy(&r, &b, 2);
    } else if constexpr (sizeof(T) % 8 == 0) {
        unsigned long long w[sizeof(T) / 8];
#pragma unroll
        for (int k = 0; k < (int)(sizeof(T) / 8); ++k) w[k] = __ldcg(reinterpret_cast<const unsigned long long *>(p) + k);
        memcpy(&r, w, sizeof(T));
    } else {
        unsigned int w[sizeof(T) / 4];
#pragma unroll
        for (int k = 0; k < (int)(sizeof(T) / 4); ++k) w[k] = __ldcg(reinterpret_cast<const unsigned int *>(p) + k);
        memcpy(&r, w, sizeof(T));
    }
    return r;
}

// ------------------------------------------------------------------ stores
template <class Tout, class Op>
__device__ __forceinline__ void store_result(Tout *out, int64_t idx, typename Op::acc_t acc, int init_mode) {
    using A = typename Op::val_t;
    if constexpr (sizeof(typename Op::acc_t) == sizeof(A)) {
        if (init_mode == B200_INIT_FROM_OUT) acc = Op::apply(Op::lift(cvt<A>(out[idx])), acc);
        out[idx] = cvt<Tout>(acc);
    } else {  // extrema: (min, max) pair
        if (init_mode == B200_INIT_FROM_OUT) {
            typename Op::acc_t prev = {cvt<A>(out[2 * idx]), cvt<A>(out[2 * idx + 1])};
            acc = Op::apply(prev, acc);
        }
        out[2 * idx] = cvt<Tout>(acc.mn);
        out[2 * idx + 1] = cvt<Tout>(acc.mx);
    }
}

// ------------------------------------------------------------------ block reduce
template <class Op, int THREADS>
__device__ __forceinline__ typename Op::acc_t block_reduce(typename Op::acc_t v, typename Op::acc_t *smem /* THREADS/32 */) {
    using acc_t = typename Op::acc_t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Op::apply(v, shfl_xor_any(v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // smem may still be read from a previous round
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        acc_t w = lane < THREADS / 32 ? smem[lane] : Op::identity();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w = Op::apply(w, shfl_xor_any(w, o));
        v = w;
    }
    return v;  // valid in warp 0
}

// ------------------------------------------------------------------ contig: per-thread streaming
// Workers w in [0, nw) cooperatively cover seg[0, red): vector v is handled by
// worker v mod nw.  Returns this worker's partial.
template <class Tin, class Op, class Map, int VB>
__device__ __forceinline__ typename Op::acc_t
contig_thread_partial(const Tin *__restrict__ seg, int64_t red, int64_t w, int64_t nw) {
    using A = typename Op::val_t;
    using acc_t = typename Op::acc_t;
    using V = typename VecOf<VB>::type;
    constexpr int VN = VB / (int)sizeof(Tin);
    constexpr int NACC = 4;
    acc_t acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = Op::identity();

    const uintptr_t addr = reinterpret_cast<uintptr_t>(seg);
    int64_t head = (int64_t)(((VB - (addr % VB)) % VB) / sizeof(Tin));
    if (head > red) head = red;
    const int64_t nvec = (red - head) / VN;
    const char *vbase = reinterpret_cast<const char *>(seg + head);

    // 1- and 2-byte integers summed into a >= 32-bit integer: IDP.4A / IDP.2A add 4 (2) elements per instruction,
    // one wide add per vector (count(A) on Bool would otherwise spend ~100 ALU instructions per 32-byte load)
    constexpr bool kDotSum = std::is_integral<Tin>::value && sizeof(Tin) <= 2 && std::is_integral<acc_t>::value &&
                             sizeof(acc_t) >= 4 && std::is_same<Op, OpAdd<acc_t>>::value &&
                             std::is_same<Map, MapIdentity>::value;
    auto consume = [&](const V &v) {
        if constexpr (kDotSum) {
            uint32_t wds[VB / 4];
            memcpy(wds, &v, VB);
            if constexpr (std::is_signed<Tin>::value) {
                int s0 = 0, s1 = 0;
#pragma unroll
                for (int k = 0; k < VB / 4; k += 2) {
                    if constexpr (sizeof(Tin) == 1) {
                        s0 = __dp4a((int)wds[k], 0x01010101, s0);
                        s1 = __dp4a((int)wds[k + 1], 0x01010101, s1);
                    } else {
                        s0 = __dp2a_lo((int)wds[k], 0x01010101, s0);
                        s1 = __dp2a_lo((int)wds[k + 1], 0x01010101, s1);
                    }
                }
                acc[0] += (acc_t)(s0 + s1);
            } else {
                unsigned s0 = 0, s1 = 0;
#pragma unroll
                for (int k = 0; k < VB / 4; k += 2) {
                    if constexpr (sizeof(Tin) == 1) {
                        s0 = __dp4a(wds[k], 0x01010101u, s0);
                        s1 = __dp4a(wds[k + 1], 0x01010101u, s1);
                    } else {
                        s0 = __dp2a_lo(wds[k], 0x01010101u, s0);
                        s1 = __dp2a_lo(wds[k + 1], 0x01010101u, s1);
                    }
                }
                acc[0] += (acc_t)(s0 + s1);
            }
        } else {
            Tin e[VN];
            memcpy(e, &v, VB);
#pragma unroll
            for (int k = 0; k < VN; ++k)
                acc[k % NACC] = Op::apply(acc[k % NACC], Op::lift(Map::template apply<A, Tin>(e[k])));
        }
    };

    int64_t v = w;
    for (; v + (int64_t)(RC_UNROLL - 1) * nw < nvec; v += (int64_t)RC_UNROLL * nw) {
        V buf[RC_UNROLL];
#pragma unroll
        for (int u = 0; u < RC_UNROLL; ++u) buf[u] = VecOf<VB>::load(vbase + (v + (int64_t)u * nw) * VB);
#pragma unroll
        for (int u = 0; u < RC_UNROLL; ++u) consume(buf[u]);
    }
    for (; v < nvec; v += nw) consume(VecOf<VB>::load(vbase + v * VB));

    // scalar leftovers: the unaligned head and the sub-vector tail
    const int64_t tail0 = head + nvec * VN;
    const int64_t nleft = head + (red - tail0);
    for (int64_t i = w; i < nleft; i += nw) {
        const int64_t idx = i < head ? i : tail0 + (i - head);
        acc[0] = Op::apply(acc[0], Op::lift(Map::template apply<A, Tin>(ldg_stream(seg + idx))));
    }
    acc[0] = Op::apply(acc[0], acc[1]);
    acc[2] = Op::apply(acc[2], acc[3]);
    return Op::apply(acc[0], acc[2]);
}

// Block mode: S CTAs per segment; blockIdx.x = s + S*seg.
template <class Tin, class Tout, class Op, class Map, int VB>
__global__ void __launch_bounds__(RC_THREADS) reduce_contig_block_kernel(ReduceParams p) {
    using acc_t = typename Op::acc_t;
    __shared__ acc_t s_warp[RC_THREADS / 32];
    __shared__ bool s_last;
    const int64_t seg = blockIdx.x / p.S;
    const int s = (int)(blockIdx.x - seg * p.S);
    const Tin *segp = static_cast<const Tin *>(p.in) + seg * p.red;
    Tout *out = static_cast<Tout *>(p.out);

    acc_t v = contig_thread_partial<Tin, Op, Map, VB>(segp, p.red, (int64_t)s * RC_THREADS + threadIdx.x,
                                                      (int64_t)p.S * RC_THREADS);
    v = block_reduce<Op, RC_THREADS>(v, s_warp);
    if (p.S == 1) {
        if (threadIdx.x == 0) store_result<Tout, Op>(out, seg, v, p.init_mode);
        return;
    }
    acc_t *partials = static_cast<acc_t *>(p.partials) + seg * p.S;
    if (threadIdx.x == 0) {
        partials[s] = v;
        __threadfence();
        const unsigned int t = atomicAdd(p.tickets + seg, 1u);
        s_last = (t == (unsigned int)p.S - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA of the segment: fold the S partials in a fixed order
    acc_t f = Op::identity();
    for (int k = threadIdx.x; k < p.S; k += RC_THREADS) f = Op::apply(f, ld_cg_any(partials + k));
    f = block_reduce<Op, RC_THREADS>(f, s_warp);
    if (threadIdx.x == 0) {
        store_result<Tout, Op>(out, seg, f, p.init_mode);
        p.tickets[seg] = 0;  // self-reset
    }
}

// Group mode: G lanes (power of two <= 32) per segment, no cross-CTA combine.
// G == 1 is the in-index-order sequential regime (mapreduce.jl:136-157).
template <class Tin, class Tout, class Op, class Map, int VB>
__global__ void __launch_bounds__(RC_THREADS) reduce_contig_group_kernel(ReduceParams p) {
    using A = typename Op::val_t;
    using acc_t = typename Op::acc_t;
    const int G = p.G;
    const int64_t gthread = (int64_t)blockIdx.x * RC_THREADS + threadIdx.x;
    const int64_t seg = gthread / G;
    const int w = (int)(gthread - seg * G);
    const bool active = seg < p.post;
    const Tin *segp = static_cast<const Tin *>(p.in) + (active ? seg : 0) * p.red;
    acc_t v = Op::identity();
    if (active) {
        if (G == 1) {
            // strictly sequential, index order, starting from op(neutral, neutral)
            v = Op::apply(Op::identity(), Op::identity());
            for (int64_t r = 0; r < p.red; ++r) v = Op::apply(v, Op::lift(Map::template apply<A, Tin>(segp[r])));
        } else {
            v = contig_thread_partial<Tin, Op, Map, VB>(segp, p.red, w, G);
        }
    }
    for (int o = G >> 1; o > 0; o >>= 1) v = Op::apply(v, shfl_xor_any(v, o));
    if (active && w == 0) store_result<Tout, Op>(static_cast<Tout *>(p.out), seg, v, p.init_mode);
}

// ------------------------------------------------------------------ strided (pre > 1)
// Thread (x, y): x indexes a group of VN consecutive outputs along pre, y splits
// red.  blockIdx.x = s + S*xtile.  Rows handled by (s, y): r = s*TY + y + k*S*TY.
template <class Tin, class Tout, class Op, class Map, int VN>
__global__ void __launch_bounds__(RS_THREADS) reduce_strided_kernel(ReduceParams p) {
    using A = typename Op::val_t;
    using acc_t = typename Op::acc_t;
    extern __shared__ __align__(16) unsigned char s_raw[];
    acc_t *s_acc = reinterpret_cast<acc_t *>(s_raw);  // [TY][BX*VN]
    __shared__ bool s_last;

    const int BX = p.BX, TY = p.TY, S = p.S;
    const int x = threadIdx.x % BX, y = threadIdx.x / BX;
    const int64_t xtile = blockIdx.x / S;
    const int s = (int)(blockIdx.x - xtile * S);
    const int64_t gpre = p.pre / VN;                 // groups per j (VN divides pre)
    const int64_t ngroups = gpre * p.post;
    const int64_t g = xtile * BX + x;
    const bool active = g < ngroups;
    const int64_t j = active ? g / gpre : 0;
    const int64_t i = active ? (g - j * gpre) * VN : 0;
    const Tin *base = static_cast<const Tin *>(p.in) + i + p.pre * p.red * j;
    const int64_t rstride = (int64_t)S * TY;

    acc_t acc[VN];
    // op(neutral, neutral), as the reference seeds every thread (mapreduce.jl:114,148)
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[k] = Op::apply(Op::identity(), Op::identity());

    if (active) {
        using LoadT = typename std::conditional<(VN * sizeof(Tin) == 16), Vec16,
                      typename std::conditional<(VN * sizeof(Tin) == 8), uint64_t,
                      typename std::conditional<(VN * sizeof(Tin) == 4), uint32_t,
                      typename std::conditional<(VN * sizeof(Tin) == 2), uint16_t, uint8_t>::type>::type>::type>::type;
        auto loadrow = [&](int64_t r) -> LoadT {
            const Tin *ptr = base + p.pre * r;
            if constexpr (VN * sizeof(Tin) == 16) return ldg_stream16(ptr);
            else return ldg_stream(reinterpret_cast<const LoadT *>(ptr));
        };
        auto consume = [&](const LoadT &v) {
            Tin e[VN];
            memcpy(e, &v, sizeof(LoadT));
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[k] = Op::apply(acc[k], Op::lift(Map::template apply<A, Tin>(e[k])));
        };
        int64_t r = (int64_t)s * TY + y;
        for (; r + (RS_UNROLL - 1) * rstride < p.red; r += RS_UNROLL * rstride) {
            LoadT buf[RS_UNROLL];
#pragma unroll
            for (int u = 0; u < RS_UNROLL; ++u) buf[u] = loadrow(r + u * rstride);
#pragma unroll
            for (int u = 0; u < RS_UNROLL; ++u) consume(buf[u]);
        }
        for (; r < p.red; r += rstride) consume(loadrow(r));
    }

    // combine the TY row-splits of this CTA in a fixed order
    if (TY > 1) {
#pragma unroll
        for (int k = 0; k < VN; ++k) s_acc[(y * BX + x) * VN + k] = acc[k];
        __syncthreads();
        if (y == 0) {
            for (int yy = 1; yy < TY; ++yy)
#pragma unroll
                for (int k = 0; k < VN; ++k) acc[k] = Op::apply(acc[k], s_acc[(yy * BX + x) * VN + k]);
        }
    }
    Tout *out = static_cast<Tout *>(p.out);
    const int64_t o0 = i + p.pre * j;  // output index of this thread's first column
    if (S == 1) {
        if (active && y == 0) {
#pragma unroll
            for (int k = 0; k < VN; ++k) store_result<Tout, Op>(out, o0 + k, acc[k], p.init_mode);
        }
        return;
    }
    // cross-CTA combine: partials[xtile][s][BX*VN]
    acc_t *partials = static_cast<acc_t *>(p.partials) + (xtile * S) * (int64_t)(BX * VN);
    if (y == 0) {
#pragma unroll
        for (int k = 0; k < VN; ++k) partials[(int64_t)s * BX * VN + x * VN + k] = acc[k];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(p.tickets + xtile, 1u);
        s_last = (t == (unsigned int)S - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA of the column tile: fold the S partials, every (x, y) thread taking the slices y, y+TY, ...
    // (a fixed assignment, so the result is deterministic), then the TY rows through shared memory
    acc_t f[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) f[k] = Op::identity();
    if (active) {
        for (int ss = y; ss < S; ss += TY)
#pragma unroll
            for (int k = 0; k < VN; ++k)
                f[k] = Op::apply(f[k], ld_cg_any(partials + (int64_t)ss * BX * VN + x * VN + k));
    }
    if (TY > 1) {
#pragma unroll
        for (int k = 0; k < VN; ++k) s_acc[(y * BX + x) * VN + k] = f[k];
        __syncthreads();
        if (y == 0) {
            for (int yy = 1; yy < TY; ++yy)
#pragma unroll
                for (int k = 0; k < VN; ++k) f[k] = Op::apply(f[k], s_acc[(yy * BX + x) * VN + k]);
        }
    }
    if (active && y == 0) {
#pragma unroll
        for (int k = 0; k < VN; ++k) store_result<Tout, Op>(out, o0 + k, f[k], p.init_mode);
    }
    if (threadIdx.x == 0) p.tickets[xtile] = 0;
}

// ------------------------------------------------------------------ host planning
struct ReducePlan {
    int family;          // 0 contig-block, 1 contig-group, 2 strided
    int S, G, TY, BX, VN;
    int64_t grid;
    int64_t ntickets;    // 0 when S == 1
    int64_t npartials;   // in acc_t units
    size_t smem;
};

// `in_addr`: address of `in` (only its alignment matters; pass 256 for "fully aligned").
static inline ReducePlan plan_reduce(int64_t pre, int64_t red, int64_t post, int in_size, int acc_size,
                                     uintptr_t in_addr) {
    const DevInfo &dev = devinfo();
    const int sms = dev.sm_count > 0 ? dev.sm_count : 148;
    ReducePlan pl{};
    pl.S = 1; pl.G = 0; pl.TY = 1; pl.BX = RS_THREADS; pl.VN = 1;
    const int64_t serial_threshold = 1024LL * sms;  // mapreduce.jl:162-166
    const int64_t target_ctas = 8LL * sms;          // 256-thread CTAs at full occupancy
    if (pre == 1) {
        const int vn = B200_REDUCE_VB / in_size;
        const int64_t nvec = red / vn;
        if (post >= serial_threshold) {
            pl.family = 1; pl.G = 1;                // one thread per output, index order
            pl.grid = cdiv(post, RC_THREADS);
        } else if (nvec < 2LL * RC_THREADS) {
            // short segments: a sub-warp group per segment, ~2 vectors per lane
            int G = 1;
            while (G < 32 && (int64_t)G * 2 < nvec) G <<= 1;
            pl.family = 1; pl.G = G;
            pl.grid = cdiv(post * G, RC_THREADS);
        } else {
            pl.family = 0;
            int64_t S = cdiv(target_ctas, post);
            const int64_t maxS = nvec / ((int64_t)RC_THREADS * RC_UNROLL);  // >= one unrolled tile per CTA
            if (S > maxS) S = maxS;
            if (S < 1) S = 1;
            pl.S = (int)S;
            pl.grid = post * S;
            if (S > 1) { pl.ntickets = post; pl.npartials = post * S; }
        }
        return pl;
    }
    // strided: lanes along pre
    pl.family = 2;
    const int vmax = 16 / in_size;
    int vn = (pre % vmax == 0 && in_addr % 16 == 0) ? vmax : 1;
    const int64_t outputs = pre * post;
    const int64_t target_threads = 512LL * sms;
    if (outputs >= serial_threshold) {
        // the reference's serial regime: one thread per output column, index order
        pl.VN = (outputs / vn >= target_threads) ? vn : 1;
        pl.BX = RS_THREADS; pl.TY = 1; pl.S = 1;
    } else {
        if (vn > 1 && outputs / vn < 32) vn = 1;
        pl.VN = vn;
        const int64_t groups = outputs / vn;
        // measured on B200 (16384^2, sum over dims=2): ~4.3 CTAs per SM of 256 threads is the sweet spot
        // (F32: BX=64,TY=4,S=10 -> 6.1 TB/s vs 5.6 TB/s at 8 CTAs/SM; BF16: BX=32,TY=8,S=10 -> 5.6 vs 4.9 TB/s)
        int bx = groups >= 4096 ? 64 : 32;
        // few columns (sum(A; dims=2) on 3 x 1_000_000): lanes along the columns would idle, so shrink the
        // column tile and give the freed lanes to the rows (a warp then reads 32/bx consecutive rows)
        while (bx > 1 && bx / 2 >= groups) bx >>= 1;
        const int64_t min_rows = 16;                // rows per (s, y) worker, at least
        int ty = RS_THREADS / bx;
        while (ty > 1 && red / ty < min_rows) ty >>= 1;
        const int64_t xtiles = cdiv(groups, bx);
        int64_t S = cdiv((13LL * sms) / 3, xtiles);
        const int64_t maxS = red / ((int64_t)ty * min_rows);
        if (S > maxS) S = maxS;
        if (S < 1) S = 1;
        pl.BX = bx; pl.TY = ty; pl.S = (int)S;
    }
    const int64_t groups = (pre / pl.VN) * post;
    const int64_t xtiles = cdiv(groups, pl.BX);
    pl.grid = xtiles * pl.S;
    if (pl.S > 1) { pl.ntickets = xtiles; pl.npartials = xtiles * pl.S * pl.BX * pl.VN; }
    pl.smem = pl.TY > 1 ? (size_t)pl.TY * pl.BX * pl.VN * acc_size : 0;
    return pl;
}

// Launch for one (Tin, Tout, Op, Map) instantiation.
template <class Tin, class Tout, class Op, class Map>
int32_t launch_reduce(void *ws, size_t ws_bytes, const void *in, void *out, int64_t pre, int64_t red, int64_t post,
                      int init_mode, cudaStream_t stream, size_t *query_bytes) {
    using acc_t = typename Op::acc_t;
    // workspace is planned for the worst alignment-independent case: plan with the real pointer
    // when launching, with a fully aligned one when querying (VN only shrinks => fewer partials? no:
    // smaller VN means more column groups; so query with the minimum VN instead).
    const bool query = query_bytes != nullptr;
    ReducePlan pl = plan_reduce(pre, red, post, (int)sizeof(Tin), (int)sizeof(acc_t),
                                query ? (uintptr_t)1 * sizeof(Tin) : reinterpret_cast<uintptr_t>(in));
    if (query) {
        // upper bound over possible alignments: take the max over aligned and unaligned plans
        ReducePlan pa = plan_reduce(pre, red, post, (int)sizeof(Tin), (int)sizeof(acc_t), 256);
        Carver c1(nullptr), c2(nullptr);
        c1.take<unsigned int>(pl.ntickets); c1.take<acc_t>(pl.npartials);
        c2.take<unsigned int>(pa.ntickets); c2.take<acc_t>(pa.npartials);
        *query_bytes = c1.bytes() > c2.bytes() ? c1.bytes() : c2.bytes();
        return B200_OK;
    }
    Carver c(ws);
    unsigned int *tickets = c.take<unsigned int>(pl.ntickets);
    acc_t *partials = c.take<acc_t>(pl.npartials);
    if (c.bytes() > ws_bytes) return B200_WORKSPACE_TOO_SMALL;
    if (c.bytes() > 0 && ws == nullptr) return B200_INVALID_VALUE;
    if (pl.grid > 2147483647LL) return B200_INVALID_VALUE;

    ReduceParams p{};
    p.in = in; p.out = out; p.partials = partials; p.tickets = tickets;
    p.pre = pre; p.red = red; p.post = post;
    p.S = pl.S; p.G = pl.G; p.TY = pl.TY; p.BX = pl.BX; p.init_mode = init_mode;
    if (pl.ntickets > 0) B200_CUDA_TRY(cudaMemsetAsync(tickets, 0, pl.ntickets * sizeof(unsigned int), stream));

    const dim3 grid((unsigned int)pl.grid);
    if (pl.family == 0) {
        reduce_contig_block_kernel<Tin, Tout, Op, Map, B200_REDUCE_VB><<<grid, RC_THREADS, 0, stream>>>(p);
    } else if (pl.family == 1) {
        reduce_contig_group_kernel<Tin, Tout, Op, Map, B200_REDUCE_VB><<<grid, RC_THREADS, 0, stream>>>(p);
    } else {
        constexpr int VMAX = 16 / (int)sizeof(Tin);
        const unsigned int threads = (unsigned int)(pl.BX * pl.TY);
        if (pl.VN == 1) reduce_strided_kernel<Tin, Tout, Op, Map, 1><<<grid, threads, pl.smem, stream>>>(p);
        else if (pl.VN == VMAX) reduce_strided_kernel<Tin, Tout, Op, Map, VMAX><<<grid, threads, pl.smem, stream>>>(p);
        else return B200_INVALID_VALUE;
    }
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // namespace b200
