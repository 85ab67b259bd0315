// scan.cuh — single-pass prefix scan (decoupled look-back) for sm_100a.
//
// Replaces partial_scan + aggregate_partial_scan + the recursive host glue of
// CUDACore/src/accumulate.jl:15-131,166-196 (>= 4 launches per level and ~4N
// traffic) with ONE kernel that reads every element once and writes it once.
//
// Layout: (pre, n, post) view of a dense column-major array, scanned along n.
//   contig  (pre == 1): `post` independent chains.  A tile is 256 threads x K
//           chunks; chunks are WARP-STRIPED (chunk c of a warp = row k, lane l,
//           c = k*32 + l) so every load/store instruction of a warp covers one
//           contiguous 512-byte span: fully coalesced with no shared-memory
//           transpose.  Per tile: chunk-serial scan, K warp scans, one block scan,
//           then Merrill-Garland decoupled look-back over the per-tile status words.
//   strided (pre > 1): one thread per column marching along n (coalesced across
//           columns).  [SURVEY §8 f2: look-back chains for strided scans are "next".]
//
// Kernels (host selection in launch_scan at the end of this file):
//   scan_chain2_kernel      default for contiguous scans: persistent CTAs, TMA ring, self-chained prefixes,
//                           control warp; 6.6 TB/s on cumsum 2^30 Float32
//   scan_pipe_kernel        decoupled look-back in a persistent TMA ring; used when its tile fits the segments
//                           better (several segments) or with B200_SCAN_CHAIN=0; scan_pipe_ws_kernel (look-back in
//                           its own warp, B200_SCAN_WS=1) is a measured experiment
//   scan_contig_kernel      LSU fallback for unaligned inputs
//   scan_warpseg_kernel     many short contiguous segments, one warp each
//   scan_strided_lb_kernel / scan_fewcols_kernel / scan_strided_kernel   scans along a non-leading dimension
//
// Tile status words live in the caller's workspace: {flag:32 | payload:32} packed in
// one 64-bit word per 32 bits of the accumulator, written/read with relaxed
// gpu-scope 64-bit accesses (single-copy atomic), so no fences are needed; a
// reader of a 64-bit accumulator retries until both halves carry the same flag.
// Tiles take their index from an atomic ticket, so a tile's predecessors are always
// resident or finished (forward progress without relying on block scheduling order).
#pragma once
#include "ops.cuh"
#include <cstdlib>
#include <cstdio>

namespace b200 {

constexpr int SC_THREADS = 256;
constexpr int SC_WARPS = SC_THREADS / 32;
constexpr uint32_t SC_INVALID = 0, SC_PARTIAL = 1, SC_INCLUSIVE = 2;
#ifndef B200_SC_LOOK
#define B200_SC_LOOK 4
#endif
constexpr int SC_LOOK = B200_SC_LOOK;  // predecessors inspected per lane per look-back round trip

struct ScanParams {
    const void *in;
    void *out;
    uint64_t *status;              // [post * tiles_per_seg * NW]
    unsigned long long *ticket;    // [1]
    int64_t pre, n, post;
    int64_t tiles_per_seg;
    int32_t exclusive;
    int32_t has_init;
    uint64_t init_bits;            // the Some(x) value, bit pattern of the accumulator type
    const void *carry_dev;         // optional DEVICE array of accumulator-typed values folded (in order) into the seed:
    int32_t carry_count;           // the totals of the shards before this one in a sharded scan (no host round trip)
    int32_t chain_at;              // scan_chain2_kernel: 0 = status prefetch before barrier (A), 1 = after the publication
};

template <class A> __device__ __forceinline__ A bits_to_acc(uint64_t bits);

// What every chain of the scan starts from: init (a host value) and / or the device-side carry values, else the
// operator's identity.  Evaluated where a segment's FIRST tile is seeded only.
template <class A, class Op> __device__ __forceinline__ A scan_seed(const ScanParams &p) {
    A s = p.has_init ? bits_to_acc<A>(p.init_bits) : Op::identity();
    if (p.carry_dev != nullptr) {
        const A *c = static_cast<const A *>(p.carry_dev);
        for (int i = 0; i < p.carry_count; ++i) s = Op::apply(s, c[i]);
    }
    return s;
}

__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

#ifdef B200_SCAN_STATS
// debug build only (tools/build_variant.sh stats -DB200_SCAN_STATS): where a persistent CTA's time goes
__device__ unsigned long long g_scan_stats[16];
#define SC_STAT_T(var) const long long var = clock64()
#define SC_STAT_ADD(i, v) do { if (threadIdx.x == 0) st_acc[i] += (unsigned long long)(v); } while (0)
#else
#define SC_STAT_T(var)
#define SC_STAT_ADD(i, v)
#endif

// Segment of global tile index g.  Persistent kernels evaluate this on their critical path (the control warp of
// scan_chain2_kernel once per ticket; measured: a 64-bit division there cost 7 % of the whole cumsum), so the common
// cases avoid the ~100-instruction 64-bit divide: one segment -> 0, indices below 2^32 -> 32-bit divide.
__device__ __forceinline__ int64_t seg_of_tile(int64_t g, const ScanParams &p) {
    if (p.post == 1) return 0;
    if (((uint64_t)g | (uint64_t)p.tiles_per_seg) <= 0xFFFFFFFFull) return (int64_t)((uint32_t)g / (uint32_t)p.tiles_per_seg);
    return g / p.tiles_per_seg;
}

template <class A> struct StatusWords { static constexpr int NW = sizeof(A) > 4 ? 2 : 1; };

template <class A> __device__ __forceinline__ void publish(uint64_t *slot, uint32_t flag, A value) {
    constexpr int NW = StatusWords<A>::NW;
    uint32_t w[NW] = {};
    memcpy(w, &value, sizeof(A));
#pragma unroll
    for (int k = 0; k < NW; ++k) st_relaxed_u64(slot + k, ((uint64_t)flag << 32) | w[k]);
}

// Spins until the slot is valid; returns its flag and value.
template <class A> __device__ __forceinline__ uint32_t wait_slot(const uint64_t *slot, A &value) {
    constexpr int NW = StatusWords<A>::NW;
    uint32_t w[NW];
    uint32_t flag;
    while (true) {
        uint64_t r[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) r[k] = ld_relaxed_u64(slot + k);
        flag = (uint32_t)(r[0] >> 32);
        bool ok = flag != SC_INVALID;
#pragma unroll
        for (int k = 0; k < NW; ++k) { ok = ok && ((uint32_t)(r[k] >> 32) == flag); w[k] = (uint32_t)r[k]; }
        if (ok) break;
    }
    memcpy(&value, w, sizeof(A));
    return flag;
}

// One attempt: returns SC_INVALID when the slot is empty or its halves disagree.
template <class A> __device__ __forceinline__ uint32_t peek_slot(const uint64_t *slot, A &value) {
    constexpr int NW = StatusWords<A>::NW;
    uint64_t r[NW];
    if constexpr (NW == 2) {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(r[0]), "=l"(r[1]) : "l"(slot) : "memory");
    } else {
        r[0] = ld_relaxed_u64(slot);
    }
    uint32_t w[NW];
    uint32_t flag = (uint32_t)(r[0] >> 32);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        if ((uint32_t)(r[k] >> 32) != flag) flag = SC_INVALID;
        w[k] = (uint32_t)r[k];
    }
    memcpy(&value, w, sizeof(A));
    return flag;
}

template <class A> __device__ __forceinline__ A bits_to_acc(uint64_t bits) {
    A a;
    memcpy(&a, &bits, sizeof(A));
    return a;
}

// Decoupled look-back, executed by one full warp.  Publishes this tile's aggregate,
// walks back over predecessor status words until an inclusive prefix is found,
// publishes this tile's inclusive prefix and returns its exclusive prefix.
// Wide window: every lane inspects SC_LOOK consecutive predecessors per round trip
// (32*SC_LOOK tiles per window).
template <class Op, int LOOK = SC_LOOK>
__device__ __forceinline__ typename Op::acc_t tile_lookback(const ScanParams &p, int64_t seg, int64_t t, int lane,
                                                            typename Op::acc_t tile_agg, bool publish_partial = true) {
    using A = typename Op::acc_t;
    constexpr int NW = StatusWords<A>::NW;
    uint64_t *slots = p.status + seg * p.tiles_per_seg * NW;
    A prefix;
    if (t == 0) {
        prefix = scan_seed<A, Op>(p);
        if (lane == 0 && p.tiles_per_seg > 1) publish<A>(slots, SC_INCLUSIVE, Op::apply(prefix, tile_agg));
        return prefix;
    }
    if (lane == 0 && publish_partial) publish<A>(slots + t * NW, SC_PARTIAL, tile_agg);
    prefix = Op::identity();
    // One window = LOOK rows of 32 consecutive predecessors; lane l of row w looks at the tile
    // at distance w*32 + l + 1.  Consecutive lanes read consecutive status words, so every
    // load instruction of the warp is 2 (4 for 64-bit accumulators) L2 line requests.
    int64_t look = t - 1;
    while (true) {
        A vals[LOOK];
        uint32_t flags[LOOK];
        while (true) {
            bool all_valid = true;
#pragma unroll
            for (int w = 0; w < LOOK; ++w) {
                const int64_t idx = look - (w * 32 + lane);
                flags[w] = SC_INCLUSIVE;          // virtual tiles before the first one: inclusive, identity
                vals[w] = Op::identity();
                if (idx >= 0) flags[w] = peek_slot<A>(slots + idx * NW, vals[w]);
                all_valid = all_valid && (flags[w] != SC_INVALID);
            }
            if (__all_sync(0xffffffffu, all_valid)) break;
        }
        A val = Op::identity();
        bool found = false;
#pragma unroll
        for (int w = 0; w < LOOK; ++w) {
            if (!found) {          // warp-uniform
                const uint32_t incl = __ballot_sync(0xffffffffu, flags[w] == SC_INCLUSIVE);
                const int nearest = __ffs(incl) - 1;      // closest predecessor of this row holding an inclusive prefix
                if (!incl || lane <= nearest) val = Op::apply(vals[w], val);
                found = incl != 0u;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val = Op::apply(val, shfl_xor_any(val, o));
        prefix = Op::apply(val, prefix);
        if (found) break;
        look -= 32 * LOOK;
    }
    if (lane == 0) publish<A>(slots + t * NW, SC_INCLUSIVE, Op::apply(prefix, tile_agg));
    return prefix;
}

// Loads EPC consecutive elements starting at element offset `off` of the tile;
// elements at or beyond `valid` are replaced by the operator's identity.
template <class Tin, class A, class Op, int EPC>
__device__ __forceinline__ void load_chunk(const Tin *tile, int64_t off, int64_t valid, A (&v)[EPC]) {
    if (off + EPC <= valid) {
        using LoadT = typename std::conditional<(EPC * sizeof(Tin) == 16), Vec16,
                      typename std::conditional<(EPC * sizeof(Tin) == 8), uint64_t,
                      typename std::conditional<(EPC * sizeof(Tin) == 4), uint32_t,
                      typename std::conditional<(EPC * sizeof(Tin) == 2), uint16_t, uint8_t>::type>::type>::type>::type;
        static_assert(sizeof(LoadT) == EPC * sizeof(Tin), "chunk size must be 1,2,4,8 or 16 bytes");
        LoadT raw;
        if constexpr (sizeof(LoadT) == 16) raw = ld_coh16(tile + off);
        else raw = ld_coh(reinterpret_cast<const LoadT *>(tile + off));
        Tin e[EPC];
        memcpy(e, &raw, sizeof(LoadT));
#pragma unroll
        for (int k = 0; k < EPC; ++k) v[k] = cvt<A>(e[k]);
    } else {
#pragma unroll
        for (int k = 0; k < EPC; ++k) v[k] = (off + k < valid) ? cvt<A>(ld_coh(tile + off + k)) : Op::identity();
    }
}

template <class Tout, class A, int EPC>
__device__ __forceinline__ void store_chunk(Tout *tile, int64_t off, int64_t valid, const A (&v)[EPC]) {
    if (off + EPC <= valid) {
        Tout e[EPC];
#pragma unroll
        for (int k = 0; k < EPC; ++k) e[k] = cvt<Tout>(v[k]);
        if constexpr (EPC * sizeof(Tout) == 16) {
            Vec16 raw;
            memcpy(&raw, e, 16);
            *reinterpret_cast<Vec16 *>(tile + off) = raw;
        } else if constexpr (EPC * sizeof(Tout) == 32) {
            Vec16 raw[2];
            memcpy(raw, e, 32);
            *reinterpret_cast<Vec16 *>(tile + off) = raw[0];
            *(reinterpret_cast<Vec16 *>(tile + off) + 1) = raw[1];
        } else if constexpr (EPC * sizeof(Tout) == 8) {
            uint64_t raw; memcpy(&raw, e, 8);
            *reinterpret_cast<uint64_t *>(tile + off) = raw;
        } else if constexpr (EPC * sizeof(Tout) == 4) {
            uint32_t raw; memcpy(&raw, e, 4);
            *reinterpret_cast<uint32_t *>(tile + off) = raw;
        } else if constexpr (EPC * sizeof(Tout) == 2) {
            uint16_t raw; memcpy(&raw, e, 2);
            *reinterpret_cast<uint16_t *>(tile + off) = raw;
        } else {
#pragma unroll
            for (int k = 0; k < EPC; ++k) tile[off + k] = e[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < EPC; ++k)
            if (off + k < valid) tile[off + k] = cvt<Tout>(v[k]);
    }
}

// One tile per CTA.  EPC = elements per chunk, K = chunk rows per warp.
template <class Tin, class Tout, class Op, int EPC, int K>
__global__ void __launch_bounds__(SC_THREADS) scan_contig_kernel(ScanParams p) {
    using A = typename Op::acc_t;
    constexpr int WARP_ELEMS = 32 * EPC * K;
    constexpr int TILE = SC_WARPS * WARP_ELEMS;
    __shared__ A s_warp[SC_WARPS];
    __shared__ A s_tile_prefix;
    __shared__ unsigned long long s_ticket;

    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1ull);
    __syncthreads();
    const int64_t g = (int64_t)s_ticket;
    const int64_t seg = g / p.tiles_per_seg;
    const int64_t t = g - seg * p.tiles_per_seg;
    const int64_t tile_off = seg * p.n + t * TILE;
    int64_t valid = p.n - t * TILE;
    if (valid > TILE) valid = TILE;
    const Tin *tin = static_cast<const Tin *>(p.in) + tile_off;
    Tout *tout = static_cast<Tout *>(p.out) + tile_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- load (warp-striped chunks) and scan inside each chunk
    A v[K][EPC];
#pragma unroll
    for (int k = 0; k < K; ++k)
        load_chunk<Tin, A, Op, EPC>(tin, (int64_t)warp * WARP_ELEMS + (k * 32 + lane) * EPC, valid, v[k]);
    A csum[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int e = 1; e < EPC; ++e) v[k][e] = Op::apply(v[k][e - 1], v[k][e]);
        csum[k] = v[k][EPC - 1];
    }
    // ---- K interleaved warp scans over the chunk sums
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            A up = shfl_up_any(csum[k], d);
            if (lane >= d) csum[k] = Op::apply(up, csum[k]);
        }
    }
    // chunk-exclusive prefix within the warp: earlier rows, then earlier lanes of this row
    A cpre[K];
    A run = Op::identity();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        A excl = shfl_up_any(csum[k], 1);
        A rtot = shfl_idx_any(csum[k], 31);
        cpre[k] = lane == 0 ? run : Op::apply(run, excl);
        run = Op::apply(run, rtot);
    }
    if (lane == 0) s_warp[warp] = run;  // warp aggregate
    __syncthreads();

    // ---- warp 0: scan of warp aggregates + decoupled look-back
    if (warp == 0) {
        A wagg = lane < SC_WARPS ? s_warp[lane] : Op::identity();
        A wincl = wagg;
#pragma unroll
        for (int d = 1; d < SC_WARPS; d <<= 1) {
            A up = shfl_up_any(wincl, d);
            if (lane >= d) wincl = Op::apply(up, wincl);
        }
        const A tile_agg = shfl_idx_any(wincl, SC_WARPS - 1);
        A wexcl = shfl_up_any(wincl, 1);
        if (lane == 0) wexcl = Op::identity();

        const A prefix = tile_lookback<Op>(p, seg, t, lane, tile_agg);
        if (lane == 0) s_tile_prefix = prefix;
        if (lane < SC_WARPS) s_warp[lane] = wexcl;
    }
    __syncthreads();
    const A base = Op::apply(s_tile_prefix, s_warp[warp]);

    // ---- apply prefixes and store
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const A pre = Op::apply(base, cpre[k]);
        A o[EPC];
        if (!p.exclusive) {
#pragma unroll
            for (int e = 0; e < EPC; ++e) o[e] = Op::apply(pre, v[k][e]);
        } else {
            o[0] = pre;
#pragma unroll
            for (int e = 1; e < EPC; ++e) o[e] = Op::apply(pre, v[k][e - 1]);
        }
        store_chunk<Tout, A, EPC>(tout, (int64_t)warp * WARP_ELEMS + (k * 32 + lane) * EPC, valid, o);
    }
}


// ------------------------------------------------------------------ TMA tile kernel
// Same tiling, but the tile travels global -> shared -> global through the TMA unit
// (cp.async.bulk, SASS UBLKCP): the bulk data never enters the SM's LSU / L1TEX
// queue, so the look-back's status-word loads see plain L2 latency instead of
// queueing behind ~100 KB of streaming loads per SM (measured: ~2 us per status
// round trip when the data moves through LDG, which capped the scan at 2.5 TB/s).
// Registers only hold the K chunk sums between the two passes over shared memory.
template <class T, int N> __device__ __forceinline__ void lds_chunk(const T *sm, T (&e)[N]) {
    if constexpr (N * sizeof(T) == 16) {
        Vec16 raw = *reinterpret_cast<const Vec16 *>(sm);
        memcpy(e, &raw, 16);
    } else if constexpr (N * sizeof(T) == 32) {
        Vec16 raw[2];
        raw[0] = *reinterpret_cast<const Vec16 *>(sm);
        raw[1] = *(reinterpret_cast<const Vec16 *>(sm) + 1);
        memcpy(e, raw, 32);
    } else if constexpr (N * sizeof(T) == 8) {
        uint64_t raw = *reinterpret_cast<const uint64_t *>(sm);
        memcpy(e, &raw, 8);
    } else if constexpr (N * sizeof(T) == 4) {
        uint32_t raw = *reinterpret_cast<const uint32_t *>(sm);
        memcpy(e, &raw, 4);
    } else if constexpr (N * sizeof(T) == 2) {
        uint16_t raw = *reinterpret_cast<const uint16_t *>(sm);
        memcpy(e, &raw, 2);
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) e[k] = sm[k];
    }
}
template <class T, int N> __device__ __forceinline__ void sts_chunk(T *sm, const T (&e)[N]) {
    if constexpr (N * sizeof(T) == 16) {
        Vec16 raw; memcpy(&raw, e, 16);
        *reinterpret_cast<Vec16 *>(sm) = raw;
    } else if constexpr (N * sizeof(T) == 32) {
        Vec16 raw[2]; memcpy(raw, e, 32);
        *reinterpret_cast<Vec16 *>(sm) = raw[0];
        *(reinterpret_cast<Vec16 *>(sm) + 1) = raw[1];
    } else if constexpr (N * sizeof(T) == 8) {
        uint64_t raw; memcpy(&raw, e, 8);
        *reinterpret_cast<uint64_t *>(sm) = raw;
    } else if constexpr (N * sizeof(T) == 4) {
        uint32_t raw; memcpy(&raw, e, 4);
        *reinterpret_cast<uint32_t *>(sm) = raw;
    } else if constexpr (N * sizeof(T) == 2) {
        uint16_t raw; memcpy(&raw, e, 2);
        *reinterpret_cast<uint16_t *>(sm) = raw;
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) sm[k] = e[k];
    }
}

// ------------------------------------------------------------------ persistent TMA pipeline
// The tile-per-CTA kernels above suffer a convoy: all ~600 resident tiles start together,
// their loads land together, and each tile waits for the slowest predecessor before any
// store is issued, so HBM idles during every look-back (measured: 7 us of look-back per
// 12 us tile lifetime).  Here each CTA is persistent and owns a ring of S shared-memory
// stages filled by cp.async.bulk loads that run S-1 tiles ahead of the tile being scanned.
// Two things keep the look-back chain short although a CTA holds S tickets at a time:
//   * phase A: as soon as a prefetched tile has LANDED (non-blocking mbarrier test) the CTA
//     reduces it and publishes its aggregate, so no other tile ever waits for this CTA to
//     "get to" a tile it merely holds;
//   * phase B (in ticket order): scan in shared memory, look-back, apply, store with ordinary
//     coalesced 16-byte STGs (the stage is free for the next load right away).
// Tiles are handed out by an atomic ticket, so a CTA only waits on tiles that some resident
// CTA already owns.
template <class Tin, class Tout, class Op, int EPC, int K, int S, int NT, int LOOK>
__global__ void __launch_bounds__(NT) scan_pipe_kernel(ScanParams p) {
    using A = typename Op::acc_t;
    constexpr int NWARPS = NT / 32;
    constexpr int NW = StatusWords<A>::NW;
    constexpr int WARP_ELEMS = 32 * EPC * K;
    constexpr int TILE = NWARPS * WARP_ELEMS;
    constexpr size_t STAGE_BYTES = (size_t)TILE * sizeof(Tin);
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) uint64_t s_full[S];
    __shared__ long long s_tile[S];        // ticket held by each stage
    __shared__ int s_use[S];               // how many times the stage has been armed (barrier parity)
    __shared__ A s_wexcl[2][NWARPS];       // warp-exclusive prefixes inside the tile: [tile parity][warp]
    __shared__ A s_wtot[NWARPS];
    __shared__ A s_tile_prefix;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t total = p.tiles_per_seg * p.post;
    const int64_t wbase = (int64_t)warp * WARP_ELEMS;
#ifdef B200_SCAN_STATS
    unsigned long long st_acc[16] = {};
#endif

    // thread 0: put ticket g into `stage` and start its load
    auto fill = [&](int stage, long long g) {
        s_tile[stage] = g;
        s_use[stage] += 1;
        if (g < total) {
            const int64_t seg = seg_of_tile(g, p), t = g - seg * p.tiles_per_seg;
            if (p.n - t * TILE >= TILE) {
                mbar_expect_tx(&s_full[stage], (uint32_t)STAGE_BYTES);
                tma_load_1d(s_dyn + stage * STAGE_BYTES, static_cast<const Tin *>(p.in) + seg * p.n + t * TILE,
                            (uint32_t)STAGE_BYTES, &s_full[stage]);
            } else {
                mbar_arrive(&s_full[stage]);   // ragged tail: loaded with ordinary loads in pass 1
            }
        }
    };

    // Pass 1 over the tile in `stage` (blocks until its data has landed): chunk sums, K warp
    // scans -> chunk-exclusive prefixes inside the warp (returned in cpre), warp totals to smem.
    // After the caller's __syncthreads, finish_pass1 (warp 0) turns the warp totals into
    // warp-exclusive prefixes and PUBLISHES the tile aggregate — one tile ahead of its turn.
    auto pass1 = [&](int stage, A (&cpre)[K]) {
        const long long g = s_tile[stage];
        const int64_t seg = seg_of_tile(g, p), t = g - seg * p.tiles_per_seg;
        int64_t valid = p.n - t * TILE;
        const bool full = valid >= TILE;
        if (valid > TILE) valid = TILE;
        Tin *s_in = reinterpret_cast<Tin *>(s_dyn + stage * STAGE_BYTES);
        SC_STAT_T(tw0);
        mbar_wait(&s_full[stage], (uint32_t)((s_use[stage] - 1) & 1));
        SC_STAT_T(tw1);
        SC_STAT_ADD(1, tw1 - tw0);          // waiting for the tile to land
        A csum[K];
        if (full) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                Tin e[EPC];
                lds_chunk<Tin, EPC>(s_in + wbase + (k * 32 + lane) * EPC, e);
                A acc = cvt<A>(e[0]);
#pragma unroll
                for (int q = 1; q < EPC; ++q) acc = Op::apply(acc, cvt<A>(e[q]));
                csum[k] = acc;
            }
        } else {
            const Tin *tin = static_cast<const Tin *>(p.in) + seg * p.n + t * TILE;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                A v[EPC];
                load_chunk<Tin, A, Op, EPC>(tin, wbase + (k * 32 + lane) * EPC, valid, v);
                Tin e[EPC];
                A acc = v[0];
                e[0] = cvt<Tin>(v[0]);
#pragma unroll
                for (int q = 1; q < EPC; ++q) { acc = Op::apply(acc, v[q]); e[q] = cvt<Tin>(v[q]); }
                csum[k] = acc;
                sts_chunk<Tin, EPC>(s_in + wbase + (k * 32 + lane) * EPC, e);   // own chunk only
            }
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                A up = shfl_up_any(csum[k], d);
                if (lane >= d) csum[k] = Op::apply(up, csum[k]);
            }
        }
        A run = Op::identity();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            A excl = shfl_up_any(csum[k], 1);
            A rtot = shfl_idx_any(csum[k], 31);
            cpre[k] = lane == 0 ? run : Op::apply(run, excl);
            run = Op::apply(run, rtot);
        }
        if (lane == 0) s_wtot[warp] = run;
    };
    // warp 0 only, after a __syncthreads that follows pass1(stage); returns the tile aggregate
    auto finish_pass1 = [&](int stage, int parity) -> A {
        const long long g = s_tile[stage];
        const int64_t seg = seg_of_tile(g, p), t = g - seg * p.tiles_per_seg;
        A wincl = lane < NWARPS ? s_wtot[lane] : Op::identity();
#pragma unroll
        for (int d = 1; d < NWARPS; d <<= 1) {
            A up = shfl_up_any(wincl, d);
            if (lane >= d) wincl = Op::apply(up, wincl);
        }
        const A tile_agg = shfl_idx_any(wincl, NWARPS - 1);
        A wexcl = shfl_up_any(wincl, 1);
        if (lane == 0) wexcl = Op::identity();
        if (lane < NWARPS) s_wexcl[parity][lane] = wexcl;
        if (lane == 0) {
            uint64_t *slots = p.status + seg * p.tiles_per_seg * NW;
            if (t == 0) {
                if (p.tiles_per_seg > 1) {
                    const A seed = scan_seed<A, Op>(p);
                    publish<A>(slots, SC_INCLUSIVE, Op::apply(seed, tile_agg));
                }
            } else {
                publish<A>(slots + t * NW, SC_PARTIAL, tile_agg);
            }
        }
        return tile_agg;
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int st = 0; st < S; ++st) { mbar_init(&s_full[st], 1); s_use[st] = 0; }
        fence_proxy_async_smem();
#pragma unroll
        for (int st = 0; st < S; ++st) fill(st, (long long)atomicAdd(p.ticket, 1ull));
    }
    __syncthreads();
    if (s_tile[0] >= total) return;

    A cpre_cur[K], cpre_nxt[K];
    A agg_cur = Op::identity(), agg_nxt = Op::identity();   // meaningful in warp 0
    pass1(0, cpre_cur);
    __syncthreads();
    if (warp == 0) agg_cur = finish_pass1(0, 0);

    for (int64_t it = 0;; ++it) {
        const int stage = (int)(it % S);
        const int nxt = (int)((it + 1) % S);
        const int par = (int)(it & 1);
        const long long g = s_tile[stage];
        const bool have_next = s_tile[nxt] < total;
        // ticket for the refill at the end of this iteration: issued now, consumed ~3 us later
        long long next_ticket = 0;
        SC_STAT_T(t_top);
        if (threadIdx.x == 0) next_ticket = (long long)atomicAdd(p.ticket, 1ull);

        // ---- pass 1 of the NEXT tile first, so that its aggregate is published a whole iteration
        // before its turn and never waits behind this CTA's own look-back (publishing after the
        // look-back re-creates the serial chain: measured 1.7 TB/s instead of 4.4-5.2 TB/s)
        if (have_next) pass1(nxt, cpre_nxt);
        SC_STAT_T(t_p1);
        SC_STAT_T(t_rdy);
        __syncthreads();                                   // (A) warp totals of the next tile visible
        SC_STAT_T(t_a);
        const int64_t seg = seg_of_tile(g, p), t = g - seg * p.tiles_per_seg;
        if (warp == 0) {
            if (have_next) agg_nxt = finish_pass1(nxt, par ^ 1);
            const A prefix = tile_lookback<Op, LOOK>(p, seg, t, lane, agg_cur, /*publish_partial=*/false);
            if (lane == 0) s_tile_prefix = prefix;
        }
        __syncthreads();                                   // (B) prefix of the current tile visible
        SC_STAT_T(t_b);
        const A base = Op::apply(s_tile_prefix, s_wexcl[par][warp]);

        // ---- pass 2 of the CURRENT tile: apply the prefix, store straight to global
        int64_t valid = p.n - t * TILE;
        if (valid > TILE) valid = TILE;
        Tout *tout = static_cast<Tout *>(p.out) + seg * p.n + t * TILE;
        const Tin *s_in = reinterpret_cast<const Tin *>(s_dyn + stage * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            Tin e[EPC];
            lds_chunk<Tin, EPC>(s_in + wbase + (k * 32 + lane) * EPC, e);
            A run2 = Op::apply(base, cpre_cur[k]);
            A o[EPC];
#pragma unroll
            for (int q = 0; q < EPC; ++q) {
                const A nxtv = Op::apply(run2, cvt<A>(e[q]));
                o[q] = p.exclusive ? run2 : nxtv;
                run2 = nxtv;
            }
            store_chunk<Tout, A, EPC>(tout, wbase + (k * 32 + lane) * EPC, valid, o);
        }
        SC_STAT_T(t_p2);
        __syncthreads();                                   // (C) stage consumed; s_wtot / s_tile_prefix free
        SC_STAT_T(t_c);
        SC_STAT_ADD(0, 1);                 // iterations
        SC_STAT_ADD(2, t_p1 - t_top);      // ticket + pass 1 (includes [1])
        SC_STAT_ADD(3, t_rdy - t_p1);      // ready check / early consume (waits for the prefetched loads)
        SC_STAT_ADD(4, t_a - t_rdy);       // barrier (A)
        SC_STAT_ADD(5, t_b - t_a);         // publish + look-back / fallback + barrier (B)
        SC_STAT_ADD(6, t_p2 - t_b);        // pass 2
        SC_STAT_ADD(8, t_c - t_p2);        // barrier (C)
        if (threadIdx.x == 0) fill(stage, next_ticket);
        if (!have_next) break;
#pragma unroll
        for (int k = 0; k < K; ++k) cpre_cur[k] = cpre_nxt[k];
        agg_cur = agg_nxt;
    }
#ifdef B200_SCAN_STATS
    if (threadIdx.x == 0)
        for (int i = 0; i < 16; ++i) if (st_acc[i]) atomicAdd(&g_scan_stats[i], st_acc[i]);
#endif
}

// ------------------------------------------------------------------ self-chained pipeline, pass 1 two tiles ahead
// Same ring and the same CHAIN arithmetic as scan_pipe_kernel<CHAIN>, but with four stages and pass 1 running
// TWO tiles ahead of pass 2.  Under a saturated memory system a status word needs ~1.7 us to come back and a
// publication about as long to become visible (measured with B200_SCAN_STATS: prefetched words not back after
// 3300 cycles, 10-16 % of them still empty), which is a whole iteration; publishing two iterations early gives
// every consumer's prefetch a full iteration of slack on both sides.
template <class Tin, class Tout, class Op, int EPC, int K, int NT, int LOOK>
__global__ void __launch_bounds__(NT + 32) scan_chain2_kernel(ScanParams p) {
    constexpr int S = 4;
    constexpr bool CHAIN = true;
    using A = typename Op::acc_t;
    constexpr int NWARPS = NT / 32;
    constexpr int NW = StatusWords<A>::NW;
    constexpr int WARP_ELEMS = 32 * EPC * K;
    constexpr int TILE = NWARPS * WARP_ELEMS;
    constexpr size_t STAGE_BYTES = (size_t)TILE * sizeof(Tin);
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) uint64_t s_full[S];
    __shared__ long long s_tile[S];        // ticket held by each stage
    __shared__ int s_use[S];               // how many times the stage has been armed (barrier parity)
    __shared__ long long s_seg[S], s_tt[S];  // segment and tile-in-segment of each stage's ticket (divided once, by the control warp)
    __shared__ A s_wexcl[3][NWARPS];       // warp-exclusive prefixes inside the tile: [tile sequence % 3][warp]
    __shared__ A s_wtot[NWARPS];
    __shared__ A s_tile_prefix;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto chunk_of = [&](int kk) { return (kk * 32 + lane) * EPC; };
    const bool ctl = warp == NWARPS;       // warp NWARPS: control (tickets, loads, publications, prefixes); 0..NWARPS-1: compute
    const int64_t total = p.tiles_per_seg * p.post;
    const int64_t wbase = (int64_t)warp * WARP_ELEMS;
#ifdef B200_SCAN_STATS
    unsigned long long st_acc[16] = {};
#endif

    // thread 0: put ticket g into `stage` and start its load
    auto fill = [&](int stage, long long g) {
        s_tile[stage] = g;
        s_use[stage] += 1;
        if (g < total) {
            int64_t seg = 0, t = g;                        // one segment: no 64-bit division on the control warp's path
            if (p.post != 1) { seg = seg_of_tile(g, p); t = g - seg * p.tiles_per_seg; }
            s_seg[stage] = seg; s_tt[stage] = t;
            if (p.n - t * TILE >= TILE) {
                mbar_expect_tx(&s_full[stage], (uint32_t)STAGE_BYTES);
                tma_load_1d(s_dyn + stage * STAGE_BYTES, static_cast<const Tin *>(p.in) + seg * p.n + t * TILE,
                            (uint32_t)STAGE_BYTES, &s_full[stage]);
            } else {
                mbar_arrive(&s_full[stage]);   // ragged tail: loaded with ordinary loads in pass 1
            }
        }
    };

    // Pass 1 over the tile in `stage` (blocks until its data has landed): chunk sums, K warp
    // scans -> chunk-exclusive prefixes inside the warp (returned in cpre), warp totals to smem.
    // After the caller's __syncthreads, finish_pass1 (warp 0) turns the warp totals into
    // warp-exclusive prefixes and PUBLISHES the tile aggregate — one tile ahead of its turn.
    auto pass1 = [&](int stage, A (&cpre)[K]) {
        const int64_t seg = s_seg[stage], t = s_tt[stage];
        int64_t valid = p.n - t * TILE;
        const bool full = valid >= TILE;
        if (valid > TILE) valid = TILE;
        Tin *s_in = reinterpret_cast<Tin *>(s_dyn + stage * STAGE_BYTES);
        SC_STAT_T(tw0);
        mbar_wait(&s_full[stage], (uint32_t)((s_use[stage] - 1) & 1));
        SC_STAT_T(tw1);
        SC_STAT_ADD(1, tw1 - tw0);          // waiting for the tile to land
        A csum[K];
        if (full) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                Tin e[EPC];
                lds_chunk<Tin, EPC>(s_in + wbase + chunk_of(k), e);
                A acc = cvt<A>(e[0]);
#pragma unroll
                for (int q = 1; q < EPC; ++q) acc = Op::apply(acc, cvt<A>(e[q]));
                csum[k] = acc;
            }
        } else {
            const Tin *tin = static_cast<const Tin *>(p.in) + seg * p.n + t * TILE;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                A v[EPC];
                load_chunk<Tin, A, Op, EPC>(tin, wbase + chunk_of(k), valid, v);
                Tin e[EPC];
                A acc = v[0];
                e[0] = cvt<Tin>(v[0]);
#pragma unroll
                for (int q = 1; q < EPC; ++q) { acc = Op::apply(acc, v[q]); e[q] = cvt<Tin>(v[q]); }
                csum[k] = acc;
                sts_chunk<Tin, EPC>(s_in + wbase + chunk_of(k), e);   // own chunk only
            }
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                A up = shfl_up_any(csum[k], d);
                if (lane >= d) csum[k] = Op::apply(up, csum[k]);
            }
        }
        A run = Op::identity();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            A excl = shfl_up_any(csum[k], 1);
            A rtot = shfl_idx_any(csum[k], 31);
            cpre[k] = lane == 0 ? run : Op::apply(run, excl);
            run = Op::apply(run, rtot);
        }
        if (lane == 0) s_wtot[warp] = run;
    };
    // warp 0 only, after a __syncthreads that follows pass1(stage); returns the tile aggregate
    auto finish_pass1 = [&](int stage, int parity) -> A {
        const int64_t seg = s_seg[stage], t = s_tt[stage];
        A wincl = lane < NWARPS ? s_wtot[lane] : Op::identity();
#pragma unroll
        for (int d = 1; d < NWARPS; d <<= 1) {
            A up = shfl_up_any(wincl, d);
            if (lane >= d) wincl = Op::apply(up, wincl);
        }
        const A tile_agg = shfl_idx_any(wincl, NWARPS - 1);
        A wexcl = shfl_up_any(wincl, 1);
        if (lane == 0) wexcl = Op::identity();
        if (lane < NWARPS) s_wexcl[parity][lane] = wexcl;
        if (lane == 0) {
            uint64_t *slots = p.status + seg * p.tiles_per_seg * NW;
            if (t == 0) {
                if (p.tiles_per_seg > 1) {
                    const A seed = scan_seed<A, Op>(p);
                    publish<A>(slots, SC_INCLUSIVE, Op::apply(seed, tile_agg));
                }
            } else {
                publish<A>(slots + t * NW, SC_PARTIAL, tile_agg);
            }
        }
        return tile_agg;
    };

    // ---- CHAIN mode (warp 0): a CTA's tiles g_0 < g_1 < ... chain through the CTA's OWN previous tile,
    //   prefix(g_{j+1}) = inclusive(g_j) (+) partial(g_j + 1) (+) ... (+) partial(g_{j+1} - 1),
    // so nobody ever waits for another tile's INCLUSIVE prefix: only for the aggregates of the tiles between
    // two of its own tickets (~#CTAs - 1 of them), which their owners publish one iteration before those
    // tiles' turn.  The status loads are ISSUED right after this CTA published tile g_{j+1}'s aggregate and
    // CONSUMED an iteration later, so the loaded-L2 round trip (~1.5 us, the critical path of the look-back
    // variant) overlaps pass 2 of g_j and pass 1 of g_{j+2}.  p.chain_at = how many chunks of pass 2 warp 0
    // processes before it issues them: later = more of the neighbours' aggregates already published (an
    // entry that comes back empty costs a full round trip when it is consumed), earlier = longer in flight.
    // Status words are kept AS LOADED and decoded only when they are folded: issuing the prefetch must not touch the
    // loaded registers, otherwise the warp waits one L2 round trip per row right there (first cut: flag tests inside
    // the issue loop, behind per-row branches: ~300 cycles x 6 rows on the control warp's path).  Lanes beyond the
    // range read the tile's own slot (always a valid address) so that the loads need no branch.
    uint64_t pf_raw[LOOK][NW];
    int pf_n = 0;                           // entries of [lo, hi] covered by the prefetch rows
    int64_t pf_lo = 0, pf_hi = -1;          // status range [lo, hi] (global tile indices) still to be folded
    int pf_base = 0;                        // 0: segment seed (first tile), 1: identity, 2: this CTA's previous inclusive
    A incl_own = Op::identity();            // inclusive prefix through this CTA's previous tile
    auto decode = [&](const uint64_t (&raw)[NW], A &value) -> uint32_t {
        uint32_t w[NW];
        uint32_t flag = (uint32_t)(raw[0] >> 32);
#pragma unroll
        for (int q = 0; q < NW; ++q) {
            if ((uint32_t)(raw[q] >> 32) != flag) flag = SC_INVALID;
            w[q] = (uint32_t)raw[q];
        }
        memcpy(&value, w, sizeof(A));
        return flag;
    };
    auto chain_issue = [&](long long g, long long t_in_seg, long long g_prev, bool have_prev) {
        const int64_t seg_start = g - t_in_seg;
        if (g == seg_start) { pf_base = 0; pf_lo = 0; pf_hi = -1; }
        else if (have_prev && g_prev >= seg_start) { pf_base = 2; pf_lo = g_prev + 1; pf_hi = g - 1; }
        else { pf_base = 1; pf_lo = seg_start; pf_hi = g - 1; }
        const int64_t span = pf_hi - pf_lo + 1;
        pf_n = (int)(span < 0 ? 0 : span > LOOK * 32 ? LOOK * 32 : span);
#pragma unroll
        for (int r = 0; r < LOOK; ++r) {
            const int e = r * 32 + lane;
            const uint64_t *slot = p.status + (e < pf_n ? pf_hi - e : (int64_t)g) * NW;
            if constexpr (NW == 2) {
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(pf_raw[r][0]), "=l"(pf_raw[r][NW - 1]) : "l"(slot) : "memory");
            } else {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(pf_raw[r][0]) : "l"(slot) : "memory");
            }
        }
    };
    // true when every prefetched status word is valid and the gap fits the prefetch (warp-uniform)
    auto chain_ready = [&]() -> bool {
        bool ok = pf_hi - pf_lo < (int64_t)LOOK * 32;
#pragma unroll
        for (int r = 0; r < LOOK; ++r) {
            A v;
            if (r * 32 + lane < pf_n) ok = ok && decode(pf_raw[r], v) != SC_INVALID;
        }
        return __all_sync(0xffffffffu, ok);
    };
    auto chain_consume = [&]() -> A {
        A val = Op::identity();
#pragma unroll
        for (int r = 0; r < LOOK; ++r) {
            const int e = r * 32 + lane;
            if (e < pf_n) {
                A v;
                uint32_t flag = decode(pf_raw[r], v);
                while (flag == SC_INVALID) flag = peek_slot<A>(p.status + (pf_hi - e) * NW, v);
                val = Op::apply(v, val);
            }
        }
        for (int64_t idx = pf_hi - (LOOK * 32 + lane); idx >= pf_lo; idx -= 32) {   // gap wider than the prefetch
            A v;
            while (peek_slot<A>(p.status + idx * NW, v) == SC_INVALID) {}
            val = Op::apply(v, val);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val = Op::apply(val, shfl_xor_any(val, o));
        const A base = pf_base == 0 ? scan_seed<A, Op>(p)
                     : pf_base == 1 ? Op::identity() : incl_own;
        return Op::apply(base, val);
    };
    __shared__ int s_chain_ready;

    const bool boss = threadIdx.x == NT;   // control warp, lane 0
    if (boss) {
#pragma unroll
        for (int st = 0; st < S; ++st) { mbar_init(&s_full[st], 1); s_use[st] = 0; }
        fence_proxy_async_smem();
#pragma unroll
        for (int st = 0; st < S; ++st) fill(st, (long long)atomicAdd(p.ticket, 1ull));
    }
    __syncthreads();
    if (s_tile[0] >= total) return;

    // cpre0 / agg0: current tile, cpre1 / agg1: next tile, cpre2 / agg2: the tile after that
    A cpre0[K], cpre1[K], cpre2[K];
    A agg0 = Op::identity(), agg1 = Op::identity(), agg2 = Op::identity();   // meaningful in the control warp
    if (!ctl) pass1(0, cpre0);
    __syncthreads();
    if (ctl) {
        agg0 = finish_pass1(0, 0);
        chain_issue(s_tile[0], s_tt[0], 0, false);
    }
    __syncthreads();                                       // s_wtot free again
    if (s_tile[1] < total) {
        if (!ctl) pass1(1, cpre1);
        __syncthreads();
        if (ctl) agg1 = finish_pass1(1, 1);
        __syncthreads();
    }

    // compute and control warps meet at these barriers from different call sites: a named barrier with an
    // explicit thread count instead of __syncthreads()
    auto cta_bar = [] { asm volatile("bar.sync 1, %0;" :: "n"(NT + 32) : "memory"); };
    for (int64_t it = 0;; ++it) {
        const int stage = (int)(it % S);
        const int st1 = (int)((it + 1) % S), st2 = (int)((it + 2) % S);
        const long long g = s_tile[stage];
        const bool have1 = s_tile[st1] < total;
        const bool have2 = have1 && s_tile[st2] < total;
        const int64_t seg = s_seg[stage], t = s_tt[stage];
        if (ctl) {
            // ================= control warp: everything serial, overlapped with the compute warps' passes
            long long next_ticket = 0;
            if (lane == 0) next_ticket = (long long)atomicAdd(p.ticket, 1ull);
            SC_STAT_T(t_top);
            const bool ready = chain_ready();              // status words requested an iteration ago
            if (ready) {
                const A prefix = chain_consume();
                incl_own = Op::apply(prefix, agg0);
                if (lane == 0) s_tile_prefix = prefix;
            }
            if (lane == 0) s_chain_ready = ready ? 1 : 0;
            // chain_at == 0 issues the next tile's status prefetch here, beside pass 1; the default (1) issues it after
            // the publication, beside pass 2 (measured on B200: 0 costs 2-3 % for Float32 and for Int64; issuing it
            // right after (A), before the publication, costs 5-10 %)
            const bool issue_early = p.chain_at == 0 && ready;
            if (issue_early && have1) chain_issue(s_tile[st1], s_tt[st1], g, true);
            SC_STAT_T(t_rdy);
            cta_bar();                               // (A)
            SC_STAT_T(t_a);
            if (have2) agg2 = finish_pass1(st2, (int)((it + 2) % 3));     // published before anything that may wait
            if (!ready) {
                const A prefix = chain_consume();
                incl_own = Op::apply(prefix, agg0);
                if (lane == 0) s_tile_prefix = prefix;
                cta_bar();                           // (B)
            }
            if (!issue_early && have1) chain_issue(s_tile[st1], s_tt[st1], g, true);
            SC_STAT_T(t_b);
            cta_bar();                               // (C)
            SC_STAT_T(t_c);
            if (lane == 0) {
#ifdef B200_SCAN_STATS
                st_acc[0] += 1; st_acc[7] += ready ? 0 : 1;
                st_acc[3] += (unsigned long long)(t_rdy - t_top);   // ready check + consume
                st_acc[4] += (unsigned long long)(t_a - t_rdy);     // waiting for pass 1 at (A)
                st_acc[5] += (unsigned long long)(t_b - t_a);       // publish + fallback + issue
                st_acc[8] += (unsigned long long)(t_c - t_b);       // waiting for pass 2 at (C)
#endif
                fill(stage, next_ticket);
            }
            if (!have1) break;
            agg0 = agg1; agg1 = agg2;
        } else {
            // ================= compute warps
            SC_STAT_T(t_top);
            if (have2) pass1(st2, cpre2);
            SC_STAT_T(t_p1);
            cta_bar();                               // (A) warp totals of tile it+2 / prefix of tile it visible
            SC_STAT_T(t_a);
            if (s_chain_ready == 0) cta_bar();       // (B) CTA-uniform
            const A base = Op::apply(s_tile_prefix, s_wexcl[it % 3][warp]);
            int64_t valid = p.n - t * TILE;
            if (valid > TILE) valid = TILE;
            Tout *tout = static_cast<Tout *>(p.out) + seg * p.n + t * TILE;
            const Tin *s_in = reinterpret_cast<const Tin *>(s_dyn + stage * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                Tin e[EPC];
                lds_chunk<Tin, EPC>(s_in + wbase + chunk_of(k), e);
                A run2 = Op::apply(base, cpre0[k]);
                A o[EPC];
#pragma unroll
                for (int q = 0; q < EPC; ++q) {
                    const A nxtv = Op::apply(run2, cvt<A>(e[q]));
                    o[q] = p.exclusive ? run2 : nxtv;
                    run2 = nxtv;
                }
                store_chunk<Tout, A, EPC>(tout, wbase + chunk_of(k), valid, o);
            }
            SC_STAT_T(t_p2);
            cta_bar();                               // (C) stage consumed; s_wtot / s_tile_prefix free
            SC_STAT_T(t_c);
            SC_STAT_ADD(2, t_p1 - t_top);                  // pass 1 (thread 0)
            SC_STAT_ADD(9, t_a - t_p1);                    // waiting at (A)
            SC_STAT_ADD(6, t_p2 - t_a);                    // pass 2
            SC_STAT_ADD(10, t_c - t_p2);                   // waiting at (C)
            if (!have1) break;
#pragma unroll
            for (int k = 0; k < K; ++k) { cpre0[k] = cpre1[k]; cpre1[k] = cpre2[k]; }
        }
    }
#ifdef B200_SCAN_STATS
    if (lane == 0 && (warp == 0 || ctl))
        for (int i = 0; i < 16; ++i) if (st_acc[i]) atomicAdd(&g_scan_stats[i], st_acc[i]);
#endif
}

// ------------------------------------------------------------------ warp-specialised pipeline
// Same persistent TMA ring, but the look-back has its own warp: NT compute threads scan tiles
// (pass 1 of tile i+1, pass 2 of tile i) while warp NT/32 resolves tile i's prefix, so a loaded
// L2 round trip (~1.5 us under 5 TB/s of traffic) is hidden behind pass 1 instead of idling the
// SM.  Compute warps and the look-back warp hand over through shared memory + mbarriers
// (s_agg / s_prefix, two slots by tile parity); compute warps synchronise among themselves with
// a named barrier so that the look-back warp never has to join a CTA-wide barrier.
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

#ifdef B200_TUNING   // measured-slower experiment (look-back in its own warp): A/B builds only
template <class Tin, class Tout, class Op, int EPC, int K, int S, int NT, int LOOK>
__global__ void __launch_bounds__(NT + 32) scan_pipe_ws_kernel(ScanParams p) {
    using A = typename Op::acc_t;
    constexpr int NWARPS = NT / 32;
    constexpr int NW = StatusWords<A>::NW;
    constexpr int WARP_ELEMS = 32 * EPC * K;
    constexpr int TILE = NWARPS * WARP_ELEMS;
    constexpr size_t STAGE_BYTES = (size_t)TILE * sizeof(Tin);
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ __align__(8) uint64_t s_full[S];
    __shared__ __align__(8) uint64_t s_agg_bar[2], s_prefix_bar[2];
    __shared__ long long s_tile[S];        // ticket held by each stage
    __shared__ int s_use[S];               // how many times the stage has been armed (barrier parity)
    __shared__ A s_wexcl[2][NWARPS];       // warp-exclusive prefixes inside the tile: [tile parity][warp]
    __shared__ A s_wtot[NWARPS];
    __shared__ A s_agg[2];                 // tile aggregate            (compute -> look-back warp)
    __shared__ long long s_agg_tile[2];    // its tile index, -1 = stop (compute -> look-back warp)
    __shared__ A s_prefix[2];              // tile exclusive prefix     (look-back warp -> compute)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool is_lb = warp == NWARPS;
    const int64_t total = p.tiles_per_seg * p.post;
    const int64_t wbase = (int64_t)warp * WARP_ELEMS;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int st = 0; st < S; ++st) { mbar_init(&s_full[st], 1); s_use[st] = 0; }
        mbar_init(&s_agg_bar[0], 1); mbar_init(&s_agg_bar[1], 1);
        mbar_init(&s_prefix_bar[0], 1); mbar_init(&s_prefix_bar[1], 1);
        fence_proxy_async_smem();
    }
    __syncthreads();                       // the only CTA-wide barrier

    if (is_lb) {
        // ================= look-back warp =================
        for (long long j = 0;; ++j) {
            const int slot = (int)(j & 1);
            mbar_wait(&s_agg_bar[slot], (uint32_t)((j >> 1) & 1));
            const long long g = s_agg_tile[slot];
            if (g < 0) break;
            const A agg = s_agg[slot];
            const int64_t seg = g / p.tiles_per_seg, t = g - seg * p.tiles_per_seg;
            const A prefix = tile_lookback<Op, LOOK>(p, seg, t, lane, agg, /*publish_partial=*/false);
            if (lane == 0) {
                s_prefix[slot] = prefix;
                mbar_arrive(&s_prefix_bar[slot]);       // release: s_prefix visible to the waiters
            }
            __syncwarp();
        }
        return;
    }

    // ================= compute warps =================
    auto fill = [&](int stage, long long g) {
        s_tile[stage] = g;
        s_use[stage] += 1;
        if (g < total) {
            const int64_t seg = g / p.tiles_per_seg, t = g - seg * p.tiles_per_seg;
            if (p.n - t * TILE >= TILE) {
                mbar_expect_tx(&s_full[stage], (uint32_t)STAGE_BYTES);
                tma_load_1d(s_dyn + stage * STAGE_BYTES, static_cast<const Tin *>(p.in) + seg * p.n + t * TILE,
                            (uint32_t)STAGE_BYTES, &s_full[stage]);
            } else {
                mbar_arrive(&s_full[stage]);
            }
        }
    };
    auto pass1 = [&](int stage, A (&cpre)[K]) {
        const long long g = s_tile[stage];
        const int64_t seg = g / p.tiles_per_seg, t = g - seg * p.tiles_per_seg;
        int64_t valid = p.n - t * TILE;
        const bool full = valid >= TILE;
        if (valid > TILE) valid = TILE;
        Tin *s_in = reinterpret_cast<Tin *>(s_dyn + stage * STAGE_BYTES);
        mbar_wait(&s_full[stage], (uint32_t)((s_use[stage] - 1) & 1));
        A csum[K];
        if (full) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                Tin e[EPC];
                lds_chunk<Tin, EPC>(s_in + wbase + (k * 32 + lane) * EPC, e);
                A acc = cvt<A>(e[0]);
#pragma unroll
                for (int q = 1; q < EPC; ++q) acc = Op::apply(acc, cvt<A>(e[q]));
                csum[k] = acc;
            }
        } else {
            const Tin *tin = static_cast<const Tin *>(p.in) + seg * p.n + t * TILE;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                A v[EPC];
                load_chunk<Tin, A, Op, EPC>(tin, wbase + (k * 32 + lane) * EPC, valid, v);
                Tin e[EPC];
                A acc = v[0];
                e[0] = cvt<Tin>(v[0]);
#pragma unroll
                for (int q = 1; q < EPC; ++q) { acc = Op::apply(acc, v[q]); e[q] = cvt<Tin>(v[q]); }
                csum[k] = acc;
                sts_chunk<Tin, EPC>(s_in + wbase + (k * 32 + lane) * EPC, e);
            }
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                A up = shfl_up_any(csum[k], d);
                if (lane >= d) csum[k] = Op::apply(up, csum[k]);
            }
        }
        A run = Op::identity();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            A excl = shfl_up_any(csum[k], 1);
            A rtot = shfl_idx_any(csum[k], 31);
            cpre[k] = lane == 0 ? run : Op::apply(run, excl);
            run = Op::apply(run, rtot);
        }
        if (lane == 0) s_wtot[warp] = run;
    };
    // warp 0, after the compute barrier that follows pass1(stage): warp-exclusive prefixes, publish the
    // aggregate to the other CTAs (status word) and hand it to the look-back warp (slot = tile parity)
    auto finish_pass1 = [&](int stage, long long j) {
        const long long g = s_tile[stage];
        const int64_t seg = g / p.tiles_per_seg, t = g - seg * p.tiles_per_seg;
        const int slot = (int)(j & 1);
        A wincl = lane < NWARPS ? s_wtot[lane] : Op::identity();
#pragma unroll
        for (int d = 1; d < NWARPS; d <<= 1) {
            A up = shfl_up_any(wincl, d);
            if (lane >= d) wincl = Op::apply(up, wincl);
        }
        const A tile_agg = shfl_idx_any(wincl, NWARPS - 1);
        A wexcl = shfl_up_any(wincl, 1);
        if (lane == 0) wexcl = Op::identity();
        if (lane < NWARPS) s_wexcl[slot][lane] = wexcl;
        if (lane == 0) {
            uint64_t *slots = p.status + seg * p.tiles_per_seg * NW;
            if (t == 0) {
                if (p.tiles_per_seg > 1) {
                    const A seed = scan_seed<A, Op>(p);
                    publish<A>(slots, SC_INCLUSIVE, Op::apply(seed, tile_agg));
                }
            } else {
                publish<A>(slots + t * NW, SC_PARTIAL, tile_agg);
            }
            s_agg[slot] = tile_agg;
            s_agg_tile[slot] = g;
            mbar_arrive(&s_agg_bar[slot]);
        }
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int st = 0; st < S; ++st) fill(st, (long long)atomicAdd(p.ticket, 1ull));
    }
    named_bar_sync(1, NT);
    if (s_tile[0] >= total) {
        if (threadIdx.x == 0) { s_agg_tile[0] = -1; mbar_arrive(&s_agg_bar[0]); }
        return;
    }

    A cpre_cur[K], cpre_nxt[K];
    pass1(0, cpre_cur);
    named_bar_sync(1, NT);
    if (warp == 0) finish_pass1(0, 0);

    for (long long j = 0;; ++j) {                           // j = sequence number of the current tile in this CTA
        const int stage = (int)(j % S);
        const int nxt = (int)((j + 1) % S);
        const int slot = (int)(j & 1);
        const long long g = s_tile[stage];
        const bool have_next = s_tile[nxt] < total;
        long long next_ticket = 0;
        if (threadIdx.x == 0) next_ticket = (long long)atomicAdd(p.ticket, 1ull);

        // pass 1 of the next tile while the look-back warp resolves the current one
        if (have_next) pass1(nxt, cpre_nxt);
        named_bar_sync(1, NT);                              // (A) warp totals of the next tile visible
        if (warp == 0) {
            if (have_next) finish_pass1(nxt, j + 1);
            else if (lane == 0) { s_agg_tile[slot ^ 1] = -1; mbar_arrive(&s_agg_bar[slot ^ 1]); }   // stop the look-back warp
        }
        mbar_wait(&s_prefix_bar[slot], (uint32_t)((j >> 1) & 1));   // prefix of the current tile
        const A base = Op::apply(s_prefix[slot], s_wexcl[slot][warp]);

        const int64_t seg = g / p.tiles_per_seg, t = g - seg * p.tiles_per_seg;
        int64_t valid = p.n - t * TILE;
        if (valid > TILE) valid = TILE;
        Tout *tout = static_cast<Tout *>(p.out) + seg * p.n + t * TILE;
        const Tin *s_in = reinterpret_cast<const Tin *>(s_dyn + stage * STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            Tin e[EPC];
            lds_chunk<Tin, EPC>(s_in + wbase + (k * 32 + lane) * EPC, e);
            A run2 = Op::apply(base, cpre_cur[k]);
            A o[EPC];
#pragma unroll
            for (int q = 0; q < EPC; ++q) {
                const A nxtv = Op::apply(run2, cvt<A>(e[q]));
                o[q] = p.exclusive ? run2 : nxtv;
                run2 = nxtv;
            }
            store_chunk<Tout, A, EPC>(tout, wbase + (k * 32 + lane) * EPC, valid, o);
        }
        named_bar_sync(1, NT);                              // (C) stage consumed; s_wtot free
        if (threadIdx.x == 0) fill(stage, next_ticket);
        if (!have_next) break;
#pragma unroll
        for (int k = 0; k < K; ++k) cpre_cur[k] = cpre_nxt[k];
    }
}

#endif  // B200_TUNING

// One thread per (i, j) column, marching along n.
template <class Tin, class Tout, class Op>
__global__ void __launch_bounds__(256) scan_strided_kernel(ScanParams p) {
    using A = typename Op::acc_t;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= p.pre * p.post) return;
    const int64_t j = o / p.pre, i = o - j * p.pre;
    const Tin *in = static_cast<const Tin *>(p.in) + i + p.pre * p.n * j;
    Tout *out = static_cast<Tout *>(p.out) + i + p.pre * p.n * j;
    A acc = scan_seed<A, Op>(p);
    constexpr int U = 8;
    int64_t r = 0;
    for (; r + U <= p.n; r += U) {
        Tin x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = ld_coh(in + p.pre * (r + u));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const A nxt = Op::apply(acc, cvt<A>(x[u]));
            out[p.pre * (r + u)] = cvt<Tout>(p.exclusive ? acc : nxt);
            acc = nxt;
        }
    }
    for (; r < p.n; ++r) {
        const A nxt = Op::apply(acc, cvt<A>(ld_coh(in + p.pre * r)));
        out[p.pre * r] = cvt<Tout>(p.exclusive ? acc : nxt);
        acc = nxt;
    }
}

// Strided scan with look-back along n: for few, long columns (e.g.
// accumulate(A; dims=2) on a 16384 x 16384 matrix) one thread per column cannot fill the GPU.
// A CTA owns a 256-column x SS_ROWS-row tile; every thread scans its column's SS_ROWS values in
// registers (row loads are coalesced across the CTA), publishes the column aggregate and resolves
// its own look-back chain over the tiles above (same status words as the 1-D scan, one chain per
// column).  Tickets are row-tile major, so every predecessor of a tile holds a smaller ticket.
constexpr int SS_ROWS = 32;

template <class Tin, class Tout, class Op>
__global__ void __launch_bounds__(256) scan_strided_lb_kernel(ScanParams p, int64_t ncb, int64_t nrt) {
    using A = typename Op::acc_t;
    constexpr int NW = StatusWords<A>::NW;
    __shared__ unsigned long long s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1ull);
    __syncthreads();
    const int64_t g = (int64_t)s_ticket;
    const int64_t rt = g / ncb, cb = g - rt * ncb;
    const int64_t cols = p.pre * p.post;
    const int64_t c = cb * 256 + threadIdx.x;
    if (c >= cols) return;
    const int64_t j = c / p.pre, i = c - j * p.pre;
    const int64_t r0 = rt * SS_ROWS;
    const int rows = (int)((p.n - r0) < SS_ROWS ? (p.n - r0) : SS_ROWS);
    const Tin *in = static_cast<const Tin *>(p.in) + i + p.pre * (r0 + p.n * j);
    Tout *out = static_cast<Tout *>(p.out) + i + p.pre * (r0 + p.n * j);
    A v[SS_ROWS];
#pragma unroll
    for (int r = 0; r < SS_ROWS; ++r) v[r] = r < rows ? cvt<A>(ld_coh(in + p.pre * r)) : Op::identity();
#pragma unroll
    for (int r = 1; r < SS_ROWS; ++r) v[r] = Op::apply(v[r - 1], v[r]);
    const A agg = v[SS_ROWS - 1];
    uint64_t *slots = p.status + c * NW;                    // slot of (rt, c) = status[(rt * cols + c) * NW]
    const int64_t rstride = cols * NW;
    A prefix;
    if (rt == 0) {
        prefix = scan_seed<A, Op>(p);
        if (nrt > 1) publish<A>(slots, SC_INCLUSIVE, Op::apply(prefix, agg));
    } else {
        publish<A>(slots + rt * rstride, SC_PARTIAL, agg);
        prefix = Op::identity();
        int64_t u = rt - 1;
        while (true) {
            A val;
            uint32_t flag;
            do { flag = peek_slot<A>(slots + u * rstride, val); } while (flag == SC_INVALID);
            prefix = Op::apply(val, prefix);
            if (flag == SC_INCLUSIVE) break;
            --u;
        }
        publish<A>(slots + rt * rstride, SC_INCLUSIVE, Op::apply(prefix, agg));
    }
    if (!p.exclusive) {
#pragma unroll
        for (int r = 0; r < SS_ROWS; ++r)
            if (r < rows) out[p.pre * r] = cvt<Tout>(Op::apply(prefix, v[r]));
    } else {
        if (rows > 0) out[0] = cvt<Tout>(prefix);
#pragma unroll
        for (int r = 1; r < SS_ROWS; ++r)
            if (r < rows) out[p.pre * r] = cvt<Tout>(Op::apply(prefix, v[r - 1]));
    }
}

// Strided scan for FEW long columns (pre * post <= 128, e.g. accumulate(A; dims=2) on a 3 x 1_000_000
// matrix, the reference's own "L" benchmark shape, perf/array.jl:8-9,72-90).  The 256-column tiling above
// would keep 3 lanes of 256 busy, so a CTA here owns `rtc` consecutive row tiles of ALL columns: thread
// (rtl, c) scans SS_ROWS rows of column c in registers, the per-thread aggregates are combined across the
// CTA's row tiles in shared memory, and one thread per column runs the look-back over the preceding CTAs
// (one status slot per (CTA, column)).
template <class Tin, class Tout, class Op>
__global__ void __launch_bounds__(256) scan_fewcols_kernel(ScanParams p, int cols, int rtc, int64_t nctas) {
    using A = typename Op::acc_t;
    constexpr int NW = StatusWords<A>::NW;
    extern __shared__ __align__(16) unsigned char s_fc_raw[];
    A *s_agg = reinterpret_cast<A *>(s_fc_raw);     // [rtc][cols]
    A *s_tot = s_agg + rtc * cols;                  // [cols] column totals of this CTA
    A *s_pref = s_tot + cols;                       // [cols] exclusive prefix of this CTA
    __shared__ unsigned long long s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1ull);
    __syncthreads();
    const int64_t g = (int64_t)s_ticket;
    const int c = threadIdx.x % cols, rtl = threadIdx.x / cols;
    const bool active = rtl < rtc;
    const int64_t j = c / p.pre, i = c - j * p.pre;
    const int64_t r0 = (g * rtc + rtl) * SS_ROWS;
    int rows = 0;
    if (active && r0 < p.n) rows = (int)((p.n - r0) < SS_ROWS ? (p.n - r0) : SS_ROWS);
    const Tin *in = static_cast<const Tin *>(p.in) + i + p.pre * (r0 + p.n * j);
    Tout *out = static_cast<Tout *>(p.out) + i + p.pre * (r0 + p.n * j);
    A v[SS_ROWS];
#pragma unroll
    for (int r = 0; r < SS_ROWS; ++r) v[r] = r < rows ? cvt<A>(ld_coh(in + p.pre * r)) : Op::identity();
#pragma unroll
    for (int r = 1; r < SS_ROWS; ++r) v[r] = Op::apply(v[r - 1], v[r]);
    const A agg = v[SS_ROWS - 1];
    if (active) s_agg[rtl * cols + c] = agg;
    __syncthreads();
    A local = Op::identity();                       // this column's aggregate over the CTA's earlier row tiles
    if (active)
        for (int k = 0; k < rtl; ++k) local = Op::apply(local, s_agg[k * cols + c]);
    if (active && rtl == rtc - 1) s_tot[c] = Op::apply(local, agg);
    __syncthreads();
    if (threadIdx.x < cols) {                       // rtl == 0: one look-back chain per column
        const A total = s_tot[c];
        uint64_t *slots = p.status + (int64_t)c * NW;
        const int64_t gstride = (int64_t)cols * NW;
        A prefix;
        if (g == 0) {
            prefix = scan_seed<A, Op>(p);
            if (nctas > 1) publish<A>(slots, SC_INCLUSIVE, Op::apply(prefix, total));
        } else {
            publish<A>(slots + g * gstride, SC_PARTIAL, total);
            prefix = Op::identity();
            int64_t u = g - 1;
            while (true) {
                A val;
                uint32_t flag;
                do { flag = peek_slot<A>(slots + u * gstride, val); } while (flag == SC_INVALID);
                prefix = Op::apply(val, prefix);
                if (flag == SC_INCLUSIVE) break;
                --u;
            }
            publish<A>(slots + g * gstride, SC_INCLUSIVE, Op::apply(prefix, total));
        }
        s_pref[c] = prefix;
    }
    __syncthreads();
    if (rows == 0) return;
    const A prefix = Op::apply(s_pref[c], local);
    if (!p.exclusive) {
#pragma unroll
        for (int r = 0; r < SS_ROWS; ++r)
            if (r < rows) out[p.pre * r] = cvt<Tout>(Op::apply(prefix, v[r]));
    } else {
        out[0] = cvt<Tout>(prefix);
#pragma unroll
        for (int r = 1; r < SS_ROWS; ++r)
            if (r < rows) out[p.pre * r] = cvt<Tout>(Op::apply(prefix, v[r - 1]));
    }
}

// One warp per contiguous segment, for many short-to-medium segments (accumulate(A; dims=1) on a
// 512 x 1000 matrix, perf/array.jl:76): no tickets, no look-back; the warp walks its segment in chunks of
// 32 * EPC elements with a running carry, the next chunk's loads issued before the current chunk's shuffles.
template <class Tin, class Tout, class Op, int EPC>
__global__ void __launch_bounds__(256) scan_warpseg_kernel(ScanParams p) {
    using A = typename Op::acc_t;
    const int lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (seg >= p.post) return;
    const Tin *tin = static_cast<const Tin *>(p.in) + seg * p.n;
    Tout *tout = static_cast<Tout *>(p.out) + seg * p.n;
    constexpr int STEP = 32 * EPC;
    A carry = scan_seed<A, Op>(p);
    A nxt[EPC];
    load_chunk<Tin, A, Op, EPC>(tin, (int64_t)lane * EPC, p.n, nxt);
    for (int64_t base = 0; base < p.n; base += STEP) {
        A v[EPC];
#pragma unroll
        for (int e = 0; e < EPC; ++e) v[e] = nxt[e];
        if (base + STEP < p.n) load_chunk<Tin, A, Op, EPC>(tin, base + STEP + (int64_t)lane * EPC, p.n, nxt);
#pragma unroll
        for (int e = 1; e < EPC; ++e) v[e] = Op::apply(v[e - 1], v[e]);
        A csum = v[EPC - 1];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            A up = shfl_up_any(csum, d);
            if (lane >= d) csum = Op::apply(up, csum);
        }
        const A excl = shfl_up_any(csum, 1);
        const A tot = shfl_idx_any(csum, 31);
        const A pre = lane == 0 ? carry : Op::apply(carry, excl);
        A o[EPC];
        if (!p.exclusive) {
#pragma unroll
            for (int e = 0; e < EPC; ++e) o[e] = Op::apply(pre, v[e]);
        } else {
            o[0] = pre;
#pragma unroll
            for (int e = 1; e < EPC; ++e) o[e] = Op::apply(pre, v[e - 1]);
        }
        store_chunk<Tout, A, EPC>(tout, base + (int64_t)lane * EPC, p.n, o);
        carry = Op::apply(carry, tot);
    }
}

// ------------------------------------------------------------------ host
template <class Tin, class Tout> struct ScanShape {
    static constexpr int WIDE = sizeof(Tin) > sizeof(Tout) ? sizeof(Tin) : sizeof(Tout);
    static constexpr int EPC = 16 / WIDE;                       // elements per 16-byte chunk of the wider type
    static constexpr int K = 8;
    static constexpr int TILE_VEC = SC_THREADS * EPC * K;
    static constexpr int K1 = 8;                                // scalar-chunk fallback
    static constexpr int TILE_SCALAR = SC_THREADS * K1;
};

template <class Tin, class Tout, class Op>
int32_t launch_scan(void *ws, size_t ws_bytes, const void *in, void *out, int64_t pre, int64_t n, int64_t post,
                    int exclusive, const void *init_host, const void *carry_dev, int carry_count, cudaStream_t stream,
                    size_t *query_bytes) {
    using A = typename Op::acc_t;
    using S = ScanShape<Tin, Tout>;
    constexpr int NW = StatusWords<A>::NW;
    ScanParams p{};
    p.in = in; p.out = out; p.pre = pre; p.n = n; p.post = post; p.exclusive = exclusive;
    p.carry_dev = carry_count > 0 ? carry_dev : nullptr;
    p.carry_count = carry_count;
    p.has_init = init_host != nullptr;
    if (init_host) {
        // init arrives as a value of the OUTPUT element type; convert to the accumulator
        Tout iv;
        memcpy(&iv, init_host, sizeof(Tout));
        A a = cvt<A>(iv);
        memcpy(&p.init_bits, &a, sizeof(A));
    }
    if (pre == 1 && n <= 128 && post >= 32LL * devinfo().sm_count) {
        // very many very short contiguous segments (accumulate(A; dims=1) on 3 x 1_000_000): a tile per
        // segment would launch a look-back chain for 3 elements; one thread per segment instead
        if (query_bytes) { *query_bytes = 0; return B200_OK; }
        const int64_t grid = cdiv(post, 256);
        if (grid > 2147483647LL) return B200_INVALID_VALUE;
        scan_strided_kernel<Tin, Tout, Op><<<(unsigned int)grid, 256, 0, stream>>>(p);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    if (pre > 1 && pre * post <= 128 && n >= 4 * SS_ROWS) {
        const int cols = (int)(pre * post);
        const int rtc = 256 / cols;                                    // row tiles per CTA
        const int64_t nctas = cdiv(n, (int64_t)rtc * SS_ROWS);
        Carver c(ws);
        unsigned long long *ticket = c.take<unsigned long long>(1);
        uint64_t *status = c.take<uint64_t>((size_t)(nctas * cols * NW));
        if (query_bytes) { *query_bytes = c.bytes(); return B200_OK; }
        if (c.bytes() > ws_bytes) return B200_WORKSPACE_TOO_SMALL;
        if (ws == nullptr) return B200_INVALID_VALUE;
        if (nctas > 2147483647LL) return B200_INVALID_VALUE;
        p.status = status; p.ticket = ticket;
        B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, c.bytes(), stream));
        const size_t smem = (size_t)(rtc * cols + 2 * cols) * sizeof(A);
        scan_fewcols_kernel<Tin, Tout, Op><<<(unsigned int)nctas, 256, smem, stream>>>(p, cols, rtc, nctas);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    if (pre > 1) {
        const int64_t cols = pre * post;
        // enough columns to fill the machine, or columns too short to tile: one thread per column
        const bool use_lb = n >= 4 * SS_ROWS && cols < 2048LL * devinfo().sm_count;
        const int64_t ncb = cdiv(cols, 256), nrt = cdiv(n, SS_ROWS);
        Carver c(ws);
        unsigned long long *ticket = c.take<unsigned long long>(1);
        uint64_t *status = c.take<uint64_t>((size_t)(use_lb ? nrt * cols * NW : 0));
        if (query_bytes) { *query_bytes = use_lb ? c.bytes() : 0; return B200_OK; }
        if (!use_lb) {
            const int64_t grid = cdiv(cols, 256);
            if (grid > 2147483647LL) return B200_INVALID_VALUE;
            scan_strided_kernel<Tin, Tout, Op><<<(unsigned int)grid, 256, 0, stream>>>(p);
            B200_CHECK_LAUNCH();
            return B200_OK;
        }
        if (c.bytes() > ws_bytes) return B200_WORKSPACE_TOO_SMALL;
        if (ws == nullptr) return B200_INVALID_VALUE;
        if (ncb * nrt > 2147483647LL) return B200_INVALID_VALUE;
        p.status = status; p.ticket = ticket;
        B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, c.bytes(), stream));
        scan_strided_lb_kernel<Tin, Tout, Op><<<(unsigned int)(ncb * nrt), 256, 0, stream>>>(p, ncb, nrt);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    // contig: vector chunks need both pointers 16-byte-chunk aligned and, with several
    // segments, every segment start aligned too
    const bool query = query_bytes != nullptr;
    const bool seg_ok = post == 1 || (n % S::EPC == 0);
    const bool aligned = query || ((reinterpret_cast<uintptr_t>(in) % (S::EPC * sizeof(Tin))) == 0 &&
                                   (reinterpret_cast<uintptr_t>(out) % (S::EPC * sizeof(Tout))) == 0);
    // the persistent kernels load tiles with cp.async.bulk: global source 16-byte aligned for every segment
    const bool tma_ok = query || ((reinterpret_cast<uintptr_t>(in) % 16) == 0 && (post == 1 || (n * sizeof(Tin)) % 16 == 0));
    const bool vec = seg_ok && aligned && S::EPC > 1 && tma_ok;
    // workspace is sized for the scalar tiling (smaller tiles => more status words) so that
    // the query does not depend on pointer alignment
    const int64_t tps_max = cdiv(n, (int64_t)(S::TILE_SCALAR < S::TILE_VEC ? S::TILE_SCALAR : S::TILE_VEC));
    Carver c(ws);
    unsigned long long *ticket = c.take<unsigned long long>(1);
    uint64_t *status = c.take<uint64_t>((size_t)(post * tps_max * NW));
    if (query) { *query_bytes = c.bytes(); return B200_OK; }
    if (c.bytes() > ws_bytes) return B200_WORKSPACE_TOO_SMALL;
    if (ws == nullptr) return B200_INVALID_VALUE;
    if (post >= 4LL * devinfo().sm_count && n <= 8192) {
        // enough segments to give every SM several warps, each short enough for one warp: no tile machinery
        const int64_t wgrid = cdiv(post, 8);
        if (wgrid > 2147483647LL) return B200_INVALID_VALUE;
        if (vec) scan_warpseg_kernel<Tin, Tout, Op, S::EPC><<<(unsigned int)wgrid, 256, 0, stream>>>(p);
        else scan_warpseg_kernel<Tin, Tout, Op, 1><<<(unsigned int)wgrid, 256, 0, stream>>>(p);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    const int64_t tile = vec ? S::TILE_VEC : S::TILE_SCALAR;
    p.tiles_per_seg = cdiv(n, tile);
    p.status = status; p.ticket = ticket;
    const int64_t grid = p.tiles_per_seg * post;
    if (grid > 2147483647LL) return B200_INVALID_VALUE;
    // one memset node clears the ticket and every status word (ticket and status are adjacent)
    const size_t clear = (size_t)((char *)(status + post * p.tiles_per_seg * NW) - (char *)ticket);
    B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, clear, stream));
    if (vec) {
        // Persistent CTAs over a ring of TMA-loaded stages.  Measured on B200 (cumsum 2^30):
        //   scan_chain2_kernel (default):  Float32 6.58 TB/s, Int64 5.65 TB/s   (480 + 32 threads, 4 stages of 52.5 KB)
        //   scan_pipe_kernel (look-back):  Float32 5.28 TB/s (one 512-thread CTA per SM, 96-wide window),
        //                                  Int64 4.82 TB/s (two 256-thread CTAs per SM)
        // The single-ahead chain and the bulk-store ring measured on the way were removed
        // again; they live in the history (commit 4518692).
        constexpr bool kWide = sizeof(A) > 4;
#ifndef B200_SCAN_NT_NARROW
#define B200_SCAN_NT_NARROW 512
#endif
#ifndef B200_SCAN_K_NARROW
#define B200_SCAN_K_NARROW 8
#endif
#ifndef B200_SCAN_NT_WIDE
#define B200_SCAN_NT_WIDE 256
#endif
#ifndef B200_SCAN_LOOK_NARROW
#define B200_SCAN_LOOK_NARROW 3
#endif
#ifndef B200_SCAN_CHAIN2_K
#define B200_SCAN_CHAIN2_K 7
#endif
        // B200_SCAN_CHAIN=0 / 2 forces the look-back kernel / chain2 for every shape; B200_SCAN_WS=1 (with the look-back
        // kernel) gives the look-back its own warp (measured slower).
        // (A/B knobs exist in -DB200_TUNING builds only; the product library reads no environment variable)
        static const int chain_knob = tune_int("B200_SCAN_CHAIN", -1);
#ifdef B200_TUNING
        static const bool use_ws = tune_int("B200_SCAN_WS", 0) != 0;
#else
        constexpr bool use_ws = false;        // scan_pipe_ws_kernel (look-back in its own warp, measured slower) is not built
#endif
        // scan_chain2_kernel: status prefetch issued 0 = before barrier (A), 1 = after the publication
        static const int chain_at = tune_int("B200_SCAN_CHAIN_AT", 1);
        p.chain_at = chain_at < 0 ? 0 : chain_at;
        int dev = 0;
        B200_CUDA_TRY(cudaGetDevice(&dev));
        // Several segments: a segment's last tile is usually ragged and ragged tiles take the slow (LDG) path of pass 1, so
        // pick the tiling that wastes less of its last tile (16384 x 16384 Float32 along dims=1: the look-back kernel's
        // 16384-element tile fits exactly, chain2's 13440 would leave every second tile ragged).
        constexpr int kLookThreads = kWide ? B200_SCAN_NT_WIDE : B200_SCAN_NT_NARROW;
        constexpr int kLookK = kWide ? S::K : B200_SCAN_K_NARROW;
        constexpr int64_t kTileChain2 = (int64_t)480 * S::EPC * B200_SCAN_CHAIN2_K;
        constexpr int64_t kTileLook = (int64_t)kLookThreads * S::EPC * kLookK;
        const int64_t pad2 = cdiv(n, kTileChain2) * kTileChain2 - n, pad1 = cdiv(n, kTileLook) * kTileLook - n;
        const bool chain2_fits = post == 1 || pad2 <= pad1 || pad2 * 8 <= n;      // at most 12.5 % of a segment in a ragged tile
        static const int chain_force = chain_knob < 0 ? -1 : (chain_knob == 2 ? 2 : 0);
        const bool use_chain2 = !use_ws && (chain_force >= 0 ? chain_force == 2 : chain2_fits);
        int64_t pgrid = 0, ptiles = 0;
#ifdef B200_SCAN_STATS
        unsigned long long zero[16] = {};
        cudaMemcpyToSymbol(g_scan_stats, zero, sizeof(zero));
#endif
        if (use_chain2) {
            constexpr int kThreads = 480, kK = B200_SCAN_CHAIN2_K;   // 15 compute warps + 1 control warp = 512 threads (128 registers each)
            constexpr size_t smem = (size_t)kTileChain2 * sizeof(Tin) * 4;
            // one CTA per SM for every type pair: the tile is sized by the OUTPUT bytes a CTA stores per iteration, and a
            // second CTA would double the status words the control warp holds in registers (12 rows: spills)
            constexpr int kRows = (148 * 5 / 4 + 31) / 32;           // status rows prefetched: the gap between two tickets + 25 %
            auto kern = scan_chain2_kernel<Tin, Tout, Op, S::EPC, kK, kThreads, kRows>;
            static thread_local int attr2_dev_mask[64];
            if (dev >= 0 && dev < 64 && !attr2_dev_mask[dev]) {
                B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr2_dev_mask[dev] = 1;
            }
            p.tiles_per_seg = cdiv(n, kTileChain2);
            ptiles = p.tiles_per_seg * post;
            pgrid = (int64_t)devinfo().sm_count;
            if (pgrid > ptiles) pgrid = ptiles;
            kern<<<(unsigned int)pgrid, kThreads + 32, smem, stream>>>(p);
        } else {
            constexpr int kStages = 3;
            constexpr int kLook = kWide ? 4 : B200_SCAN_LOOK_NARROW;
            constexpr size_t smem = (size_t)kTileLook * sizeof(Tin) * kStages;
            constexpr int kFit = (200 * 1024) / (int)smem > 0 ? (200 * 1024) / (int)smem : 1;   // CTAs per SM by shared memory
#ifdef B200_TUNING
            auto kern = use_ws ? scan_pipe_ws_kernel<Tin, Tout, Op, S::EPC, kLookK, kStages, (kLookThreads > 992 ? 992 : kLookThreads), kLook>
                               : scan_pipe_kernel<Tin, Tout, Op, S::EPC, kLookK, kStages, kLookThreads, kLook>;
#else
            auto kern = scan_pipe_kernel<Tin, Tout, Op, S::EPC, kLookK, kStages, kLookThreads, kLook>;
#endif
            static thread_local int attr_dev_mask[64][2];  // cudaFuncSetAttribute once per (thread, device, variant)
            if (dev >= 0 && dev < 64 && !attr_dev_mask[dev][use_ws ? 1 : 0]) {
                B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_dev_mask[dev][use_ws ? 1 : 0] = 1;
            }
            p.tiles_per_seg = cdiv(n, kTileLook);
            ptiles = p.tiles_per_seg * post;
            pgrid = (int64_t)devinfo().sm_count * kFit;
            if (pgrid > ptiles) pgrid = ptiles;
            kern<<<(unsigned int)pgrid, use_ws ? kLookThreads + 32 : kLookThreads, smem, stream>>>(p);
        }
#ifdef B200_SCAN_STATS
        cudaStreamSynchronize(stream);
        unsigned long long h[16];
        cudaMemcpyFromSymbol(h, g_scan_stats, sizeof(h));
        if (h[0] && use_chain2)
            fprintf(stderr, "[scan stats] chain2 grid=%lld tiles=%lld iters=%llu | compute cycles/iter: mbar-wait %.0f  pass1 %.0f  wait(A) %.0f  pass2 %.0f  wait(C) %.0f"
                    " | control: ready+consume %.0f  wait(A) %.0f  publish+fallback+issue %.0f  wait(C) %.0f | fallbacks %.1f%%\n", (long long)pgrid, (long long)ptiles, h[0],
                    (double)h[1] / h[0], (double)h[2] / h[0], (double)h[9] / h[0], (double)h[6] / h[0], (double)h[10] / h[0],
                    (double)h[3] / h[0], (double)h[4] / h[0], (double)h[5] / h[0], (double)h[8] / h[0], 100.0 * h[7] / h[0]);
        else if (h[0])
            fprintf(stderr, "[scan stats] look-back grid=%lld tiles=%lld iters=%llu | cycles/iter: mbar-wait %.0f  ticket+pass1 %.0f  barA %.0f  "
                    "publish+lookback+barB %.0f  pass2 %.0f  barC %.0f\n", (long long)pgrid, (long long)ptiles, h[0],
                    (double)h[1] / h[0], (double)h[2] / h[0], (double)h[4] / h[0], (double)h[5] / h[0], (double)h[6] / h[0], (double)h[8] / h[0]);
#endif
    }
    else scan_contig_kernel<Tin, Tout, Op, 1, S::K1><<<(unsigned int)grid, SC_THREADS, 0, stream>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // namespace b200
