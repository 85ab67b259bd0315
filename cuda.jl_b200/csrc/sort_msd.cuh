// sort_msd.cuh — hybrid MSD radix sort for 32-bit keys (keys-only `sort!`, 2^27 keys and more), sm_100a.
//
// Why: an LSD radix pass must rank stably, and a stable warp rank costs ~37 issue slots per key row
// (8 ballots) — three such passes put `sort!` at ~290 instructions per key (profiles/r1_sort_opcode_mix.md).
// Keys-only sorts do not need stability when they go most-significant-digit first (equal keys are
// indistinguishable), so every rank can be ONE shared-memory atomic.  The path:
//
//   H   msd_hist16_kernel   one read of the keys -> histogram of the top 16 key bits (65536 bins, packed 16-bit
//                           shared-memory counters, flushed to global before a field can overflow)
//   PL  msd_plan_kernel     exclusive scan of the histogram = start of every 16-bit segment; list of non-empty
//                           segments; work list of the 16-bit leaf; tile table of the 256 level-1 buckets; flags for
//                           heavy digits (skewed keys); decides whether the path runs at all (largest segment at most
//                           max(65535, n / 64) keys — one CTA sorts a segment —, at least min_avg keys per non-empty
//                           segment) — otherwise `mode` stays 0, every kernel below returns at once and the LSD passes
//                           of sort.cuh (gated on the same word) run instead
//   P1  msd_scatter_kernel<1>  partition by key byte 3 (atomic rank inside the tile, decoupled look-back over
//                           the tiles, digit runs written from shared memory) : keys -> alt
//   P2  msd_scatter_kernel<2>  the same inside each level-1 bucket by key byte 2 (tiles never straddle a bucket,
//                           the look-back chain restarts at a bucket's first tile)       : alt -> keys
//                           Both levels exist in two instantiations: plain (28 keys per thread) and heavy-digit-aware
//                           (20 keys per thread; digits holding >= 1/8 of their chain are ranked lane-privately, without
//                           atomics); the plan's flag bits let exactly one of them run.
//   LB  msd_leaf_block_kernel  inputs that are sparse at 16 bits: segments of up to 1536 / 3072 / 5120 keys sorted by one
//                           CTA in shared memory (two stable radix-256 passes over the low 16 bits)
//   L8  msd_leaf_kernel     persistent CTAs take one 16-bit segment (up to 2^18 keys) at a time: histogram of its low 16
//                           bits in a packed 8-bit-bin shared-memory table (one atomic per key), then the table is walked
//                           in value order and every value written `count` times, in place (a keys-only sort of a 16-bit
//                           residue IS its histogram)
//   L16 msd_leaf16_kernel   the same with 16-bit bins for longer segments (any length, one ticket each) and for heavy
//                           duplicates; values with more than 32767 copies spill into a per-CTA table in global memory
//
// ~170 instructions per key in total instead of ~290, three passes over HBM plus the histogram read: 2^30 UInt32 keys in
// 9.9 ms (108 Gkeys/s), Float32 `rand` keys in 10.9 ms (profiles/r2_sort_final.md).
// Replaces (for this input class) CUDACore/src/sorting.jl:909-974 `bitonic_sort!`; keys keep Julia's `isless`
// order through key_encode / key_decode exactly as in the LSD path.
#pragma once
#include <type_traits>
#include "sort.cuh"

namespace b200 {

constexpr int MSD_BINS = 65536;
constexpr int MSD_H16_THREADS = 1024;
constexpr int MSD_H16_VPT = 8;                 // 16-byte vectors per thread and iteration
constexpr int MSD_H16_WORDS = MSD_BINS / 2;    // two 16-bit counters per word

using MsdCtl = SortCtl;

// N 16-bit ranks in N/2 registers (indices are compile-time constants after unrolling).
template <int N> struct PackedRanks {
    uint32_t w[(N + 1) / 2];
    __device__ __forceinline__ void set(int j, unsigned int r) {
        if (j & 1) w[j >> 1] |= r << 16;
        else w[j >> 1] = r;
    }
    __device__ __forceinline__ unsigned int get(int j) const { return (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu); }
};

// Phase A of the leaf reads keys that phase B reads again a few microseconds later: plain L2-cached load (the
// L1::no_allocate streaming form, SASS LDG.E.NA, was measured to send 85 % of the re-reads to DRAM).
__device__ __forceinline__ uint32_t ld_keep_l2(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned int ld_cg_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Store v to p[OFF] when c > TH.  The address is computed once, unconditionally, and only the store is predicated
// (left to itself the compiler re-derives every 64-bit address under its store's predicate: 55 instructions per table
// word of the leaf's expand loop instead of ~30).
template <int TH, int OFF> __device__ __forceinline__ void st_if_gt(uint32_t *p, uint32_t v, unsigned int c) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.gt.u32 q, %2, %3;\n\t@q st.global.u32 [%0+%4], %1;\n\t}"
                 :: "l"(p), "r"(v), "r"(c), "n"(TH), "n"(OFF * 4) : "memory");
}
// Inclusive warp scan with the shuffle's own range predicate (SHFL.UP P + predicated add: 2 instructions per step).
__device__ __forceinline__ unsigned int warp_incl_scan(unsigned int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        asm volatile("{\n\t.reg .pred q;\n\t.reg .u32 t;\n\tshfl.sync.up.b32 t|q, %0, %1, 0, 0xffffffff;\n\t@q add.u32 %0, %0, t;\n\t}"
                     : "+r"(v) : "r"(o));
    return v;
}

// ------------------------------------------------------------------ H: histogram of the top 16 bits
// IDENT: the key transform is the identity (unsigned ascending) and is skipped.
template <bool IDENT>
__global__ void __launch_bounds__(MSD_H16_THREADS, 1)
msd_hist16_kernel(const uint32_t *__restrict__ keys, int64_t n, SortXform xf, unsigned int *__restrict__ hist16) {
    extern __shared__ __align__(16) unsigned int sh16[];   // [MSD_H16_WORDS]
    __shared__ int s_flush;
    const int tid = threadIdx.x;
    for (int i = tid; i < MSD_H16_WORDS / 4; i += MSD_H16_THREADS) reinterpret_cast<uint4 *>(sh16)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) s_flush = 0;
    __syncthreads();
    // A field is flushed before it can reach 2^16: an iteration adds at most 1024 * 32 = 2^15 to one field, and an
    // iteration in which some add saw a field at or above 2^14 is followed by a flush — so every iteration starts with
    // all fields <= 2^14 and ends with all fields <= 2^14 + 2^15 < 2^16.
    auto count_enc = [&](uint32_t key) {
        const uint32_t inc = (key & 0x10000u) ? 0x10000u : 1u;
        const uint32_t old = atomicAdd(&sh16[key >> 17], inc);
        if (old & 0xC000C000u) s_flush = 1;
    };
    auto count = [&](uint32_t raw) { count_enc(IDENT ? raw : key_encode<uint32_t>(raw, xf)); };
    auto flush = [&]() {
        for (int w = tid; w < MSD_H16_WORDS; w += MSD_H16_THREADS) {
            const uint32_t v = sh16[w];
            if (v) {
                sh16[w] = 0;
                if (v & 0xFFFFu) atomicAdd(hist16 + 2 * w, v & 0xFFFFu);
                if (v >> 16) atomicAdd(hist16 + 2 * w + 1, v >> 16);
            }
        }
    };
    int64_t head = (int64_t)(((16 - (reinterpret_cast<uintptr_t>(keys) & 15)) & 15) / 4);
    if (head > n) head = n;
    const uint32_t *body = keys + head;
    const int64_t nvec = (n - head) / 4;
    const int64_t tail0 = head + nvec * 4;
    if (blockIdx.x == 0 && tid < 4) {           // unaligned head and tail (< 4 elements each)
        if (tid < head) count(keys[tid]);
        if (tail0 + tid < n) count(keys[tail0 + tid]);
    }
    constexpr int64_t CHUNK_V = (int64_t)MSD_H16_THREADS * MSD_H16_VPT;
    const int64_t nchunks = (nvec + CHUNK_V - 1) / CHUNK_V;
    const int lane = tid & 31, warp = tid >> 5;
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        // warp w reads the contiguous vectors [c CHUNK_V + w 32 VPT, + 32 VPT): row k = 32 vectors = 128 consecutive keys
        const int64_t v0 = c * CHUNK_V + (int64_t)warp * (32 * MSD_H16_VPT) + lane;
        Vec16 r[MSD_H16_VPT];
        const bool whole = (c + 1) * CHUNK_V <= nvec;        // uniform: every thread has all its vectors
#pragma unroll
        for (int k = 0; k < MSD_H16_VPT; ++k) {
            const int64_t v = v0 + (int64_t)k * 32;
            if (whole || v < nvec) r[k] = ldg_stream16(body + v * 4);
        }
        if (whole) {
            if constexpr (!IDENT) {
#pragma unroll
                for (int k = 0; k < MSD_H16_VPT; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[k].w[j] = key_encode<uint32_t>(r[k].w[j], xf);
            }
            // Sorted / clustered input: a row's 128 keys fall into one bin and 32 lanes adding to one word serialise
            // (measured 4.5 ms instead of 0.9 ms at 2^30).  Row 0 is the probe; behind it every row tests itself and a
            // uniform row is counted by lane 0 alone.
            const uint32_t ref0 = __shfl_sync(0xffffffffu, r[0].w[0], 0);
            const uint32_t d0 = (r[0].w[0] ^ ref0) | (r[0].w[1] ^ ref0) | (r[0].w[2] ^ ref0) | (r[0].w[3] ^ ref0);
            if (__all_sync(0xffffffffu, (d0 >> 16) == 0)) {
#pragma unroll
                for (int k = 0; k < MSD_H16_VPT; ++k) {
                    const uint32_t ref = __shfl_sync(0xffffffffu, r[k].w[0], 0);
                    const uint32_t d = (r[k].w[0] ^ ref) | (r[k].w[1] ^ ref) | (r[k].w[2] ^ ref) | (r[k].w[3] ^ ref);
                    if (__all_sync(0xffffffffu, (d >> 16) == 0)) {
                        if (lane == 0) {
                            const uint32_t inc = ((ref & 0x10000u) ? 0x10000u : 1u) * 128u;
                            const uint32_t old = atomicAdd(&sh16[ref >> 17], inc);
                            if (old & 0xC000C000u) s_flush = 1;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) count_enc(r[k].w[j]);
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < MSD_H16_VPT; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j) count_enc(r[k].w[j]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < MSD_H16_VPT; ++k) {
                const int64_t v = v0 + (int64_t)k * 32;
                if (v < nvec) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) count(r[k].w[j]);
                }
            }
        }
        __syncthreads();
        if (s_flush) {                           // uniform: nobody writes the flag between the two barriers
            flush();
            __syncthreads();
            if (tid == 0) s_flush = 0;
            __syncthreads();
        }
    }
    __syncthreads();
    flush();
}

// ------------------------------------------------------------------ PL: scan + plan (one CTA)
struct MsdPlanParams {
    const unsigned int *hist16;     // [65536]
    unsigned int *seg_start;        // [65537] exclusive scan, seg_start[65536] = n
    unsigned int *seg_list;         // [65536] ids of the non-empty segments, ascending
    unsigned int *l1_tile_start;    // [257] first level-2 tile of every level-1 bucket
    unsigned int *list16;           // [65536] segments longer than cap8 (the 16-bit leaf's work list), ascending
    MsdCtl *ctl;
    unsigned int n;
    unsigned int tile2;             // keys per level-2 tile of the heavy-digit instantiation (also the unit of the heavy rule)
    unsigned int tile2_plain;       // keys per level-2 tile of the plain instantiation (no heavy digit at level 2)
    unsigned int cap8;              // longest segment the 8-bit leaf takes
    unsigned int cap_a;             // longest segment admitted at all (one CTA sorts a segment: load balance)
    unsigned int min_avg;           // hybrid only when n / non-empty segments >= min_avg
    unsigned int dense_avg;         // below this many keys per non-empty segment the short segments are sorted on chip ...
    unsigned int small_cap;         // ... those of at most small_cap keys (msd_leaf_block_kernel)
};

static __global__ void __launch_bounds__(1024, 1) msd_plan_kernel(MsdPlanParams p) {
    __shared__ unsigned int s_wsum[32], s_wne[32], s_wmax[32], s_wnb[32];
    __shared__ unsigned int s_excl[1024];
    __shared__ unsigned int s_t[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4 *h4 = reinterpret_cast<const uint4 *>(p.hist16) + tid * 16;   // 64 consecutive bins per thread
    unsigned int sum = 0, ne = 0, mx = 0, nb = 0;
    const unsigned int cap8 = p.cap8;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const uint4 v = h4[k];
        sum += v.x + v.y + v.z + v.w;
        ne += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
        nb += (v.x > cap8) + (v.y > cap8) + (v.z > cap8) + (v.w > cap8);
        mx = max(max(mx, v.x), max(v.y, max(v.z, v.w)));
    }
    unsigned int isum = sum, ine = ne, imx = mx, inb = nb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int a = __shfl_up_sync(0xffffffffu, isum, o);
        const unsigned int b = __shfl_up_sync(0xffffffffu, ine, o);
        const unsigned int c = __shfl_xor_sync(0xffffffffu, imx, o);
        const unsigned int d = __shfl_up_sync(0xffffffffu, inb, o);
        if (lane >= o) { isum += a; ine += b; inb += d; }
        imx = max(imx, c);
    }
    if (lane == 31) { s_wsum[warp] = isum; s_wne[warp] = ine; s_wnb[warp] = inb; }
    if (lane == 0) s_wmax[warp] = imx;
    __syncthreads();
    unsigned int bsum = 0, bne = 0, bnb = 0, tot_ne = 0, tot_nb = 0, gmx = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
        if (w < warp) { bsum += s_wsum[w]; bne += s_wne[w]; bnb += s_wnb[w]; }
        tot_ne += s_wne[w];
        tot_nb += s_wnb[w];
        gmx = max(gmx, s_wmax[w]);
    }
    unsigned int run = bsum + isum - sum;       // exclusive prefix of this thread's first bin
    unsigned int idx = bne + ine - ne;
    unsigned int idx16 = bnb + inb - nb;
    s_excl[tid] = run;
    uint4 *o4 = reinterpret_cast<uint4 *>(p.seg_start) + tid * 16;
    const unsigned int bin0 = (unsigned int)tid * 64;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const uint4 v = h4[k];
        uint4 o;
        o.x = run; if (v.x) p.seg_list[idx++] = bin0 + k * 4 + 0; if (v.x > cap8) p.list16[idx16++] = bin0 + k * 4 + 0; run += v.x;
        o.y = run; if (v.y) p.seg_list[idx++] = bin0 + k * 4 + 1; if (v.y > cap8) p.list16[idx16++] = bin0 + k * 4 + 1; run += v.y;
        o.z = run; if (v.z) p.seg_list[idx++] = bin0 + k * 4 + 2; if (v.z > cap8) p.list16[idx16++] = bin0 + k * 4 + 2; run += v.z;
        o.w = run; if (v.w) p.seg_list[idx++] = bin0 + k * 4 + 3; if (v.w > cap8) p.list16[idx16++] = bin0 + k * 4 + 3; run += v.w;
        o4[k] = o;
    }
    if (tid == 1023) p.seg_start[MSD_BINS] = run;
    if (tid == 0) s_t[0] = 0;                     // reused below as the tile table; first as the heavy-digit flags
    __syncthreads();
    {   // heavy digits (same rule as msd_scatter_kernel's `detect`): a digit holding >= 1/8 of a chain of >= 4 tiles
        const unsigned int b0 = (unsigned int)tid & ~3u;                       // first thread of this thread's level-1 bucket
        const unsigned int lo1 = s_excl[b0], hi1 = b0 + 4 < 1024 ? s_excl[b0 + 4] : p.n;
        const unsigned int tot1 = hi1 - lo1;
        unsigned int flags = 0;
        if ((tid & 3) == 0 && p.n >= 4u * p.tile2 && tot1 >= (p.n >> 3)) flags |= 1u;
        if (tot1 >= 4u * p.tile2) {
            const unsigned int th = tot1 >> 3;
#pragma unroll 4
            for (int k = 0; k < 16; ++k) {
                const uint4 v = h4[k];
                if (v.x >= th || v.y >= th || v.z >= th || v.w >= th) flags |= 2u;
            }
        }
        if (flags) atomicOr(&s_t[0], flags);
    }
    __syncthreads();
    const unsigned int heavy_flags = s_t[0];
    if (tid == 0) p.ctl->heavy = heavy_flags;
    const unsigned int tile2 = (heavy_flags & 2u) ? p.tile2 : p.tile2_plain;      // the level-2 instantiation that will run
    __syncthreads();
    // level-1 bucket b = bins [256 b, 256 b + 256) = threads 4 b .. 4 b + 3
    unsigned int tiles = 0;
    if (tid < 256) {
        const unsigned int a = s_excl[tid * 4];
        const unsigned int b = tid == 255 ? p.n : s_excl[tid * 4 + 4];
        tiles = (b - a + tile2 - 1) / tile2;
    }
    unsigned int it = tiles;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int a = __shfl_up_sync(0xffffffffu, it, o);
        if (lane >= o) it += a;
    }
    if (tid < 256 && lane == 31) s_t[warp] = it;
    __syncthreads();
    if (tid < 256) {
        unsigned int base = 0;
        for (int w = 0; w < 8; ++w) if (w < warp) base += s_t[w];
        p.l1_tile_start[tid] = base + it - tiles;
        if (tid == 255) {
            p.l1_tile_start[256] = base + it;
            p.ctl->tiles2 = base + it;
        }
    }
    if (tid == 0) {
        unsigned int mode = gmx <= p.cap_a ? 1u : 0u;
        if (tot_ne == 0 || (unsigned long long)p.n < (unsigned long long)p.min_avg * tot_ne) mode = 0;
        p.ctl->nne = tot_ne;
        p.ctl->n16 = tot_nb;             // the 8-bit leaf appends the segments it gives up on
        p.ctl->small_cap = (mode && (unsigned long long)p.n < (unsigned long long)p.dense_avg * tot_ne) ? p.small_cap : 0u;
        p.ctl->max_seg = gmx;
        p.ctl->mode = mode;
    }
}

// ------------------------------------------------------------------ P1 / P2: atomic-ranked MSD partition
struct MsdScatterParams {
    const uint32_t *keys_in;
    uint32_t *keys_out;
    unsigned int n;
    const unsigned int *seg_start;       // [65537]
    const unsigned int *l1_tile_start;   // [257]
    unsigned int *status;                // [tiles][256], zeroed
    MsdCtl *ctl;
    SortXform xf;
};

// LEVEL 1: digit = key byte 3 of the encoded key, one chain over all tiles, input = raw keys (encoded here).
// LEVEL 2: digit = key byte 2, tile t lies inside level-1 bucket `seg`, chain restarts at the bucket's first tile.
//
// Persistent CTAs, software-pipelined over tiles: once a tile's keys sit in the shared staging buffer their
// registers are dead, so the NEXT tile's global loads are issued right there — before the look-back and the
// write-out of the current tile — and the DRAM round trip hides behind them (measured before this change: 32 %
// issue utilisation, barrier + scoreboard stalls on top, profiles/r2_sort_msd_v1.md).
struct MsdTile {
    unsigned int t, first, seg, base;   // ticket, chain's first ticket, level-1 bucket, index of the tile's first key
    int valid;                          // keys in the tile
};

// HEAVY: the instantiation that knows about heavy digits (skewed keys, see `detect`); the plan kernel's ctl->heavy bit of
// this level says which of the two instantiations runs — keeping the extra code out of the common one is worth ~3 %
// on uniform keys (register allocation of the hot loop).
template <int LEVEL, int NT, int ITEMS, bool IDENT, bool HEAVY>
__global__ void __launch_bounds__(NT, (NT == 512 && ITEMS <= 16) ? 3 : 1024 / NT) msd_scatter_kernel(MsdScatterParams p) {
    using LB = LookBits<uint32_t>;
    constexpr int TILE = NT * ITEMS;
    constexpr int SHIFT = LEVEL == 1 ? 24 : 16;
    static_assert(NT >= 256 && NT % 32 == 0, "one thread per digit");
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_raw);      // [TILE] digit-sorted staging
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned int s_doff[256];
    __shared__ unsigned int s_gbase[256];
    __shared__ unsigned int s_wsum[8];
    __shared__ unsigned int s_tk[2];
    __shared__ unsigned int s_seg[2][4];                     // {bucket, first tile, first key, end key} of a ticket
    __shared__ unsigned int s_l1ts[LEVEL == 2 ? 257 : 1];    // first tile of every level-1 bucket (loaded once)
    __shared__ unsigned int s_l1lo[LEVEL == 2 ? 257 : 1];    // first key of every level-1 bucket
    __shared__ unsigned int s_hv[HEAVY ? 2 : 1][4];          // heavy digits of the chain a ticket belongs to (see `detect`)
    __shared__ unsigned int s_nh[2];

    if (ld_cg_u32(&p.ctl->mode) == 0) return;
    if ((((ld_cg_u32(&p.ctl->heavy) >> (LEVEL - 1)) & 1u) != 0u) != HEAVY) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool dthread = NT == 256 || tid < 256;
    const unsigned int tiles = LEVEL == 1 ? (p.n + TILE - 1) / TILE : ld_cg_u32(&p.ctl->tiles2);
    unsigned int *ticket = LEVEL == 1 ? &p.ctl->ticket1 : &p.ctl->ticket2;

    // which bucket does ticket t fall into?  (LEVEL 2; the one digit thread whose range holds t answers)
    auto locate = [&](unsigned int t, int slot) {
        if constexpr (LEVEL == 2) {
            if (dthread && t < tiles) {
                const unsigned int a = s_l1ts[tid], b = s_l1ts[tid + 1];
                if (a <= t && t < b) {
                    s_seg[slot][0] = tid; s_seg[slot][1] = a;
                    s_seg[slot][2] = s_l1lo[tid]; s_seg[slot][3] = s_l1lo[tid + 1];
                }
            }
        }
    };
    auto describe = [&](unsigned int t, int slot) {
        MsdTile d;
        d.t = t;
        unsigned int lo = 0, hi = p.n;
        d.seg = 0; d.first = 0;
        if constexpr (LEVEL == 2) { d.seg = s_seg[slot][0]; d.first = s_seg[slot][1]; lo = s_seg[slot][2]; hi = s_seg[slot][3]; }
        d.base = lo + (t - d.first) * TILE;
        const unsigned int remain = hi - d.base;
        d.valid = remain > (unsigned int)TILE ? TILE : (int)remain;
        return d;
    };
    uint32_t keys[ITEMS];
    auto load_tile = [&](const MsdTile &d) {
        const uint32_t *kin = p.keys_in + d.base + tid;
        if (d.valid == TILE) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) keys[i] = ldg_stream(kin + i * NT);
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) keys[i] = (i * NT + tid < d.valid) ? ldg_stream(kin + i * NT) : 0u;
        }
    };

    // Skewed digits (Float32 keys: half of uniform [0, 1) floats share ONE top byte): 16 lanes of a warp adding to
    // one shared counter serialise.  A digit that holds at least 1/8 of its chain's keys is "heavy": its keys are ranked
    // by ballot + ONE atomic of the leading lane (at most 4 heavy digits get this treatment; the digit threads find
    // them from the bucket's exact counts, an iteration before the tile is ranked).
    auto detect = [&](int slot, unsigned int my_cnt, unsigned int total) {
        if (total >= 4u * (unsigned int)TILE && my_cnt >= (total >> 3)) {
            const unsigned int k = atomicAdd(&s_nh[slot], 1u);
            if (k < 4) s_hv[slot][k] = (unsigned int)tid;
        }
    };

    // ---- prologue: tables, first ticket, first tile's loads
    unsigned int next_t = 0;                       // thread 0: ticket after the current one, requested a tile early
    if (tid == 0) { s_tk[0] = atomicAdd(ticket, 1u); s_nh[0] = 0; s_nh[1] = 0; }   // s_nh stays 0 in the !HEAVY instantiation
    if constexpr (LEVEL == 2) {
        for (int i = tid; i < 257; i += NT) {
            s_l1ts[i] = p.l1_tile_start[i];
            s_l1lo[i] = p.seg_start[i * 256];
        }
    }
    if (dthread) s_hist[tid] = 0;
    __syncthreads();
    if (s_tk[0] >= tiles) return;
    if (tid == 0) next_t = atomicAdd(ticket, 1u);
    locate(s_tk[0], 0);
    if constexpr (LEVEL == 2) __syncthreads();
    MsdTile cur = describe(s_tk[0], 0);
    load_tile(cur);
    if constexpr (HEAVY) {
        if (dthread) {
            if constexpr (LEVEL == 1) detect(0, p.seg_start[(tid + 1) * 256] - p.seg_start[tid * 256], p.n);
            else detect(0, p.seg_start[cur.seg * 256 + tid + 1] - p.seg_start[cur.seg * 256 + tid], s_seg[0][3] - s_seg[0][2]);
        }
    }
    __syncthreads();

    for (unsigned int it = 0;; ++it) {
        const bool full = cur.valid == TILE;
        const bool chain_head = cur.t == cur.first;
        if (tid == 0) {
            s_tk[(it + 1) & 1] = next_t;          // the ticket requested during the previous tile has long arrived
            if constexpr (LEVEL == 2 && HEAVY) s_nh[(it + 1) & 1] = 0;     // last read two tiles ago; next written after (C)
        }
        // first output index of this bucket's digit `tid` (loaded early, used after the look-back)
        unsigned int bin = 0;
        if (dthread) bin = LEVEL == 1 ? p.seg_start[tid * 256] : p.seg_start[cur.seg * 256 + tid];

        // ---- rank with one shared atomic per key (s_hist is zero: prologue / zeroed below by its digit thread)
        if constexpr (LEVEL == 1 && !IDENT) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) keys[i] = key_encode<uint32_t>(keys[i], p.xf);
        }
        PackedRanks<ITEMS> ranks;
        // Sorted / clustered input: whole rows (32 consecutive keys) carry ONE digit, and thousands of atomics on one
        // shared counter would serialise (measured: P2 of a sorted 2^30 input 36 ms instead of 2.7 ms).  Such rows
        // reserve their 32 ranks with a single atomic by lane 0.
        // Row 0 is the probe: random data never has 32 equal digits in a row, sorted / clustered data nearly always.
        bool clustered = false;
        if (full) {
            const unsigned int d = (keys[0] >> SHIFT) & 0xFFu;
            clustered = __all_sync(0xffffffffu, d == __shfl_sync(0xffffffffu, d, 0));
        }
        if (clustered) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const unsigned int d = (keys[i] >> SHIFT) & 0xFFu;
                const unsigned int d0 = __shfl_sync(0xffffffffu, d, 0);
                unsigned int r;
                if (__all_sync(0xffffffffu, d == d0)) {            // warp-uniform branch
                    unsigned int b = 0;
                    if (lane == 0) b = atomicAdd(&s_hist[d0], 32u);
                    r = __shfl_sync(0xffffffffu, b, 0) + lane;
                } else {
                    r = atomicAdd(&s_hist[d], 1u);                 // the row that holds a digit boundary
                }
                ranks.set(i, r);
            }
        } else if (HEAVY && full && s_nh[LEVEL == 1 ? 0 : (it & 1)] != 0) {
            // heavy digits: no atomic and no ballot per key.  Pass 1: every lane counts its own keys of each heavy digit;
            // one warp scan + ONE atomic per heavy digit reserve a block of ranks per lane; pass 2 hands a lane's ranks
            // out in item order (equal keys are indistinguishable, any unique rank will do).  The other digits keep
            // their per-key atomics.
            const int hslot = LEVEL == 1 ? 0 : (int)(it & 1);
            auto heavy_rank = [&](auto nh_tag) {
                constexpr int NH = decltype(nh_tag)::value;
                unsigned int hv[NH], nxt_rank[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) { hv[h] = s_hv[hslot][h]; nxt_rank[h] = 0; }
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const unsigned int d = (keys[i] >> SHIFT) & 0xFFu;
#pragma unroll
                    for (int h = 0; h < NH; ++h) nxt_rank[h] += (d == hv[h]) ? 1u : 0u;
                }
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const unsigned int incl = warp_incl_scan(nxt_rank[h]);
                    unsigned int b = 0;
                    if (lane == 31 && incl) b = atomicAdd(&s_hist[hv[h]], incl);
                    nxt_rank[h] = __shfl_sync(0xffffffffu, b, 31) + incl - nxt_rank[h];
                }
#pragma unroll
                for (int i = 0; i < ITEMS; ++i) {
                    const unsigned int d = (keys[i] >> SHIFT) & 0xFFu;
                    unsigned int r = 0xFFFFFFFFu;
#pragma unroll
                    for (int h = 0; h < NH; ++h)
                        if (d == hv[h]) { r = nxt_rank[h]; nxt_rank[h] += 1u; }
                    if (r == 0xFFFFFFFFu) r = atomicAdd(&s_hist[d], 1u);
                    ranks.set(i, r);
                }
            };
            const unsigned int nh = s_nh[hslot];                   // uniform over the CTA
            if (nh == 1) heavy_rank(std::integral_constant<int, 1>{});
            else if (nh == 2) heavy_rank(std::integral_constant<int, 2>{});
            else if (nh == 3) heavy_rank(std::integral_constant<int, 3>{});
            else heavy_rank(std::integral_constant<int, 4>{});
        } else if (full) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) ranks.set(i, atomicAdd(&s_hist[(keys[i] >> SHIFT) & 0xFFu], 1u));
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                unsigned int r = 0;
                if (i * NT + tid < cur.valid) r = atomicAdd(&s_hist[(keys[i] >> SHIFT) & 0xFFu], 1u);
                ranks.set(i, r);
            }
        }
        __syncthreads();                                                                     // (A)

        // ---- per digit: publish the tile's count, exclusive scan over the digits.  No barrier between the warp scans
        // and the warp bases: every digit warp sums the lower warps' counters itself (<= 7 conflict-free loads per lane
        // and one warp reduction), which is cheaper than parking 16 warps at a barrier for 8 numbers.
        unsigned int cnt = 0, doff = 0;
        if (dthread) {                             // whole warps (tid < 256)
            cnt = s_hist[tid];
            st_relaxed<uint32_t>(p.status + (size_t)cur.t * 256 + tid, (chain_head ? LB::INCLUSIVE : LB::PARTIAL) | cnt);
            const unsigned int incl = warp_incl_scan(cnt);
            unsigned int lower = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k) lower += (k < warp) ? s_hist[32 * k + lane] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lower += __shfl_xor_sync(0xffffffffu, lower, o);
            doff = lower + incl - cnt;
            s_doff[tid] = doff;
        }
        const unsigned int tn = s_tk[(it + 1) & 1];       // written before barrier (A)
        const bool have_next = tn < tiles;
        locate(tn, (it + 1) & 1);
        __syncthreads();                                                                     // (C)
        if (dthread) s_hist[tid] = 0;              // every digit warp has read the counters; the next ranking is behind barrier (D)

        // ---- the first look-back round trip is requested before the scatter and consumed after it: the predecessors
        // published their counts before their own scatter, so these words are (at least) PARTIAL when they arrive
        uint32_t wpre[RX_LOOK];
#pragma unroll
        for (int k = 0; k < RX_LOOK; ++k) wpre[k] = LB::INCLUSIVE;
        if (dthread && !chain_head) {
#pragma unroll
            for (int k = 0; k < RX_LOOK; ++k)
                if ((int)cur.t - 1 - k >= (int)cur.first) wpre[k] = ld_relaxed<uint32_t>(p.status + (size_t)(cur.t - 1 - k) * 256 + tid);
        }

        // ---- scatter into shared memory, grouped by digit
        if (full) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) s_keys[s_doff[(keys[i] >> SHIFT) & 0xFFu] + ranks.get(i)] = keys[i];
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                if (i * NT + tid < cur.valid) s_keys[s_doff[(keys[i] >> SHIFT) & 0xFFu] + ranks.get(i)] = keys[i];
        }

        // ---- the key registers are dead: start the next tile's loads now
        MsdTile nxt = cur;
        if (have_next) {
            if (tid == 0) next_t = atomicAdd(ticket, 1u);
            nxt = describe(tn, (it + 1) & 1);
            load_tile(nxt);
            if constexpr (LEVEL == 2 && HEAVY) {
                if (nxt.seg != cur.seg) {                 // a new chain: its heavy digits
                    if (dthread) {
                        const unsigned int a = p.seg_start[nxt.seg * 256 + tid], b = p.seg_start[nxt.seg * 256 + tid + 1];
                        detect((int)((it + 1) & 1), b - a, s_seg[(it + 1) & 1][3] - s_seg[(it + 1) & 1][2]);
                    }
                } else if (tid < 5) {                     // same chain: the list carries over
                    if (tid < 4) s_hv[(it + 1) & 1][tid] = s_hv[it & 1][tid];
                    else s_nh[(it + 1) & 1] = s_nh[it & 1];
                }
            }
        }

        // ---- decoupled look-back over this chain's earlier tiles
        if (dthread) {
            unsigned int excl = 0;
            if (!chain_head) {
                int u = (int)cur.t - 1;
                const int floor_t = (int)cur.first;
                bool done = false;
                bool pre = true;                           // first round: the words requested before the scatter
                while (!done) {
                    uint32_t w[RX_LOOK];
                    bool all_valid = true;
                    if (pre) {
#pragma unroll
                        for (int k = 0; k < RX_LOOK; ++k) { w[k] = wpre[k]; all_valid = all_valid && ((w[k] & LB::FLAGS) != 0); }
                        pre = false;
                    } else {
                        all_valid = false;
                    }
                    while (!all_valid) {
                        all_valid = true;
#pragma unroll
                        for (int k = 0; k < RX_LOOK; ++k) {
                            w[k] = LB::INCLUSIVE;          // before the chain's first tile: inclusive, zero
                            if (u - k >= floor_t) w[k] = ld_relaxed<uint32_t>(p.status + (size_t)(u - k) * 256 + tid);
                            all_valid = all_valid && ((w[k] & LB::FLAGS) != 0);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < RX_LOOK; ++k) {
                        if (!done) {
                            excl += w[k] & LB::MASK;
                            done = (w[k] & LB::FLAGS) == LB::INCLUSIVE;
                        }
                    }
                    u -= RX_LOOK;
                }
                st_relaxed<uint32_t>(p.status + (size_t)cur.t * 256 + tid, LB::INCLUSIVE | (excl + cnt));
            }
            s_gbase[tid] = bin + excl - doff;
        }
        __syncthreads();                                                                     // (D)

        // ---- write out: consecutive threads -> consecutive addresses inside each digit run
        uint32_t *kout = p.keys_out;
        if (full) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int i = tid + k * NT;
                const uint32_t key = s_keys[i];
                kout[s_gbase[(key >> SHIFT) & 0xFFu] + (unsigned int)i] = key;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int i = tid + k * NT;
                if (i < cur.valid) {
                    const uint32_t key = s_keys[i];
                    kout[s_gbase[(key >> SHIFT) & 0xFFu] + (unsigned int)i] = key;
                }
            }
        }
        if (!have_next) break;
        cur = nxt;
        // Hazards into the next iteration: s_hist was re-zeroed by its owner before (B); the next scatter into s_keys
        // and the next s_gbase / s_doff / s_seg / s_tk writes all sit behind barrier (A) of the next iteration, which no
        // thread passes before every thread has finished the write-out above.
    }
}

// ------------------------------------------------------------------ L: leaf, one 16-bit segment per CTA iteration
// All keys of a segment share their top 16 bits, and equal keys are indistinguishable: the sorted segment is fully
// described by the HISTOGRAM of the low 16 bits.  With dense keys (2^30 keys over a 32-bit range: 4 bins per key) the
// cheapest sort is therefore: count (one shared atomic per key into a packed 65536-bin table), then walk the table in
// order and emit every value `count` times.
//
//   msd_leaf_kernel   : 8-bit bins, four per word, 64 KB table, 3 CTAs per SM; segments up to `cap8` keys, dealt
//                       round-robin.  A bin that reaches 256 copies is detected by the table's byte sum and the segment
//                       is appended to `list16`.
//   msd_leaf16_kernel : 16-bit bins, two per word, 128 KB table, 1 CTA per SM; takes `list16` — the segments longer than
//                       `cap8` (listed by the plan kernel) and the ones the 8-bit kernel gave up on — ONE SEGMENT PER
//                       TICKET, whatever its length (Float32 keys: the 128 segments of [0.5, 1) hold 2^22 keys each).
//                       Values with 32768 copies or more spill into a per-CTA table in global memory.
struct MsdLeafParams {
    uint32_t *keys;                   // encoded keys grouped by their top 16 bits; sorted + decoded in place
    const unsigned int *seg_start;    // [65537]
    const unsigned int *seg_list;     // non-empty segment ids
    unsigned int *list16;             // [65536] segment ids for the 16-bit kernel; ctl->n16 entries
    unsigned int *spill;              // [grid of the 16-bit kernel][65536] all zero between segments
    MsdCtl *ctl;
    SortXform xf;
    unsigned int cap8;                // longest segment the 8-bit kernel takes
};

constexpr int MSD_LEAF_CHUNK = 8;      // global loads in flight per thread in the count phase
constexpr unsigned int MSD_LEAF_HEAVY = 1024;   // 16-bit kernel: values with at least this many copies (and 1/64 of the segment) are written by the whole CTA

template <int NT, bool IDENT>
__global__ void __launch_bounds__(NT, 3) msd_leaf_kernel(MsdLeafParams p) {
    constexpr int PER_WORD = 4;                       // bins per 32-bit word
    constexpr int WORDS = MSD_BINS / PER_WORD;        // 16384
    constexpr int NW = NT / 32;
    constexpr int WORDS_PER_WARP = WORDS / NW;        // contiguous slice of the table walked by one warp
    constexpr int CH = MSD_LEAF_CHUNK;
    extern __shared__ __align__(16) unsigned int s_bins[];   // [WORDS] packed counters, all zero between segments
    __shared__ unsigned int s_wtot[NW];

    if (ld_cg_u32(&p.ctl->mode) == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int nne = ld_cg_u32(&p.ctl->nne);
    const int small_cap = (int)ld_cg_u32(&p.ctl->small_cap);       // segments up to here were sorted by msd_leaf_block_kernel
    for (int i = tid; i < WORDS / 4; i += NT) reinterpret_cast<uint4 *>(s_bins)[i] = make_uint4(0, 0, 0, 0);

    // Segments are dealt round-robin to the persistent CTAs (independent work: no ticket).  Bounds of the next segment
    // and the id of the one after are loaded an iteration early; every thread loads the same words (broadcast).
    const unsigned int G = gridDim.x;
    unsigned int k = blockIdx.x;
    if (k >= nne) return;
    unsigned int seg, lo, hi, seg_n1 = 0;
    seg = __ldg(p.seg_list + k);
    lo = __ldg(p.seg_start + seg);
    hi = __ldg(p.seg_start + seg + 1);
    if (k + G < nne) seg_n1 = __ldg(p.seg_list + k + G);
    __syncthreads();

    for (; k < nne; k += G) {
        unsigned int nlo = 0, nhi = 0, seg_n2 = 0;
        if (k + G < nne) { nlo = __ldg(p.seg_start + seg_n1); nhi = __ldg(p.seg_start + seg_n1 + 1); }
        if (k + 2 * G < nne) seg_n2 = __ldg(p.seg_list + k + 2 * G);
        const int m = (int)(hi - lo);
        if (m <= (int)p.cap8 && m > small_cap) {             // uniform over the CTA (longer ones: the plan kernel listed them)
            uint32_t *base = p.keys + lo;
            // ---- count: one shared atomic per key.  Loads go out CH at a time (a load -> atomic -> load chain would
            // expose one memory round trip per key).  The add needs no return value and no per-key overflow test — a
            // field that wraps loses 255 (carry into the next field) or 256 (carry out of the word) from the table's byte
            // sum, so "sum of all bins != m" after the count detects every overflow (fields hold up to 255 copies).
            auto count_one = [&](uint32_t key) {
                const unsigned int inc = __funnelshift_l(0u, 1u, key << 3);              // 1 << 8 * (key & 3)
                atomicAdd(reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned char *>(s_bins) + (key & 0xFFFCu)), inc);
            };
            for (int i0 = 0; i0 < m; i0 += CH * NT) {
                uint32_t kk[CH];
                if (i0 + CH * NT <= m) {
#pragma unroll
                    for (int c = 0; c < CH; ++c) kk[c] = ld_coh(base + i0 + c * NT + tid);
#pragma unroll
                    for (int c = 0; c < CH; ++c) count_one(kk[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int i = i0 + c * NT + tid;
                        kk[c] = i < m ? ld_coh(base + i) : 0u;
                    }
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        if (i0 + c * NT + tid < m) count_one(kk[c]);
                }
            }
            __syncthreads();
            // ---- warp totals: warp w owns table words [w * WORDS_PER_WARP, (w + 1) * WORDS_PER_WARP)
            {
                const unsigned int *wb = s_bins + warp * WORDS_PER_WARP;
                unsigned int acc = 0;
#pragma unroll 8
                for (int j = lane; j < WORDS_PER_WARP; j += 32) acc += __dp4a(wb[j], 0x01010101u, 0u);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) s_wtot[warp] = acc;
            }
            __syncthreads();
            unsigned int run = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const unsigned int t = s_wtot[w];
                run += (w < warp) ? t : 0u;
                total += t;
            }
            if (total != (unsigned int)m) {                               // uniform: some 8-bit field wrapped
                // a value occurs 256 times or more: hand the segment to the 16-bit kernel and clear the table
                __syncthreads();                                          // every thread has read s_wtot / the table
                for (int i = tid; i < WORDS / 4; i += NT) reinterpret_cast<uint4 *>(s_bins)[i] = make_uint4(0, 0, 0, 0);
                if (tid == 0) p.list16[atomicAdd(&p.ctl->n16, 1u)] = seg;
            } else {
                // ---- expand: walk the table in value order; 2 x 32 words (256 values) per step and warp.
                // A lane writes its word's values at consecutive positions, lanes back to back.  Counts are nearly always
                // 0 or 1: first and second copies are straight-line predicated stores at precomputed addresses (98 % of the
                // non-empty bins hold one key, 2.4 % two); a bin with three or more copies anywhere in the word takes a loop.
                const uint32_t hi16 = seg << 16;                          // the segment id IS the keys' top 16 bits
                unsigned int *wbw = s_bins + warp * WORDS_PER_WARP;
                uint32_t *const kall = p.keys;                            // 32-bit element index + kernel parameter: one IMAD.WIDE per address
                auto emit = [&](unsigned int w, unsigned int pos, uint32_t v0) {
                    const unsigned int g0 = lo + pos;                     // element index into the whole key array (< 2^30)
                    const unsigned int c0 = w & 0xFFu, c1 = (w >> 8) & 0xFFu, c2 = (w >> 16) & 0xFFu, c3 = w >> 24;
                    const unsigned int o1 = c0, o2 = o1 + c1, o3 = o2 + c2;
                    const uint32_t k0 = IDENT ? v0 : key_decode<uint32_t>(v0, p.xf);
                    const uint32_t k1 = IDENT ? v0 + 1 : key_decode<uint32_t>(v0 + 1, p.xf);
                    const uint32_t k2 = IDENT ? v0 + 2 : key_decode<uint32_t>(v0 + 2, p.xf);
                    const uint32_t k3 = IDENT ? v0 + 3 : key_decode<uint32_t>(v0 + 3, p.xf);
                    uint32_t *out = kall + g0, *a1 = kall + (g0 + o1), *a2 = kall + (g0 + o2), *a3 = kall + (g0 + o3);
                    st_if_gt<0, 0>(out, k0, c0);
                    st_if_gt<0, 0>(a1, k1, c1);
                    st_if_gt<0, 0>(a2, k2, c2);
                    st_if_gt<0, 0>(a3, k3, c3);
                    st_if_gt<1, 1>(out, k0, c0);
                    st_if_gt<1, 1>(a1, k1, c1);
                    st_if_gt<1, 1>(a2, k2, c2);
                    st_if_gt<1, 1>(a3, k3, c3);
                    if ((w & 0xFCFCFCFCu) | (w & (w >> 1) & 0x01010101u)) {       // some count >= 3
                        const unsigned int mx = max(max(c0, c1), max(c2, c3));
#pragma unroll 1
                        for (unsigned int kk = 2; kk < mx; ++kk) {
                            if (c0 > kk) out[kk] = k0;
                            if (c1 > kk) a1[kk] = k1;
                            if (c2 > kk) a2[kk] = k2;
                            if (c3 > kk) a3[kk] = k3;
                        }
                    }
                };
                // two table rows (2 x 32 words) per step: their word counts share ONE warp scan (16-bit halves;
                // a row holds at most 32 * 4 * 255 keys)
                for (int j = lane; j < WORDS_PER_WARP; j += 64) {
                    const unsigned int wA = wbw[j], wB = wbw[j + 32];
                    wbw[j] = 0; wbw[j + 32] = 0;                      // table is clean again for the next segment
                    const unsigned int cA = __dp4a(wA, 0x01010101u, 0u), cB = __dp4a(wB, 0x01010101u, 0u);
                    const unsigned int incl = warp_incl_scan(cA | (cB << 16));
                    const unsigned int tot = __shfl_sync(0xffffffffu, incl, 31);
                    const unsigned int totA = tot & 0xFFFFu, totB = tot >> 16;
                    const uint32_t v0 = hi16 | (uint32_t)((warp * WORDS_PER_WARP + j) * PER_WORD);
                    // no `if (w)` around the calls: an empty word's stores are all predicated off
                    emit(wA, run + (incl & 0xFFFFu) - cA, v0);
                    emit(wB, run + totA + (incl >> 16) - cB, v0 + 32 * PER_WORD);
                    run += totA + totB;
                }
            }
            __syncthreads();        // table clean / list16 written before the next segment's atomics
        }
        lo = nlo; hi = nhi; seg = seg_n1; seg_n1 = seg_n2;
    }
}

// ------------------------------------------------------------------ L (sparse): short segments sorted on chip
// The counting leaves walk 65536 bins per segment whatever its length — 2^32 bins per sort, i.e. 4 per key at 2^30 keys
// but 32 per key at 2^27.  When the input is sparse at 16 bits (fewer than dense_avg keys per non-empty segment), the
// segments of up to `small_cap` keys are instead sorted by ONE CTA entirely in shared memory: two stable radix-256
// passes over the low 16 key bits with the ballot ranking of the block sort (sort_block.cuh, radix_block_sort_kernel —
// the same code path the sort!(A; dims) tests exercise), keys already in ordered-key space, decoded on the way out.
// Stable ranking makes duplicates a non-issue; padding keys (all ones) sort behind every real key.
template <int ITEMS, bool IDENT>
__global__ void __launch_bounds__(RX_THREADS, ITEMS > 12 ? 3 : 4) msd_leaf_block_kernel(MsdLeafParams p) {
    constexpr int TILE = RX_THREADS * ITEMS;
    constexpr int WARP_ITEMS = 32 * ITEMS;
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned int *s_whist = reinterpret_cast<unsigned int *>(s_raw);                 // [8][256]
    uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_whist + RX_WARPS * RX_RADIX);  // [TILE]
    __shared__ unsigned int s_wsum[RX_WARPS];
    if (ld_cg_u32(&p.ctl->mode) == 0) return;
    int small_cap = (int)ld_cg_u32(&p.ctl->small_cap);
    if (small_cap == 0) return;
    if (small_cap > TILE) small_cap = TILE;                    // (the host passes small_cap == TILE)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int nne = ld_cg_u32(&p.ctl->nne);
    const unsigned int G = gridDim.x;
    unsigned int k = blockIdx.x;
    if (k >= nne) return;
    unsigned int seg = __ldg(p.seg_list + k), lo = __ldg(p.seg_start + seg), hi = __ldg(p.seg_start + seg + 1), seg_n1 = 0;
    if (k + G < nne) seg_n1 = __ldg(p.seg_list + k + G);
    uint32_t *const kall = p.keys;
    const int idx0 = warp * WARP_ITEMS + lane;
    unsigned int *whist = s_whist + warp * RX_RADIX;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int d = tid;
    for (; k < nne; k += G) {
        unsigned int nlo = 0, nhi = 0, seg_n2 = 0;
        if (k + G < nne) { nlo = __ldg(p.seg_start + seg_n1); nhi = __ldg(p.seg_start + seg_n1 + 1); }
        if (k + 2 * G < nne) seg_n2 = __ldg(p.seg_list + k + 2 * G);
        const int m = (int)(hi - lo);
        if (m <= small_cap) {                                  // uniform over the CTA
            uint32_t keys[ITEMS];
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                const int idx = idx0 + it * 32;
                keys[it] = idx < m ? ld_coh(kall + lo + idx) : 0xFFFFFFFFu;
            }
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int shift = 8 * pass;
#pragma unroll
                for (int q = 0; q < RX_RADIX / 32; ++q) whist[q * 32 + lane] = 0;
                __syncwarp();
                unsigned short ranks[ITEMS];
#pragma unroll
                for (int it = 0; it < ITEMS; ++it) {
                    const unsigned int digit = (keys[it] >> shift) & 0xFFu;
                    unsigned int differ = 0;
#pragma unroll
                    for (int bit = 0; bit < 8; ++bit) {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t.reg .b32 t, bal, ext;\n\t"
                            "and.b32 t, %1, %2;\n\t"
                            "setp.ne.u32 p, t, 0;\n\t"
                            "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
                            "selp.b32 ext, 0xffffffff, 0, p;\n\t"
                            "lop3.b32 %0, %0, bal, ext, 0xF6;\n\t}"
                            : "+r"(differ) : "r"(digit), "r"(1u << bit));
                    }
                    const unsigned int peers = ~differ;
                    const unsigned int lower = peers & lt_mask;
                    volatile unsigned int *ctr = whist + digit;
                    const unsigned int old = *ctr;
                    __syncwarp();
                    if (lower == 0u) *ctr = old + __popc(peers);
                    __syncwarp();
                    ranks[it] = (unsigned short)(old + __popc(lower));
                }
                __syncthreads();
                // per digit: offsets of each warp inside the digit run, then the run's start
                unsigned int tile_count = 0;
#pragma unroll
                for (int w = 0; w < RX_WARPS; ++w) {
                    const unsigned int c = s_whist[w * RX_RADIX + d];
                    s_whist[w * RX_RADIX + d] = tile_count;
                    tile_count += c;
                }
                const unsigned int incl = warp_incl_scan(tile_count);
                if (lane == 31) s_wsum[warp] = incl;
                __syncthreads();
                unsigned int wbase = 0;
#pragma unroll
                for (int w = 0; w < RX_WARPS; ++w) wbase += (w < warp) ? s_wsum[w] : 0u;
                const unsigned int doff = wbase + incl - tile_count;
#pragma unroll
                for (int w = 0; w < RX_WARPS; ++w) s_whist[w * RX_RADIX + d] += doff;
                __syncthreads();
#pragma unroll
                for (int it = 0; it < ITEMS; ++it) s_keys[whist[(keys[it] >> shift) & 0xFFu] + ranks[it]] = keys[it];
                __syncthreads();
                if (pass == 0) {
#pragma unroll
                    for (int it = 0; it < ITEMS; ++it) keys[it] = s_keys[idx0 + it * 32];
                    __syncthreads();   // everyone has reloaded before the next pass rewrites s_whist / s_keys
                }
            }
            // padding (all-ones keys) sorted last: positions >= m
            for (int idx = tid; idx < m; idx += RX_THREADS) kall[lo + idx] = IDENT ? s_keys[idx] : key_decode<uint32_t>(s_keys[idx], p.xf);
            __syncthreads();                                   // s_keys is rewritten by the next segment's first pass
        }
        lo = nlo; hi = nhi; seg = seg_n1; seg_n1 = seg_n2;
    }
}

// 16-bit bins.  One ticket = one entry of list16 = one whole segment of any length (a CTA's count loop walks it).
template <int NT, bool IDENT>
__global__ void __launch_bounds__(NT, 1) msd_leaf16_kernel(MsdLeafParams p) {
    constexpr int WORDS = MSD_BINS / 2;               // 32768 words, two 16-bit bins each
    constexpr int NW = NT / 32;
    constexpr int WORDS_PER_WARP = WORDS / NW;
    constexpr int CH = 16;                            // one CTA per SM: more loads in flight per thread
    constexpr int MAXH = 64;                          // heavy values per segment (each holds >= 1/64 of it)
    extern __shared__ __align__(16) unsigned int s_bins[];   // [WORDS] packed counters, all zero between segments
    __shared__ unsigned int s_wtot[NW];
    __shared__ unsigned int s_nheavy, s_spilled;
    __shared__ unsigned int s_tk[2];
    __shared__ uint4 s_heavy[MAXH + 1];               // {value, first position, copies}

    if (ld_cg_u32(&p.ctl->mode) == 0) return;
    const unsigned int n16 = ld_cg_u32(&p.ctl->n16);  // final: the 8-bit kernel has finished (stream order)
    if (n16 == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < WORDS / 4; i += NT) reinterpret_cast<uint4 *>(s_bins)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { s_nheavy = 0; s_spilled = 0; s_tk[0] = atomicAdd(&p.ctl->ticket_leaf, 1u); }
    unsigned int *const spill = p.spill + (size_t)blockIdx.x * MSD_BINS;
    for (int i = tid; i < MSD_BINS / 4; i += NT) reinterpret_cast<uint4 *>(spill)[i] = make_uint4(0, 0, 0, 0);   // workspace is uninitialised
    uint32_t *const kall = p.keys;
    __threadfence();
    __syncthreads();
    unsigned int next_t = 0;
    if (tid == 0 && s_tk[0] < n16) next_t = atomicAdd(&p.ctl->ticket_leaf, 1u);

    for (unsigned int it = 0;; ++it) {
        const unsigned int t = s_tk[it & 1];
        if (t >= n16) break;
        const unsigned int seg = ld_cg_u32(p.list16 + t);
        const unsigned int lo = __ldg(p.seg_start + seg), hi = __ldg(p.seg_start + seg + 1);
        const unsigned int m = hi - lo;
        const uint32_t *base = p.keys + lo;
        // ---- count.  Optimistic first: plain adds without return value; a 16-bit field that wraps loses 65535 or 65536
        // from the table's sum, so "sum != m" detects it and the segment is counted again carefully: the thread whose
        // add takes a field across 0x8000 moves 0x8000 counts to the CTA's spill table in global memory (exactly one
        // thread sees the crossing; the counts in flight are bounded by max(1024 threads x 16 loads, 32 warps
        // x 16 rows x 32) = 16384 < 0x8000, so a field never reaches 0x10000).  Rows of 32 equal keys are counted by lane 0 alone
        // (heavy duplicates would serialise 32 atomics on one word); row 0 is the probe.
        auto count_all = [&](auto careful_tag) {
            constexpr bool CAREFUL = decltype(careful_tag)::value;
            auto add = [&](uint32_t key, unsigned int inc) {
                const unsigned int sh = (key & 1u) << 4;
                unsigned int *word = &s_bins[(key & 0xFFFFu) >> 1];
                if constexpr (!CAREFUL) {
                    atomicAdd(word, inc << sh);
                } else {
                    const unsigned int old = (atomicAdd(word, inc << sh) >> sh) & 0xFFFFu;
                    if (old < 0x8000u && old + inc >= 0x8000u) {
                        atomicSub(word, 0x8000u << sh);
                        atomicAdd(spill + (key & 0xFFFFu), 0x8000u);
                        s_spilled = 1u;
                    }
                }
            };
            for (unsigned int i0 = 0; i0 < m; i0 += CH * NT) {
                uint32_t kk[CH];
                if (i0 + CH * NT <= m) {
#pragma unroll
                    for (int c = 0; c < CH; ++c) kk[c] = ld_coh(base + i0 + c * NT + tid);
                    const bool dup_rows = __all_sync(0xffffffffu, kk[0] == __shfl_sync(0xffffffffu, kk[0], 0));
                    if (dup_rows) {
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            if (__all_sync(0xffffffffu, kk[c] == __shfl_sync(0xffffffffu, kk[c], 0))) {
                                if (lane == 0) add(kk[c], 32u);
                            } else {
                                add(kk[c], 1u);
                            }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < CH; ++c) add(kk[c], 1u);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const unsigned int i = i0 + c * NT + tid;
                        kk[c] = i < m ? ld_coh(base + i) : 0u;
                    }
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        if (i0 + c * NT + tid < m) add(kk[c], 1u);
                }
            }
        };
        unsigned int *wbw = s_bins + warp * WORDS_PER_WARP;
        // warp totals: warp w owns table words [w * WORDS_PER_WARP, (w + 1) * WORDS_PER_WARP)
        auto warp_totals = [&](bool spilled) {
            unsigned int acc = 0;
            for (int j = lane; j < WORDS_PER_WARP; j += 32) {
                const unsigned int w = wbw[j];
                acc += (w & 0xFFFFu) + (w >> 16);
                if (spilled) {
                    const uint2 sp = __ldcg(reinterpret_cast<const uint2 *>(spill + 2 * (warp * WORDS_PER_WARP + j)));
                    acc += sp.x + sp.y;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s_wtot[warp] = acc;
        };
        count_all(std::false_type{});
        if (tid == 0) s_tk[(it + 1) & 1] = next_t;
        __syncthreads();
        const unsigned int tn = s_tk[(it + 1) & 1];
        if (tid == 0 && tn < n16) next_t = atomicAdd(&p.ctl->ticket_leaf, 1u);
        warp_totals(false);
        __syncthreads();
        unsigned int run = 0, total = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const unsigned int t2 = s_wtot[w];
            run += (w < warp) ? t2 : 0u;
            total += t2;
        }
        bool spilled = false;
        if (total != m) {                                         // uniform: some value has 65536 copies or more
            __syncthreads();                                      // everyone has read s_wtot
            for (int i = tid; i < WORDS / 4; i += NT) reinterpret_cast<uint4 *>(s_bins)[i] = make_uint4(0, 0, 0, 0);
            __syncthreads();
            count_all(std::true_type{});
            __threadfence();                                      // spill adds visible to the CTA's other threads
            __syncthreads();
            spilled = s_spilled != 0;
            warp_totals(spilled);
            __syncthreads();
            if (tid == 0) s_spilled = 0;                          // everyone holds `spilled`; the next careful count is barriers away
            run = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) run += (w < warp) ? s_wtot[w] : 0u;
        }
        const uint32_t hi16 = seg << 16;
        unsigned int heavy_th = m >> 6;
        if (heavy_th < MSD_LEAF_HEAVY) heavy_th = MSD_LEAF_HEAVY;
        // ---- expand: 32 words (64 values) per step and warp
        for (int j = lane; j < WORDS_PER_WARP; j += 32) {
            const unsigned int w = wbw[j];
            wbw[j] = 0;
            unsigned int c0 = w & 0xFFFFu, c1 = w >> 16;
            if (spilled) {
                uint2 *sp = reinterpret_cast<uint2 *>(spill + 2 * (warp * WORDS_PER_WARP + j));
                const uint2 v = __ldcg(sp);
                if (v.x | v.y) { c0 += v.x; c1 += v.y; *sp = make_uint2(0u, 0u); }
            }
            const unsigned int cnt = c0 + c1;
            const unsigned int incl = warp_incl_scan(cnt);
            const unsigned int tot = __shfl_sync(0xffffffffu, incl, 31);
            const unsigned int pos = lo + run + incl - cnt;                       // element index of this word's first key
            const uint32_t v0 = hi16 | (uint32_t)((warp * WORDS_PER_WARP + j) * 2);
            const uint32_t k0 = IDENT ? v0 : key_decode<uint32_t>(v0, p.xf);
            const uint32_t k1 = IDENT ? v0 + 1 : key_decode<uint32_t>(v0 + 1, p.xf);
            if (tot < 256u) {
                // few copies per value: every lane writes its own two runs — the first three copies as straight-line
                // predicated stores at precomputed addresses, the rest in a loop
                uint32_t *o0 = kall + pos, *o1 = kall + (pos + c0);
                st_if_gt<0, 0>(o0, k0, c0);
                st_if_gt<0, 0>(o1, k1, c1);
                st_if_gt<1, 1>(o0, k0, c0);
                st_if_gt<1, 1>(o1, k1, c1);
                st_if_gt<2, 2>(o0, k0, c0);
                st_if_gt<2, 2>(o1, k1, c1);
                if (max(c0, c1) > 3u) {
#pragma unroll 1
                    for (unsigned int q = 3; q < c0; ++q) o0[q] = k0;
#pragma unroll 1
                    for (unsigned int q = 3; q < c1; ++q) o1[q] = k1;
                }
            } else {
                // many copies per value (Float32 keys: 64 on average): the warp writes one lane's two adjacent runs at
                // a time with coalesced stores; runs of at least heavy_th keys (at most 64 per segment) are left to the
                // whole CTA
#pragma unroll 1
                for (int src = 0; src < 32; ++src) {
                    const unsigned int sc0 = __shfl_sync(0xffffffffu, c0, src), sc = __shfl_sync(0xffffffffu, cnt, src);
                    const unsigned int sp0 = __shfl_sync(0xffffffffu, pos, src);
                    const uint32_t sk0 = __shfl_sync(0xffffffffu, k0, src), sk1 = __shfl_sync(0xffffffffu, k1, src);
                    if (sc == 0) continue;                                       // uniform
                    if (max(sc0, sc - sc0) < heavy_th) {
                        for (unsigned int q = lane; q < sc; q += 32) kall[sp0 + q] = q < sc0 ? sk0 : sk1;
                    } else {
                        const unsigned int sc1 = sc - sc0;
                        if (sc0 >= heavy_th) { if (lane == 0) s_heavy[min(atomicAdd(&s_nheavy, 1u), (unsigned int)MAXH)] = make_uint4(sk0, sp0, sc0, 0u); }
                        else for (unsigned int q = lane; q < sc0; q += 32) kall[sp0 + q] = sk0;
                        if (sc1 >= heavy_th) { if (lane == 0) s_heavy[min(atomicAdd(&s_nheavy, 1u), (unsigned int)MAXH)] = make_uint4(sk1, sp0 + sc0, sc1, 0u); }
                        else for (unsigned int q = lane; q < sc1; q += 32) kall[sp0 + sc0 + q] = sk1;
                    }
                }
            }
            run += tot;
        }
        __syncthreads();
        {
            const unsigned int nh = min(s_nheavy, (unsigned int)MAXH);   // at most 64 values hold >= m / 64 keys each
            for (unsigned int e = 0; e < nh; ++e) {
                const uint4 h = s_heavy[e];
                for (unsigned int i = tid; i < h.z; i += NT) kall[h.y + i] = h.x;
            }
        }
        __syncthreads();            // table clean, heavy list consumed before the next segment
        if (tid == 0) s_nheavy = 0;
    }
}

}  // namespace b200
