// sort_msd.cuh — hybrid MSD radix sort for DENSE 32-bit keys (keys-only `sort!`), sm_100a.
//
// Why: an LSD radix pass must rank stably, and a stable warp rank costs ~37 issue slots per key row
// (8 ballots) — three such passes put `sort!` at ~290 instructions per key (profiles/r1_sort_opcode_mix.md).
// Keys-only sorts do not need stability when they go most-significant-digit first (equal keys are
// indistinguishable), so every rank can be ONE shared-memory atomic.  The path:
//
//   H   msd_hist16_kernel   one read of the keys -> histogram of the top 16 key bits (65536 bins, packed 16-bit
//                           shared-memory counters, flushed to global before a field can overflow)
//   PL  msd_plan_kernel     exclusive scan of the histogram = start of every 16-bit segment; list of non-empty
//                           segments; tile table of the 256 level-1 buckets; decides whether the data is dense
//                           enough (largest segment fits the leaf's shared memory, segments long enough for a
//                           counting leaf) — otherwise `mode` stays 0, every kernel below returns at once and the
//                           LSD passes of sort.cuh (gated on the same word) run instead
//   P1  msd_scatter_kernel<1>  partition by key byte 3 (atomic rank inside the tile, decoupled look-back over
//                           the tiles, digit runs written from shared memory) : keys -> alt
//   P2  msd_scatter_kernel<2>  the same inside each level-1 bucket by key byte 2 (tiles never straddle a bucket,
//                           the look-back chain restarts at a bucket's first tile)       : alt -> keys
//   L   msd_leaf_kernel     persistent CTAs take one 16-bit segment (<= CAP keys) at a time: partition it by byte 1
//                           into shared memory (atomic rank), then every warp finishes whole byte-1 groups with a
//                           256-counter counting sort on byte 0 straight into the segment's final place (in place)
//
// ~145 instructions per key in total instead of ~290, three passes over HBM plus the histogram read.
// Replaces (for this input class) CUDACore/src/sorting.jl:909-974 `bitonic_sort!`; keys keep Julia's `isless`
// order through key_encode / key_decode exactly as in the LSD path.
#pragma once
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

__device__ __forceinline__ unsigned int ld_cg_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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
    // iteration that started with a field above 2^14 is followed by a flush (some add saw bit 14 or 15 set).
    auto count = [&](uint32_t raw) {
        const uint32_t key = IDENT ? raw : key_encode<uint32_t>(raw, xf);
        const uint32_t inc = (key & 0x10000u) ? 0x10000u : 1u;
        const uint32_t old = atomicAdd(&sh16[key >> 17], inc);
        if (old & 0xC000C000u) s_flush = 1;
    };
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
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t v0 = c * CHUNK_V + tid;
        Vec16 r[MSD_H16_VPT];
#pragma unroll
        for (int k = 0; k < MSD_H16_VPT; ++k) {
            const int64_t v = v0 + (int64_t)k * MSD_H16_THREADS;
            if (v < nvec) r[k] = ldg_stream16(body + v * 4);
        }
#pragma unroll
        for (int k = 0; k < MSD_H16_VPT; ++k) {
            const int64_t v = v0 + (int64_t)k * MSD_H16_THREADS;
            if (v < nvec) {
#pragma unroll
                for (int j = 0; j < 4; ++j) count(r[k].w[j]);
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
    MsdCtl *ctl;
    unsigned int n;
    unsigned int tile2;             // keys per level-2 tile
    unsigned int cap_a, cap_b;      // leaf capacities (cap_b == 0: no second leaf)
    unsigned int min_avg;           // hybrid only when n / non-empty segments >= min_avg
};

__global__ void __launch_bounds__(1024, 1) msd_plan_kernel(MsdPlanParams p) {
    __shared__ unsigned int s_wsum[32], s_wne[32], s_wmax[32];
    __shared__ unsigned int s_excl[1024];
    __shared__ unsigned int s_t[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4 *h4 = reinterpret_cast<const uint4 *>(p.hist16) + tid * 16;   // 64 consecutive bins per thread
    unsigned int sum = 0, ne = 0, mx = 0;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const uint4 v = h4[k];
        sum += v.x + v.y + v.z + v.w;
        ne += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
        mx = max(max(mx, v.x), max(v.y, max(v.z, v.w)));
    }
    unsigned int isum = sum, ine = ne, imx = mx;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int a = __shfl_up_sync(0xffffffffu, isum, o);
        const unsigned int b = __shfl_up_sync(0xffffffffu, ine, o);
        const unsigned int c = __shfl_xor_sync(0xffffffffu, imx, o);
        if (lane >= o) { isum += a; ine += b; }
        imx = max(imx, c);
    }
    if (lane == 31) { s_wsum[warp] = isum; s_wne[warp] = ine; }
    if (lane == 0) s_wmax[warp] = imx;
    __syncthreads();
    unsigned int bsum = 0, bne = 0, tot_ne = 0, gmx = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
        if (w < warp) { bsum += s_wsum[w]; bne += s_wne[w]; }
        tot_ne += s_wne[w];
        gmx = max(gmx, s_wmax[w]);
    }
    unsigned int run = bsum + isum - sum;       // exclusive prefix of this thread's first bin
    unsigned int idx = bne + ine - ne;
    s_excl[tid] = run;
    uint4 *o4 = reinterpret_cast<uint4 *>(p.seg_start) + tid * 16;
    const unsigned int bin0 = (unsigned int)tid * 64;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const uint4 v = h4[k];
        uint4 o;
        o.x = run; if (v.x) p.seg_list[idx++] = bin0 + k * 4 + 0; run += v.x;
        o.y = run; if (v.y) p.seg_list[idx++] = bin0 + k * 4 + 1; run += v.y;
        o.z = run; if (v.z) p.seg_list[idx++] = bin0 + k * 4 + 2; run += v.z;
        o.w = run; if (v.w) p.seg_list[idx++] = bin0 + k * 4 + 3; run += v.w;
        o4[k] = o;
    }
    if (tid == 1023) p.seg_start[MSD_BINS] = run;
    __syncthreads();
    // level-1 bucket b = bins [256 b, 256 b + 256) = threads 4 b .. 4 b + 3
    unsigned int tiles = 0;
    if (tid < 256) {
        const unsigned int a = s_excl[tid * 4];
        const unsigned int b = tid == 255 ? p.n : s_excl[tid * 4 + 4];
        tiles = (b - a + p.tile2 - 1) / p.tile2;
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
        unsigned int mode = 0;
        if (gmx <= p.cap_a) mode = 1;
        else if (p.cap_b && gmx <= p.cap_b) mode = 2;
        if (tot_ne == 0 || (unsigned long long)p.n < (unsigned long long)p.min_avg * tot_ne) mode = 0;
        p.ctl->nne = tot_ne;
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
template <int LEVEL, int NT, int ITEMS, bool IDENT>
__global__ void __launch_bounds__(NT, 1024 / NT) msd_scatter_kernel(MsdScatterParams p) {
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
    __shared__ unsigned int s_seg[4];

    if (ld_cg_u32(&p.ctl->mode) == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool dthread = NT == 256 || tid < 256;
    const unsigned int tiles = LEVEL == 1 ? (p.n + TILE - 1) / TILE : ld_cg_u32(&p.ctl->tiles2);
    unsigned int *ticket = LEVEL == 1 ? &p.ctl->ticket1 : &p.ctl->ticket2;

    for (unsigned int it = 0;; ++it) {
        if (tid == 0) s_tk[it & 1] = atomicAdd(ticket, 1u);
        if (dthread) s_hist[tid] = 0;
        __syncthreads();
        const unsigned int t = s_tk[it & 1];
        if (t >= tiles) break;
        unsigned int seg = 0, first = 0, lo = 0, hi = p.n;
        if constexpr (LEVEL == 2) {
            if (dthread) {
                const unsigned int a = p.l1_tile_start[tid], b = p.l1_tile_start[tid + 1];
                if (a <= t && t < b) {
                    s_seg[0] = tid; s_seg[1] = a;
                    s_seg[2] = p.seg_start[tid * 256]; s_seg[3] = p.seg_start[tid * 256 + 256];
                }
            }
            __syncthreads();
            seg = s_seg[0]; first = s_seg[1]; lo = s_seg[2]; hi = s_seg[3];
        }
        const unsigned int tile_base = lo + (t - first) * TILE;
        const unsigned int remain = hi - tile_base;
        const int valid = remain > (unsigned int)TILE ? TILE : (int)remain;
        const bool full = valid == TILE;
        const bool chain_head = t == first;
        // first output index of this bucket's digit `tid` (loaded early, used after the look-back)
        unsigned int bin = 0;
        if (dthread) bin = LEVEL == 1 ? p.seg_start[tid * 256] : p.seg_start[seg * 256 + tid];

        // ---- load (coalesced, thread-strided: no order to preserve) and rank with one shared atomic per key
        const uint32_t *kin = p.keys_in + tile_base + tid;
        uint32_t keys[ITEMS];
        PackedRanks<ITEMS> ranks;
        if (full) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) keys[i] = ldg_stream(kin + i * NT);
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) keys[i] = (i * NT + tid < valid) ? ldg_stream(kin + i * NT) : 0u;
        }
        if constexpr (LEVEL == 1 && !IDENT) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) keys[i] = key_encode<uint32_t>(keys[i], p.xf);
        }
        if (full) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) ranks.set(i, atomicAdd(&s_hist[(keys[i] >> SHIFT) & 0xFFu], 1u));
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                unsigned int r = 0;
                if (i * NT + tid < valid) r = atomicAdd(&s_hist[(keys[i] >> SHIFT) & 0xFFu], 1u);
                ranks.set(i, r);
            }
        }
        __syncthreads();

        // ---- per digit: publish the tile's count, exclusive scan over the digits
        unsigned int cnt = 0;
        if (dthread) {
            cnt = s_hist[tid];
            st_relaxed<uint32_t>(p.status + (size_t)t * 256 + tid, (chain_head ? LB::INCLUSIVE : LB::PARTIAL) | cnt);
        }
        unsigned int incl = cnt;
        if (warp < 8) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) s_wsum[warp] = incl;
        }
        __syncthreads();
        unsigned int doff = 0;
        if (dthread) {
            unsigned int wbase = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) wbase += (w < warp) ? s_wsum[w] : 0u;
            doff = wbase + incl - cnt;
            s_doff[tid] = doff;
        }
        __syncthreads();

        // ---- scatter into shared memory, grouped by digit
        if (full) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) s_keys[s_doff[(keys[i] >> SHIFT) & 0xFFu] + ranks.get(i)] = keys[i];
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                if (i * NT + tid < valid) s_keys[s_doff[(keys[i] >> SHIFT) & 0xFFu] + ranks.get(i)] = keys[i];
        }

        // ---- decoupled look-back over this chain's earlier tiles
        if (dthread) {
            unsigned int excl = 0;
            if (!chain_head) {
                int u = (int)t - 1;
                const int floor_t = (int)first;
                bool done = false;
                while (!done) {
                    uint32_t w[RX_LOOK];
                    bool all_valid;
                    do {
                        all_valid = true;
#pragma unroll
                        for (int k = 0; k < RX_LOOK; ++k) {
                            w[k] = LB::INCLUSIVE;          // before the chain's first tile: inclusive, zero
                            if (u - k >= floor_t) w[k] = ld_relaxed<uint32_t>(p.status + (size_t)(u - k) * 256 + tid);
                            all_valid = all_valid && ((w[k] & LB::FLAGS) != 0);
                        }
                    } while (!all_valid);
#pragma unroll
                    for (int k = 0; k < RX_LOOK; ++k) {
                        if (!done) {
                            excl += w[k] & LB::MASK;
                            done = (w[k] & LB::FLAGS) == LB::INCLUSIVE;
                        }
                    }
                    u -= RX_LOOK;
                }
                st_relaxed<uint32_t>(p.status + (size_t)t * 256 + tid, LB::INCLUSIVE | (excl + cnt));
            }
            s_gbase[tid] = bin + excl - doff;
        }
        __syncthreads();

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
                if (i < valid) {
                    const uint32_t key = s_keys[i];
                    kout[s_gbase[(key >> SHIFT) & 0xFFu] + (unsigned int)i] = key;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ L: leaf, one 16-bit segment per CTA iteration
struct MsdLeafParams {
    uint32_t *keys;                   // encoded keys grouped by their top 16 bits; sorted + decoded in place
    const unsigned int *seg_start;    // [65537]
    const unsigned int *seg_list;     // non-empty segment ids
    MsdCtl *ctl;
    SortXform xf;
    unsigned int want_mode;
};

template <int NT, int ITEMS, bool IDENT>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) msd_leaf_kernel(MsdLeafParams p) {
    constexpr int CAP = NT * ITEMS;
    constexpr int NW = NT / 32;
    static_assert(CAP < 65536, "16-bit ranks");
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t *S = reinterpret_cast<uint32_t *>(s_raw);                 // [CAP] keys grouped by byte 1
    unsigned int *s_wc = reinterpret_cast<unsigned int *>(S + CAP);    // [NW][256] per-warp byte-0 counters
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned int s_gstart[256];
    __shared__ unsigned int s_gcnt[256];
    __shared__ unsigned int s_wsum[8];
    __shared__ unsigned int s_tk[2];

    if (ld_cg_u32(&p.ctl->mode) != p.want_mode) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int nne = ld_cg_u32(&p.ctl->nne);
    unsigned int *wc = s_wc + warp * 256;

    for (unsigned int it = 0;; ++it) {
        if (tid == 0) s_tk[it & 1] = atomicAdd(&p.ctl->ticket_leaf, 1u);
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        const unsigned int k = s_tk[it & 1];
        if (k >= nne) break;
        const unsigned int seg = p.seg_list[k];
        const unsigned int lo = p.seg_start[seg];
        const int m = (int)(p.seg_start[seg + 1] - lo);      // 1 <= m <= CAP (plan kernel)
        uint32_t *base = p.keys + lo;

        // ---- A: rank by byte 1 (one shared atomic per key); the keys are re-read from L2 in B
        PackedRanks<ITEMS> ranks;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (j * NT >= m) break;
            const int i = j * NT + tid;
            unsigned int r = 0;
            if (i < m) r = atomicAdd(&s_hist[(ld_coh(base + i) >> 8) & 0xFFu], 1u);
            ranks.set(j, r);
        }
        __syncthreads();
        unsigned int cnt = 0, incl = 0;
        if (tid < 256) cnt = s_hist[tid];
        incl = cnt;
        if (warp < 8) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (lane == 31) s_wsum[warp] = incl;
        }
        __syncthreads();
        if (tid < 256) {
            unsigned int wbase = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) wbase += (w < warp) ? s_wsum[w] : 0u;
            s_gstart[tid] = wbase + incl - cnt;
            s_gcnt[tid] = cnt;
        }
        __syncthreads();

        // ---- B: scatter into shared memory, grouped by byte 1
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            if (j * NT >= m) break;
            const int i = j * NT + tid;
            if (i < m) {
                const uint32_t key = ld_coh(base + i);
                S[s_gstart[(key >> 8) & 0xFFu] + ranks.get(j)] = key;
            }
        }
        __syncthreads();

        // ---- C: every warp finishes whole byte-1 groups: counting sort on byte 0 into the final place.
        // All keys of a group share their top 24 bits, so keys with equal byte 0 are identical: any order.
        for (int g = warp; g < 256; g += NW) {
            const int gc = (int)s_gcnt[g];
            if (gc == 0) continue;
            const unsigned int gs = s_gstart[g];
            if (gc == 1) {
                if (lane == 0) {
                    const uint32_t key = S[gs];
                    base[gs] = IDENT ? key : key_decode<uint32_t>(key, p.xf);
                }
                continue;
            }
            reinterpret_cast<uint4 *>(wc)[lane] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4 *>(wc)[lane + 32] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            for (int r = lane; r < gc; r += 32) atomicAdd(&wc[S[gs + r] & 0xFFu], 1u);
            __syncwarp();
            // lane owns counters 8 lane .. 8 lane + 7
            uint4 a = reinterpret_cast<uint4 *>(wc)[2 * lane], b = reinterpret_cast<uint4 *>(wc)[2 * lane + 1];
            const unsigned int tot = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
            unsigned int inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += up;
            }
            unsigned int run = inc - tot;
            uint4 oa, ob;
            oa.x = run; run += a.x; oa.y = run; run += a.y; oa.z = run; run += a.z; oa.w = run; run += a.w;
            ob.x = run; run += b.x; ob.y = run; run += b.y; ob.z = run; run += b.z; ob.w = run;
            reinterpret_cast<uint4 *>(wc)[2 * lane] = oa;
            reinterpret_cast<uint4 *>(wc)[2 * lane + 1] = ob;
            __syncwarp();
            for (int r = lane; r < gc; r += 32) {
                const uint32_t key = S[gs + r];
                const unsigned int pos = atomicAdd(&wc[key & 0xFFu], 1u);
                base[gs + pos] = IDENT ? key : key_decode<uint32_t>(key, p.xf);
            }
            __syncwarp();
        }
    }
}

}  // namespace b200
