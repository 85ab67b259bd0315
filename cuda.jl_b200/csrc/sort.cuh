// sort.cuh — onesweep LSD radix sort (8-bit digits) for sm_100a.
//
// Replaces the bitonic network of CUDACore/src/sorting.jl:525-975 (~240 full-array
// passes for 2^30 keys) with: one histogram pass over the keys, then
// one kernel per key byte in which every tile ranks its keys, publishes its digit
// counts, resolves its global offsets by decoupled look-back over the preceding
// tiles' counts, and scatters through shared memory so that global writes are
// contiguous per digit run.  Stable, so sortperm ties keep ascending index
// (sorting.jl:582-592) and `rev` is a key complement.
//
// Keys are mapped to unsigned integers whose order is Julia's `isless`
// (float sign twiddle, NaNs last; signed: sign-bit flip) when they are first
// read and mapped back when they are last written.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int RX_THREADS = 256;
constexpr int RX_WARPS = RX_THREADS / 32;
constexpr int RX_RADIX = 256;
#ifndef B200_RX_MATCH_MASK
#define B200_RX_MATCH_MASK 0x8888   // items 3, 7, 11, 15 of every 16
#endif
#ifndef B200_RX_LOOK
#define B200_RX_LOOK 4
#endif
constexpr int RX_LOOK = B200_RX_LOOK;   // predecessor tiles inspected per look-back round trip
// The ranking loop's warp-level ordering points: __syncwarp() costs an issue slot (SASS NOP) per call; the warp runs
// converged here (predicated store, no branch), so a compiler barrier is enough when B200_RX_NOSYNCWARP is set.
#ifdef B200_RX_NOSYNCWARP
#define RX_WARP_ORDER() asm volatile("" ::: "memory")
#else
#define RX_WARP_ORDER() __syncwarp()
#endif

template <int KB> struct KeyOf;
template <> struct KeyOf<1> { using type = uint8_t; };
template <> struct KeyOf<2> { using type = uint16_t; };
template <> struct KeyOf<4> { using type = uint32_t; };
template <> struct KeyOf<8> { using type = uint64_t; };

struct NoPayload {};

// Branch-free key transform, parameterised by uniform masks (built by make_xform in api_sort.cu):
//   encode: key = bits ^ (base_enc | (sar(bits) & fmn));  NaN (float only): key = nan_or | mag | sign;  key ^= desc
//   decode: key ^= desc;  bits = key ^ (base_dec ^ (sar(key) & fmn))
// float:    base_enc = sign, fmn = ~sign, base_dec = ~0   (negative: ~bits, positive: bits | sign)
// signed:   base_enc = sign, fmn = 0,     base_dec = sign
// unsigned: all zero
struct SortXform {
    uint64_t base_enc, base_dec, fmn;
    uint64_t mag_mask;      // ~sign for floats, 0 otherwise (so the NaN test never fires)
    uint64_t inf_bits;      // exponent mask of the float format
    uint64_t nan_or;        // ~0: every NaN maps to one key (sortperm); 0: keep the payload, drop the sign (sort!)
    uint64_t sign;
    uint64_t desc;          // ~0 when descending
};

template <class K> __device__ __forceinline__ K key_encode(K bits, const SortXform &x) {
    using SK = typename std::make_signed<K>::type;
    const K sar = (K)((SK)bits >> (8 * sizeof(K) - 1));
    K key = bits ^ ((K)x.base_enc | (sar & (K)x.fmn));
    const K mag = bits & (K)x.mag_mask;
    if (mag > (K)x.inf_bits) key = (K)x.nan_or | mag | (K)x.sign;
    return key ^ (K)x.desc;
}

template <class K> __device__ __forceinline__ K key_decode(K key, const SortXform &x) {
    using SK = typename std::make_signed<K>::type;
    key ^= (K)x.desc;
    const K sar = (K)((SK)key >> (8 * sizeof(K) - 1));
    return key ^ ((K)x.base_dec ^ (sar & (K)x.fmn));
}

// Device-side gate word (MsdCtl::mode, written by an earlier kernel of the same call): L2-coherent load.
__device__ __forceinline__ unsigned int ld_gate(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------ histogram of every digit
// One read of the keys builds all sizeof(K) digit histograms (shared-memory atomics,
// then one global atomic per (pass, bin) per CTA).
constexpr int RX_HIST_REPL = 8;   // histogram replicas (by lane & 7): spreads random digits over more banks

template <class K>
__global__ void __launch_bounds__(512) radix_hist_kernel(const K *__restrict__ keys, int64_t n, SortXform xf,
                                                         unsigned long long *__restrict__ ghist,
                                                         const unsigned int *__restrict__ gate) {
    constexpr int P = (int)sizeof(K);
    if (gate && ld_gate(gate) != 0) return;   // the hybrid MSD path (sort_msd.cuh) has taken this call
    constexpr int VN = 16 / (int)sizeof(K);
    extern __shared__ unsigned int sh[];   // [P][256][RX_HIST_REPL]
    for (int i = threadIdx.x; i < P * RX_RADIX * RX_HIST_REPL; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const bool aligned = (reinterpret_cast<uintptr_t>(keys) % 16) == 0;
    const int64_t nvec = aligned ? n / VN : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned int *mine = sh + (threadIdx.x & (RX_HIST_REPL - 1));
    auto count = [&](K raw) {
        const K key = key_encode<K>(raw, xf);
#pragma unroll
        for (int p = 0; p < P; ++p)
            atomicAdd(mine + (p * RX_RADIX + (int)((key >> (8 * p)) & 0xFF)) * RX_HIST_REPL, 1u);
    };
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        Vec16 raw = ldg_stream16(reinterpret_cast<const char *>(keys) + v * 16);
        K e[VN];
        memcpy(e, &raw, 16);
#pragma unroll
        for (int k = 0; k < VN; ++k) count(e[k]);
    }
    for (int64_t i = nvec * VN + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) count(keys[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < P * RX_RADIX; i += blockDim.x) {
        unsigned int c = 0;
#pragma unroll
        for (int r = 0; r < RX_HIST_REPL; ++r) c += sh[i * RX_HIST_REPL + r];
        if (c) atomicAdd(ghist + i, (unsigned long long)c);
    }
}

// Exclusive scan of each pass's 256-bin histogram: bins[p][d] = first output index of digit d.
static __global__ void __launch_bounds__(RX_RADIX) radix_bins_kernel(const unsigned long long *__restrict__ ghist,
                                                              unsigned long long *__restrict__ bins,
                                                              const unsigned int *__restrict__ gate) {
    __shared__ unsigned long long s[RX_RADIX];
    if (gate && ld_gate(gate) != 0) return;
    const int d = threadIdx.x;
    const unsigned long long c = ghist[blockIdx.x * RX_RADIX + d];
    s[d] = c;
    __syncthreads();
    for (int o = 1; o < RX_RADIX; o <<= 1) {
        unsigned long long v = d >= o ? s[d - o] : 0ull;
        __syncthreads();
        s[d] += v;
        __syncthreads();
    }
    bins[blockIdx.x * RX_RADIX + d] = s[d] - c;
}

// ------------------------------------------------------------------ partition points
// lower_bound of each needle in a sorted key array, in the ordered-key space (so it
// honours isless and `rev`).  One thread per needle; used by the multi-GPU sample sort
// to cut a locally sorted shard at the splitters.
template <class K>
__global__ void radix_lower_bound_kernel(const K *__restrict__ sorted, int64_t n, const K *__restrict__ needles,
                                         int64_t m, SortXform xf, long long *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const K needle = key_encode<K>(needles[j], xf);
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (key_encode<K>(sorted[mid], xf) < needle) lo = mid + 1;
        else hi = mid;
    }
    out[j] = lo;
}

// Device-side control block of one sort call (first 256 bytes of the workspace, zeroed by the call's one memset).
struct SortCtl {
    unsigned int mode;          // hybrid MSD path (sort_msd.cuh): 0 = declined (the LSD passes run); 1 / 2 = leaf A / B
    unsigned int nne;           // non-empty 16-bit segments
    unsigned int tiles2;        // level-2 tiles
    unsigned int max_seg;       // longest 16-bit segment (diagnostic)
    unsigned int heavy, small_cap;   // hybrid path: bit 0 / 1 = level 1 / level 2 has digits holding >= 1/8 of their chain (skewed keys);
                                     // small_cap: segments of up to this many keys are sorted on chip by msd_leaf_block_kernel (0: none)
    unsigned int ticket1, ticket2, ticket_leaf, n16;   // n16: entries of the 16-bit leaf's work list
    unsigned long long lsd_ticket[8];   // one tile ticket per LSD pass (persistent CTAs, no per-pass memset)
};

// ------------------------------------------------------------------ onesweep pass
struct OnesweepParams {
    const void *keys_in;
    void *keys_out;
    const void *vals_in;
    void *vals_out;
    int64_t n;
    int32_t shift;
    const unsigned long long *bins;   // [256] for this pass
    void *status;                     // LookT[tiles][256]
    unsigned long long *ticket;
    SortXform xf;
    int32_t encode_in;                // first pass: map raw keys to ordered unsigned keys
    int32_t decode_out;               // last pass: map back
    int32_t store_keys;               // 0 on sortperm's last pass
    int32_t gen_payload;              // first sortperm pass: payload = element index
    int32_t index_dtype;              // last sortperm pass: store index_base + payload as this dtype; -1: raw payload
    int64_t index_base;               // 1 for per-segment 1-based indices (+ segment offset when linear)
    int32_t unstable_ok;              // keys-only FIRST pass: equal keys are indistinguishable and no earlier pass has
                                      // established an order, so ranks inside a digit may come from shared-memory atomics
    int64_t tiles;                    // tiles of this pass; CTAs are persistent and take tiles by ticket
    const unsigned int *gate;         // non-null: run only while *gate == 0 (hybrid MSD path declined the call)
    // Look-back flag codes of THIS pass (already shifted into the top two bits).  The status array is zeroed once per
    // sort call; pass p uses inclusive = 1 + p % 3, partial = 1 + (p + 1) % 3, and what the previous pass left behind
    // (every tile ends "inclusive", code 1 + (p - 1) % 3) reads as "not yet published" — no memset between passes.
    unsigned long long flag_incl, flag_part;
};

template <class LookT> struct LookBits;
template <> struct LookBits<uint32_t> {
    static constexpr uint32_t PARTIAL = 1u << 30, INCLUSIVE = 2u << 30, MASK = (1u << 30) - 1, FLAGS = 3u << 30;
};
template <> struct LookBits<uint64_t> {
    static constexpr uint64_t PARTIAL = 1ull << 62, INCLUSIVE = 2ull << 62, MASK = (1ull << 62) - 1, FLAGS = 3ull << 62;
};

template <class LookT> __device__ __forceinline__ LookT ld_relaxed(const LookT *p) {
    LookT v;
    if constexpr (sizeof(LookT) == 4) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <class LookT> __device__ __forceinline__ void st_relaxed(LookT *p, LookT v) {
    if constexpr (sizeof(LookT) == 4) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
    else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// 8-bit digit at a byte-aligned shift: one PRMT for keys of up to 32 bits.
template <class K> __device__ __forceinline__ unsigned int digit_of(K key, int shift) {
    if constexpr (sizeof(K) <= 4) return __byte_perm((unsigned int)key, 0u, 0x4440u | (unsigned int)(shift >> 3));
    else return (unsigned int)(key >> shift) & 0xFFu;
}

template <class V> struct PayTraits { static constexpr bool HAS = true; using type = V; };
template <> struct PayTraits<NoPayload> { static constexpr bool HAS = false; using type = uint32_t; };

// ARANK (keys-only FIRST pass, p.unstable_ok): ranks inside a digit come from shared-memory atomics.  A separate
// instantiation: with both ranking loops in one kernel it needs 76 registers instead of 63 (3 CTAs per SM, -9 %).
// minimum CTAs per SM requested for the atomic-ranked instantiation (0 = compiler default = 64 registers, 4 CTAs).
// NB: an explicit 1 lets ptxas use 115 registers for the ballot-ranked kernel (measured: 16.4 -> 20.2 ms), hence 0.
#ifndef B200_RX_ARANK_CTAS
#define B200_RX_ARANK_CTAS 0
#endif
// NT = threads per CTA (256, or 512 as an experiment: the per-digit bookkeeping and the look-back are done by 256
// threads per TILE, so a 512-thread tile halves their cost per key; B200_SORT_NT=512, keys-only, not yet measured).
// PERSIST (only instantiated for the gated fallback of the hybrid MSD path, where a grid of `tiles` CTAs that all
// return at once would cost more than the sort): persistent CTAs, registers capped so that 4 CTAs stay resident.
// Measured on B200, sort! 2^30 UInt32: one CTA per tile 15.7 ms, persistent 17.8 ms — so everything else keeps the
// one-tile-per-CTA form, whose SASS is the round-1 kernel plus the flag-code compares.
template <class K, class V, class LookT, int ITEMS, bool ARANK = false, int NT = RX_THREADS, bool PERSIST = false>
__global__ void __launch_bounds__(NT, ((PERSIST || sizeof(K) + (PayTraits<V>::HAS ? sizeof(V) : 0) <= 8) ? 1024 / NT : 0))
radix_onesweep_kernel(OnesweepParams p) {
    using LB = LookBits<LookT>;
    constexpr bool HAS_V = PayTraits<V>::HAS;
    using VT = typename PayTraits<V>::type;
    constexpr int TILE = NT * ITEMS;
    constexpr int WARP_ITEMS = 32 * ITEMS;
    constexpr int NWARPS = NT / 32;
    constexpr int DWARPS = RX_RADIX / 32;        // warps whose threads own one digit each
    static_assert(NT >= RX_RADIX && NT % 32 == 0, "one thread per digit");

    extern __shared__ __align__(16) unsigned char s_raw[];
    // layout: [warp hist 8*256 u32][digit offset 256 u32][global base 256 i64][keys TILE][vals TILE]
    unsigned int *s_whist = reinterpret_cast<unsigned int *>(s_raw);
    unsigned int *s_doff = s_whist + NWARPS * RX_RADIX;
    // global index of the tile's first key of each digit, minus its shared-memory position;
    // 32-bit when n < 2^30 (LookT == uint32_t), which saves 64-bit address arithmetic per key
    using BaseT = typename std::conditional<sizeof(LookT) == 4, int, long long>::type;
    long long *s_gbase_raw = reinterpret_cast<long long *>(s_doff + RX_RADIX);
    BaseT *s_gbase = reinterpret_cast<BaseT *>(s_gbase_raw);
    K *s_keys = reinterpret_cast<K *>(s_gbase_raw + RX_RADIX);
    VT *s_vals = reinterpret_cast<VT *>(reinterpret_cast<unsigned char *>(s_keys) + ((TILE * sizeof(K) + 15) / 16) * 16);
    __shared__ unsigned long long s_ticket[2];
    __shared__ unsigned int s_wsum[DWARPS];

    if (p.gate && ld_gate(p.gate) != 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool dthread = NT == RX_RADIX || tid < RX_RADIX;   // this thread owns digit `tid`
    const LookT flag_incl = (LookT)p.flag_incl, flag_part = (LookT)p.flag_part;
    unsigned int *whist = s_whist + warp * RX_RADIX;
    // PERSIST: the CTA loops, taking tiles by ticket (two ticket slots: the next iteration's write cannot overtake
    // a slow reader of this one, because a barrier separates them).  !PERSIST: one tile per CTA, grid == tiles.
  for (unsigned int iter = 0;; ++iter) {
    if (tid == 0) s_ticket[iter & 1] = atomicAdd(p.ticket, 1ull);
    // zero this warp's digit counters while the ticket is in flight
#pragma unroll
    for (int k = 0; k < RX_RADIX / 32; ++k) whist[k * 32 + lane] = 0;
    __syncthreads();
    const int64_t t = (int64_t)s_ticket[iter & 1];
    if (PERSIST && t >= p.tiles) return;
    const int64_t tile_base = t * TILE;
    const int64_t remain = p.n - tile_base;
    const int valid = remain > TILE ? TILE : (int)remain;
    const bool full = valid == TILE;

    // ---- load keys, warp-striped: item i of lane l is element warp*WARP_ITEMS + i*32 + l
    const K *kin = static_cast<const K *>(p.keys_in) + tile_base + warp * WARP_ITEMS + lane;
    const int idx0 = warp * WARP_ITEMS + lane;
    K keys[ITEMS];
    if (full) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) keys[i] = ldg_stream(kin + i * 32);
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) keys[i] = (idx0 + i * 32 < valid) ? ldg_stream(kin + i * 32) : (K)0;
    }
    if (p.encode_in) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) keys[i] = key_encode<K>(keys[i], p.xf);
    }
    if (!full) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (idx0 + i * 32 >= valid) keys[i] = (K)~(K)0;   // padding ranks last inside digit 255
    }

    // ---- early counts: a block-level digit histogram (shared-memory atomics, LSU pipe) is
    // published BEFORE the expensive ranking, so that by the time any tile reaches its look-back
    // the counts of its predecessors have long been visible (they only waited for load + count,
    // not for load + rank): this removes the convoy in which all resident tiles finish ranking
    // together and then sit in the look-back.
    LookT *status = static_cast<LookT *>(p.status);
    // measured on B200 (2^30 keys): pairs 26.7 -> 25.0 ms, keys-only 20.1 -> 20.6 ms, so only pairs use it
    constexpr bool EARLY_COUNTS = HAS_V;
    if constexpr (EARLY_COUNTS) {
        unsigned int *s_bhist = s_doff;                 // reused below as the digit-offset table
        if (dthread) s_bhist[tid] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) atomicAdd(&s_bhist[digit_of<K>(keys[i], p.shift)], 1u);
        __syncthreads();
        if (dthread) {
            unsigned int c = s_bhist[tid];
            if (tid == RX_RADIX - 1) c -= (unsigned int)(TILE - valid);
            st_relaxed<LookT>(status + t * RX_RADIX + tid, (LookT)((t > 0 ? flag_part : flag_incl) | (LookT)c));
        }
    }

    // ---- rank inside the warp: stable by (item, lane) = element order.
    // Peer masks come from 8 ballots (one per digit bit) rather than MATCH.ANY: MATCH runs on
    // the ADU pipe at ~1 warp instruction per 64 cycles per SM on sm_100 and capped a pass at
    // ~140 Gkeys/s (ncu: sm__inst_executed_pipe_adu 78%); VOTE issues at 2 per cycle per SM.
    unsigned short ranks[ITEMS];
    const unsigned int lt_mask = (1u << lane) - 1u;
    if constexpr (ARANK) {
        // one ATOMS per key (LSU pipe) instead of ~37 ALU/ADU instructions: the returned old count is a unique rank
        // among this warp's keys of the digit, and the counter ends up as the warp's digit count like below
        if (full) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                ranks[i] = (unsigned short)atomicAdd(whist + digit_of<K>(keys[i], p.shift), 1u);
        } else {
            // ragged tile: the padding (all in digit 255, highest element indices) must rank after every real key;
            // per item the real lanes' atomics are issued before the padding lanes' (same warp: program order)
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const bool real = idx0 + i * 32 < valid;
                unsigned int r = 0;
                if (real) r = atomicAdd(whist + digit_of<K>(keys[i], p.shift), 1u);
                __syncwarp();
                if (!real) r = atomicAdd(whist + RX_RADIX - 1, 1u);
                ranks[i] = (unsigned short)r;
            }
        }
    } else
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const unsigned int digit = digit_of<K>(keys[i], p.shift);
        unsigned int peers;
        if ((B200_RX_MATCH_MASK >> (i & 15)) & 1) {
            // a quarter of the items go through MATCH.ANY so that the ADU pipe shares the load with the ALU
            peers = __match_any_sync(0xffffffffu, digit);
        } else {
            unsigned int differ = 0;   // lanes whose digit differs from mine in some bit
#pragma unroll
            for (int bit = 0; bit < 8; ++bit) {
                // and+setp -> one LOP3 with predicate output; VOTE; SEL; LOP3 (differ | (bal ^ ext))
                asm volatile(
                    "{\n\t.reg .pred p;\n\t.reg .b32 t, bal, ext;\n\t"
                    "and.b32 t, %1, %2;\n\t"
                    "setp.ne.u32 p, t, 0;\n\t"
                    "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
                    "selp.b32 ext, 0xffffffff, 0, p;\n\t"
                    "lop3.b32 %0, %0, bal, ext, 0xF6;\n\t}"
                    : "+r"(differ) : "r"(digit), "r"(1u << bit));
            }
            peers = ~differ;
        }
        // every peer reads the digit's counter, then the lowest peer lane bumps it (same warp:
        // shared-memory accesses execute in program order; __syncwarp keeps the compiler honest)
        const unsigned int lower = peers & lt_mask;
        const unsigned int below = __popc(lower);
        volatile unsigned int *ctr = whist + digit;
        const unsigned int old = *ctr;
        RX_WARP_ORDER();
        if (lower == 0u) *ctr = old + __popc(peers);
        RX_WARP_ORDER();
        ranks[i] = (unsigned short)(old + below);
    }
    __syncthreads();

    // ---- per digit (thread d): offsets of each warp inside the digit run, tile count
    const int d = dthread ? tid : 0;
    unsigned int tile_count = 0;
    if (dthread) {
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) {
            const unsigned int c = s_whist[w * RX_RADIX + d];
            s_whist[w * RX_RADIX + d] = tile_count;
            tile_count += c;
        }
    }
    const unsigned int pad = (unsigned int)(TILE - valid);
    const unsigned int real_count = (d == RX_RADIX - 1) ? tile_count - pad : tile_count;
    if constexpr (!EARLY_COUNTS) {  // keys-only: publish here; tile 0 knows its inclusive prefix already
        if (dthread)
            st_relaxed<LookT>(status + t * RX_RADIX + d, (LookT)((t > 0 ? flag_part : flag_incl) | (LookT)real_count));
    }

    // exclusive scan of tile_count over the 256 digits -> start of each digit run in shared memory
    unsigned int incl = tile_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31 && dthread) s_wsum[warp] = incl;
    __syncthreads();
    unsigned int wbase = 0;
#pragma unroll
    for (int w = 0; w < DWARPS; ++w) wbase += (w < warp) ? s_wsum[w] : 0u;
    const unsigned int doff = wbase + incl - tile_count;
    // fold the digit run's start into every warp's offset: the scatter then needs one lookup
    if (dthread) {
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) s_whist[w * RX_RADIX + d] += doff;
    }
    __syncthreads();

    // ---- scatter keys (and payload) into shared memory, sorted by digit
    VT vals[HAS_V ? ITEMS : 1];
    if constexpr (HAS_V) {
        const VT *vin = static_cast<const VT *>(p.vals_in) + tile_base;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int idx = idx0 + i * 32;
            if (p.gen_payload) vals[i] = (VT)(tile_base + idx);
            else vals[i] = idx < valid ? ldg_stream(vin + idx) : (VT)0;
        }
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const unsigned int digit = digit_of<K>(keys[i], p.shift);
        const unsigned int pos = whist[digit] + ranks[i];
        s_keys[pos] = keys[i];
        if constexpr (HAS_V) s_vals[pos] = vals[i];
    }

    // ---- decoupled look-back for digit d
    unsigned long long excl = 0;
    if (t > 0 && dthread) {
        int64_t u = t - 1;
        bool done = false;
        while (!done) {
            LookT w[RX_LOOK];
            bool all_valid;
            do {
                all_valid = true;
#pragma unroll
                for (int k = 0; k < RX_LOOK; ++k) {
                    w[k] = flag_incl;   // virtual tiles before the first one: inclusive, zero
                    if (u - k >= 0) w[k] = ld_relaxed<LookT>(status + (u - k) * RX_RADIX + d);
                    const LookT f = w[k] & LB::FLAGS;
                    all_valid = all_valid && (f == flag_incl || f == flag_part);
                }
            } while (!all_valid);
#pragma unroll
            for (int k = 0; k < RX_LOOK; ++k) {
                if (!done) {
                    excl += (unsigned long long)(w[k] & LB::MASK);
                    done = (w[k] & LB::FLAGS) == flag_incl;
                }
            }
            u -= RX_LOOK;
        }
        st_relaxed<LookT>(status + t * RX_RADIX + d, (LookT)(flag_incl | (LookT)(excl + real_count)));
    }
    if (dthread) s_gbase[d] = (BaseT)((long long)(p.bins[d] + excl) - (long long)doff);
    __syncthreads();

    // ---- write out: consecutive threads -> consecutive addresses inside each digit run.
    // The uniform choices (full tile? decode? keys? index conversion?) are hoisted out of
    // the per-key loop: the common case is load-shared, one table lookup, one store.
    K *kout = static_cast<K *>(p.keys_out);
    auto emit = [&](auto full_c, auto mode_c) {
        constexpr bool kFull = decltype(full_c)::value;
        constexpr int kMode = decltype(mode_c)::value;   // 0: keys as is, 1: keys decoded, 2: no keys
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = tid + k * NT;
            if (kFull || i < valid) {
                const K key = s_keys[i];
                const BaseT dst = s_gbase[digit_of<K>(key, p.shift)] + i;
                if constexpr (kMode == 0) kout[dst] = key;
                if constexpr (kMode == 1) kout[dst] = key_decode<K>(key, p.xf);
                if constexpr (HAS_V) {
                    const VT v = s_vals[i];
                    if (p.index_dtype < 0) {
                        static_cast<VT *>(p.vals_out)[dst] = v;
                    } else if (p.index_dtype == B200_I64 || p.index_dtype == B200_U64) {
                        static_cast<long long *>(p.vals_out)[dst] = (long long)v + p.index_base;
                    } else {
                        static_cast<int *>(p.vals_out)[dst] = (int)((long long)v + p.index_base);
                    }
                }
            }
        }
    };
    auto emit_mode = [&](auto full_c) {
        if (!p.store_keys) emit(full_c, std::integral_constant<int, 2>{});
        else if (p.decode_out) emit(full_c, std::integral_constant<int, 1>{});
        else emit(full_c, std::integral_constant<int, 0>{});
    };
    if (full) emit_mode(std::true_type{});
    else emit_mode(std::false_type{});
    if constexpr (!PERSIST) return;
  }   // persistent tile loop
}

}  // namespace b200
