// sort_ctr.cuh — onesweep pass with thread-private digit counters (keys-only, 32-bit keys).
//
// The ballot/MATCH ranking of sort.cuh spends ~45 ALU instructions per key; the pass is
// issue-bound at ~270 Gkeys/s.  With 227 KB of shared memory per SM a different classic
// becomes possible for full 8-bit digits: every thread owns a CONTIGUOUS run of ITEMS keys
// and a private column of 256 16-bit counters (two digits per 32-bit word, word index =
// row * NT + thread, so a warp's counter accesses never bank-conflict whatever the digits).
//   count : per key  PRMT (digit -> byte offset) + LDS.U16 + IADD + STS.U16; the old value
//           (the key's rank among the thread's earlier keys of that digit) is kept in a byte
//   scan  : the [128 rows][NT] matrix is prefix-summed along the thread axis (packed 2x16-bit
//           adds, one warp shuffle scan per row); row totals give the tile's digit counts
//   rank  : position = digit start + prefix(digit, thread) + remembered byte  -> shared staging
//   then the same decoupled look-back and digit-run write-out as sort.cuh.
// Tile order = thread-major, item-minor = element order, so the pass stays stable.
#pragma once
#include "sort.cuh"

namespace b200 {

constexpr int CT_REP = 4;        // replicas of the digit-start table (by lane & 3): random digits, fewer bank conflicts
constexpr int CT_ROWS = 128;     // digit d lives in row d & 127, half d >> 7

__device__ __forceinline__ unsigned int lds_u16(unsigned int a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(unsigned int a, unsigned int v) {
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ unsigned int lds_u32(unsigned int a) {
    unsigned int v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned int a, unsigned int v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

// p1 = (digit << 8) | (digit bit 7 replicated over the low byte): one PRMT
// (inline PTX: __byte_perm masks the selector to 3 bits per nibble and would drop the sign-replicate flag)
__device__ __forceinline__ unsigned int ctr_p1(unsigned int key, unsigned int sel) {
    unsigned int r;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(key), "r"(sel));
    return r;
}
// byte offset of the counter of p1's digit inside a thread's column
template <int NT> __device__ __forceinline__ unsigned int ctr_off(unsigned int p1) {
    if constexpr (NT == 64) return p1 & 0x7F02u;        // row * 256 + half * 2
    else return (p1 << 1) & 0xFE02u;                    // row * 512 + half * 2
}

template <int NT, int ITEMS> constexpr size_t ctr_smem_bytes() {
    return (size_t)CT_ROWS * NT * 4 + (size_t)NT * ITEMS * 4 + 256 * CT_REP * 4 + CT_ROWS * 4 + 256 * 4 + 256 * 8;
}

template <class LookT, int NT, int ITEMS>
__global__ void __launch_bounds__(NT) radix_onesweep_ctr_kernel(OnesweepParams p) {
    using K = uint32_t;
    using LB = LookBits<LookT>;
    using BaseT = typename std::conditional<sizeof(LookT) == 4, int, long long>::type;
    constexpr int TILE = NT * ITEMS;
    constexpr int NW = NT / 32;
    constexpr int DPT = 256 / NT;                 // digits per thread
    static_assert(NT == 64 || NT == 128, "counter layout is written for 64 or 128 threads");
    static_assert(ITEMS % 8 == 0 && ITEMS <= 128, "ITEMS: multiple of 8 (LDG.256), remembered ranks are bytes");
    static_assert(TILE <= 16384, "packed 16-bit prefixes");

    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned int *s_ctr = reinterpret_cast<unsigned int *>(s_raw);        // [128][NT] packed counters / prefixes
    K *s_keys = reinterpret_cast<K *>(s_ctr + CT_ROWS * NT);              // [TILE] digit-sorted staging
    unsigned int *s_doff = reinterpret_cast<unsigned int *>(s_keys + TILE);   // [256][CT_REP] digit start in the tile
    unsigned int *s_tot = s_doff + 256 * CT_REP;                          // [128] packed row totals
    unsigned int *s_dst = s_tot + CT_ROWS;                                // [256] digit start (one copy)
    long long *s_gbase_raw = reinterpret_cast<long long *>(s_dst + 256);
    BaseT *s_gbase = reinterpret_cast<BaseT *>(s_gbase_raw);              // [256] global index - staging position
    __shared__ unsigned long long s_ticket;
    __shared__ unsigned int s_wsum[NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1ull);
    __syncthreads();
    const int64_t t = (int64_t)s_ticket;
    const int64_t tile_base = t * TILE;
    const int64_t remain = p.n - tile_base;
    const int valid = remain > TILE ? TILE : (int)remain;
    const bool full = valid == TILE;

    // ---- load: thread `tid` owns elements [tid*ITEMS, (tid+1)*ITEMS) of the tile (full 32-byte sectors per lane)
    const K *kin = static_cast<const K *>(p.keys_in) + tile_base + tid * ITEMS;
    K keys[ITEMS];
    if (full && (reinterpret_cast<uintptr_t>(p.keys_in) % 32) == 0) {
#pragma unroll
        for (int j = 0; j < ITEMS / 8; ++j) {
            const Vec32 v = ldg_stream32(kin + 8 * j);
#pragma unroll
            for (int e = 0; e < 8; ++e) keys[8 * j + e] = v.w[e];
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) keys[i] = (tid * ITEMS + i < valid) ? ldg_stream(kin + i) : (K)0;
    }
    // zero the counters while the keys are in flight
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        uint4 *c4 = reinterpret_cast<uint4 *>(s_ctr);
#pragma unroll
        for (int k = 0; k < CT_ROWS / 4; ++k) c4[k * NT + tid] = z;
    }
    __syncthreads();
    if (p.encode_in) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) keys[i] = key_encode<K>(keys[i], p.xf);
    }
    if (!full) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (tid * ITEMS + i >= valid) keys[i] = ~(K)0;     // padding: end of the tile, ranks last inside digit 255
    }

    // ---- count: private counters, remembered old value = rank among this thread's earlier keys of the digit
    const unsigned int b = (unsigned int)p.shift >> 3;
    const unsigned int sel = 0x4400u | (b << 4) | (8u | b);
    const unsigned int tbase = smem_u32(s_ctr) + (unsigned int)tid * 4u;
    unsigned int memo[ITEMS / 4];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const unsigned int a = tbase + ctr_off<NT>(ctr_p1(keys[i], sel));
        const unsigned int c = lds_u16(a);
        sts_u16(a, c + 1u);
        if ((i & 3) == 0) memo[i >> 2] = c;
        else memo[i >> 2] |= c << (8 * (i & 3));
    }
    __syncthreads();

    // ---- scan the counter matrix along the thread axis (both packed digits at once): one thread per row,
    // sequential packed adds (no shuffles: the shuffle version was half of the kernel's time).  All lanes of a
    // warp would hit the same bank group at the same chunk index, so lane l walks its row starting at chunk l
    // (mod CH): sweep 1 finds the sum W of the chunks before the start, sweep 2 writes the exclusive prefixes.
    {
        constexpr int CH = NT / 4;                // 16-byte chunks per row
        constexpr int BATCH = 8;
        const int s0 = lane & (CH - 1);
#pragma unroll 1
        for (int row = tid; row < CT_ROWS; row += NT) {
            uint4 *rowp = reinterpret_cast<uint4 *>(s_ctr + row * NT);
            unsigned int W = 0, T = 0;
#pragma unroll
            for (int i0 = 0; i0 < CH; i0 += BATCH) {
                uint4 w[BATCH];
#pragma unroll
                for (int k = 0; k < BATCH; ++k) w[k] = rowp[(i0 + k + s0) & (CH - 1)];
#pragma unroll
                for (int k = 0; k < BATCH; ++k) {
                    const unsigned int sum = (w[k].x + w[k].y) + (w[k].z + w[k].w);
                    T += sum;
                    if (i0 + k >= CH - s0) W += sum;        // chunk index wrapped: it lies before the start
                }
            }
            unsigned int R = W;
#pragma unroll
            for (int i0 = 0; i0 < CH; i0 += BATCH) {
                uint4 w[BATCH];
#pragma unroll
                for (int k = 0; k < BATCH; ++k) w[k] = rowp[(i0 + k + s0) & (CH - 1)];
#pragma unroll
                for (int k = 0; k < BATCH; ++k) {
                    const int c = (i0 + k + s0) & (CH - 1);
                    if (c == 0) R = 0;
                    uint4 e;
                    e.x = R; e.y = e.x + w[k].x; e.z = e.y + w[k].y; e.w = e.z + w[k].z;
                    R = e.w + w[k].w;
                    rowp[c] = e;
                }
            }
            s_tot[row] = T;
        }
    }
    __syncthreads();

    // ---- publish this tile's digit counts (thread owns digits tid, tid+NT, ...: coalesced status words)
    LookT *status = static_cast<LookT *>(p.status);
    const unsigned int pad = (unsigned int)(TILE - valid);
    unsigned int real_count[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int d = tid + j * NT;
        const unsigned int tt = s_tot[d & 127];
        unsigned int c = (d & 128) ? (tt >> 16) : (tt & 0xFFFFu);
        if (d == 255) c -= pad;
        real_count[j] = c;
        st_relaxed<LookT>(status + t * RX_RADIX + d, (LookT)((t > 0 ? LB::PARTIAL : LB::INCLUSIVE) | (LookT)c));
    }

    // ---- digit starts inside the tile: thread owns the consecutive digits tid*DPT .. tid*DPT+DPT-1
    {
        unsigned int cnt[DPT], sum = 0;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = tid * DPT + j;
            const unsigned int tt = s_tot[d & 127];
            cnt[j] = (d & 128) ? (tt >> 16) : (tt & 0xFFFFu);
            sum += cnt[j];
        }
        unsigned int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        unsigned int run = incl - sum;
#pragma unroll
        for (int w = 0; w < NW; ++w) run += (w < warp) ? s_wsum[w] : 0u;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = tid * DPT + j;
            s_dst[d] = run;
            *reinterpret_cast<uint4 *>(s_doff + d * CT_REP) = make_uint4(run, run, run, run);
            run += cnt[j];
        }
    }
    __syncthreads();

    // ---- rank + scatter into the digit-sorted staging buffer (loads of a batch issued together)
    {
        const unsigned int kbase = smem_u32(s_keys);
        const unsigned int dbase = smem_u32(s_doff) + (unsigned int)(lane & (CT_REP - 1)) * 4u;
        constexpr int RB = 8;
#pragma unroll
        for (int i0 = 0; i0 < ITEMS; i0 += RB) {
            unsigned int c0[RB], dof[RB];
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const unsigned int p1 = ctr_p1(keys[i0 + k], sel);
                c0[k] = lds_u16(tbase + ctr_off<NT>(p1));
                dof[k] = lds_u32(dbase + ((p1 >> 4) & 0xFF0u));      // digit * CT_REP * 4
            }
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int i = i0 + k;
                const unsigned int pos = c0[k] + dof[k] + ((memo[i >> 2] >> (8 * (i & 3))) & 0xFFu);
                sts_u32(kbase + pos * 4u, keys[i]);
            }
        }
    }

    // ---- decoupled look-back for this thread's DPT digits, all in flight together
    unsigned long long excl[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) excl[j] = 0;
    if (t > 0) {
        bool done[DPT];
#pragma unroll
        for (int j = 0; j < DPT; ++j) done[j] = false;
        int64_t u = t - 1;
        bool all_done = false;
        while (!all_done) {
            LookT w[DPT][RX_LOOK];
            bool all_valid;
            do {
                all_valid = true;
#pragma unroll
                for (int j = 0; j < DPT; ++j) {
#pragma unroll
                    for (int k = 0; k < RX_LOOK; ++k) {
                        w[j][k] = LB::INCLUSIVE;           // finished digits / virtual tiles before the first: inclusive zero
                        if (!done[j] && u - k >= 0) w[j][k] = ld_relaxed<LookT>(status + (u - k) * RX_RADIX + tid + j * NT);
                        all_valid = all_valid && ((w[j][k] & LB::FLAGS) != 0);
                    }
                }
            } while (!all_valid);
            all_done = true;
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
#pragma unroll
                for (int k = 0; k < RX_LOOK; ++k) {
                    if (!done[j]) {
                        excl[j] += (unsigned long long)(w[j][k] & LB::MASK);
                        done[j] = (w[j][k] & LB::FLAGS) == LB::INCLUSIVE;
                    }
                }
                all_done = all_done && done[j];
            }
            u -= RX_LOOK;
        }
#pragma unroll
        for (int j = 0; j < DPT; ++j)
            st_relaxed<LookT>(status + t * RX_RADIX + tid + j * NT, (LookT)(LB::INCLUSIVE | (LookT)(excl[j] + real_count[j])));
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int d = tid + j * NT;
        s_gbase[d] = (BaseT)((long long)(p.bins[d] + excl[j]) - (long long)s_dst[d]);
    }
    __syncthreads();

    // ---- write out: consecutive threads -> consecutive addresses inside each digit run
    K *kout = static_cast<K *>(p.keys_out);
    auto emit = [&](auto full_c, auto dec_c) {
        constexpr bool kFull = decltype(full_c)::value;
        constexpr bool kDec = decltype(dec_c)::value;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = tid + k * NT;
            if (kFull || i < valid) {
                const K key = s_keys[i];
                const BaseT dst = s_gbase[digit_of<K>(key, p.shift)] + i;
                kout[dst] = kDec ? key_decode<K>(key, p.xf) : key;
            }
        }
    };
    if (full) {
        if (p.decode_out) emit(std::true_type{}, std::true_type{});
        else emit(std::true_type{}, std::false_type{});
    } else {
        if (p.decode_out) emit(std::false_type{}, std::true_type{});
        else emit(std::false_type{}, std::false_type{});
    }
}

}  // namespace b200
