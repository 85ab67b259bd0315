// sort_pipe.cuh — persistent, software-pipelined onesweep pass.
//
// The tile-per-CTA kernel in sort.cuh has the same convoy as the first scan: every resident
// tile ranks at the same time, then all of them sit in the digit look-back while HBM idles
// (ncu: ~26 % of warp samples in the look-back loop).  Here each CTA is persistent:
//   * keys arrive through a ring of cp.async.bulk (TMA) stages, S tiles ahead;
//   * iteration i RANKS tile i+1 first (ballots, per-digit scan), publishes its digit counts
//     and scatters it into one of two shared-memory staging buffers — so a tile's counts are
//     visible a whole iteration before anyone needs them;
//   * only then resolves tile i's per-digit prefix by look-back (its predecessors' counts were
//     published just as early, so the walk rarely waits) and writes tile i out.
// Tiles are handed out by an atomic ticket (tile index = ticket), so a CTA only waits on
// tiles some resident CTA already owns.
#pragma once
#include "sort.cuh"

namespace b200 {

constexpr int RXP_STAGES = 3;

template <class K, class V, class LookT, int ITEMS>
__global__ void __launch_bounds__(RX_THREADS) radix_onesweep_pipe_kernel(OnesweepParams p, long long ntiles) {
    using LB = LookBits<LookT>;
    constexpr bool HAS_V = PayTraits<V>::HAS;
    using VT = typename PayTraits<V>::type;
    constexpr int TILE = RX_THREADS * ITEMS;
    constexpr int WARP_ITEMS = 32 * ITEMS;
    constexpr int S = RXP_STAGES;
    constexpr size_t STAGE_BYTES = (size_t)TILE * sizeof(K);
    using BaseT = typename std::conditional<sizeof(LookT) == 4, int, long long>::type;

    extern __shared__ __align__(128) unsigned char s_raw[];
    // layout: [S key stages][2 staging key buffers][2 staging payload buffers][warp hist][gbase x2]
    unsigned char *s_stage = s_raw;
    K *s_keys = reinterpret_cast<K *>(s_raw + S * STAGE_BYTES);                           // [2][TILE]
    VT *s_vals = reinterpret_cast<VT *>(reinterpret_cast<unsigned char *>(s_keys) + 2 * STAGE_BYTES);   // [2][TILE] (pairs only)
    unsigned int *s_whist = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned char *>(s_vals) +
                                                             (HAS_V ? 2 * (size_t)TILE * sizeof(VT) : 0));
    long long *s_gbase_raw = reinterpret_cast<long long *>(s_whist + RX_WARPS * RX_RADIX);  // [2][256]
    BaseT *s_gbase = reinterpret_cast<BaseT *>(s_gbase_raw);
    __shared__ __align__(8) uint64_t s_full[S];
    __shared__ long long s_tile[S];
    __shared__ int s_use[S];
    __shared__ unsigned int s_wsum[RX_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = tid;                         // digit owned by this thread in the per-digit steps
    const int idx0 = warp * WARP_ITEMS + lane;
    unsigned int *whist = s_whist + warp * RX_RADIX;
    const unsigned int lt_mask = (1u << lane) - 1u;
    LookT *status = static_cast<LookT *>(p.status);
    const unsigned long long bin_d = p.bins[d];

    auto tile_valid = [&](long long t) -> int {
        const long long remain = p.n - t * TILE;
        return remain > TILE ? TILE : (int)remain;
    };
    // thread 0: put ticket t into `stage` and start its key load (full tiles only)
    auto fill = [&](int stage, long long t) {
        s_tile[stage] = t;
        s_use[stage] += 1;
        if (t < ntiles) {
            if (tile_valid(t) == TILE) {
                mbar_expect_tx(&s_full[stage], (uint32_t)STAGE_BYTES);
                tma_load_1d(s_stage + stage * STAGE_BYTES, static_cast<const K *>(p.keys_in) + t * TILE,
                            (uint32_t)STAGE_BYTES, &s_full[stage]);
            } else {
                mbar_arrive(&s_full[stage]);   // ragged last tile: read with ordinary loads
            }
        }
    };

    // Rank the tile in `stage`, publish its digit counts, scatter it (sorted by digit) into staging
    // buffer `buf`.  Returns this thread's digit count (`cnt`, padding excluded) and the digit
    // run's start inside the staging buffer (`doff`).  next_ticket refills the stage.
    auto rank_tile = [&](int stage, int buf, long long next_ticket, unsigned int &cnt, unsigned int &doff) {
        const long long t = s_tile[stage];
        const long long tile_base = t * TILE;
        const int valid = tile_valid(t);
        const bool full = valid == TILE;
#pragma unroll
        for (int k = 0; k < RX_RADIX / 32; ++k) whist[k * 32 + lane] = 0;
        mbar_wait(&s_full[stage], (uint32_t)((s_use[stage] - 1) & 1));
        K keys[ITEMS];
        if (full) {
            const K *sk = reinterpret_cast<const K *>(s_stage + stage * STAGE_BYTES) + idx0;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) keys[i] = sk[i * 32];
        } else {
            const K *kin = static_cast<const K *>(p.keys_in) + tile_base + idx0;
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
                if (idx0 + i * 32 >= valid) keys[i] = (K)~(K)0;
        }
        // payload loads are issued now and consumed after the ranking
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
        __syncwarp();
        unsigned short ranks[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const unsigned int digit = digit_of<K>(keys[i], p.shift);
            unsigned int peers;
            if ((i & 3) == 3) {
                peers = __match_any_sync(0xffffffffu, digit);
            } else {
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
                peers = ~differ;
            }
            const unsigned int lower = peers & lt_mask;
            const unsigned int below = __popc(lower);
            volatile unsigned int *ctr = whist + digit;
            const unsigned int old = *ctr;
            __syncwarp();
            if (lower == 0u) *ctr = old + __popc(peers);
            __syncwarp();
            ranks[i] = (unsigned short)(old + below);
        }
        __syncthreads();                               // (1) warp histograms final, stage fully read
        if (tid == 0) fill(stage, next_ticket);        // the stage is free: prefetch S tiles ahead

        unsigned int tile_count = 0;
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) {
            const unsigned int c = s_whist[w * RX_RADIX + d];
            s_whist[w * RX_RADIX + d] = tile_count;
            tile_count += c;
        }
        const unsigned int pad = (unsigned int)(TILE - valid);
        cnt = (d == RX_RADIX - 1) ? tile_count - pad : tile_count;
        // publish one iteration before this tile's own look-back
        st_relaxed<LookT>(status + t * RX_RADIX + d, (LookT)((t > 0 ? LB::PARTIAL : LB::INCLUSIVE) | (LookT)cnt));

        unsigned int incl = tile_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();                               // (2)
        unsigned int wbase = 0;
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) wbase += (w < warp) ? s_wsum[w] : 0u;
        doff = wbase + incl - tile_count;
#pragma unroll
        for (int w = 0; w < RX_WARPS; ++w) s_whist[w * RX_RADIX + d] += doff;
        __syncthreads();                               // (3) per-warp scatter offsets ready

        K *sk_out = s_keys + (size_t)buf * TILE;
        VT *sv_out = s_vals + (size_t)buf * TILE;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const unsigned int digit = digit_of<K>(keys[i], p.shift);
            const unsigned int pos = whist[digit] + ranks[i];
            sk_out[pos] = keys[i];
            if constexpr (HAS_V) sv_out[pos] = vals[i];
        }
    };

    // per-digit decoupled look-back for tile t; writes the global base of the digit run
    auto lookback = [&](long long t, int buf, unsigned int cnt, unsigned int doff) {
        unsigned long long excl = 0;
        if (t > 0) {
            long long u = t - 1;
            bool done = false;
            while (!done) {
                LookT w[RX_LOOK];
                bool all_valid;
                do {
                    all_valid = true;
#pragma unroll
                    for (int k = 0; k < RX_LOOK; ++k) {
                        w[k] = LB::INCLUSIVE;
                        if (u - k >= 0) w[k] = ld_relaxed<LookT>(status + (u - k) * RX_RADIX + d);
                        all_valid = all_valid && ((w[k] & LB::FLAGS) != 0);
                    }
                } while (!all_valid);
#pragma unroll
                for (int k = 0; k < RX_LOOK; ++k) {
                    if (!done) {
                        excl += (unsigned long long)(w[k] & LB::MASK);
                        done = (w[k] & LB::FLAGS) == LB::INCLUSIVE;
                    }
                }
                u -= RX_LOOK;
            }
            st_relaxed<LookT>(status + t * RX_RADIX + d, (LookT)(LB::INCLUSIVE | (LookT)(excl + cnt)));
        }
        s_gbase[buf * RX_RADIX + d] = (BaseT)((long long)(bin_d + excl) - (long long)doff);
    };

    auto write_out = [&](long long t, int buf) {
        const int valid = tile_valid(t);
        const K *sk = s_keys + (size_t)buf * TILE;
        const VT *sv = s_vals + (size_t)buf * TILE;
        const BaseT *gb = s_gbase + buf * RX_RADIX;
        K *kout = static_cast<K *>(p.keys_out);
        auto emit = [&](auto full_c, auto mode_c) {
            constexpr bool kFull = decltype(full_c)::value;
            constexpr int kMode = decltype(mode_c)::value;   // 0: keys as is, 1: keys decoded, 2: no keys
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const int i = tid + k * RX_THREADS;
                if (kFull || i < valid) {
                    const K key = sk[i];
                    const BaseT dst = gb[digit_of<K>(key, p.shift)] + i;
                    if constexpr (kMode == 0) kout[dst] = key;
                    if constexpr (kMode == 1) kout[dst] = key_decode<K>(key, p.xf);
                    if constexpr (HAS_V) {
                        const VT v = sv[i];
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
        if (valid == TILE) emit_mode(std::true_type{});
        else emit_mode(std::false_type{});
    };

    // ---- prologue: S tickets, then rank the first tile
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < S; ++st) { mbar_init(&s_full[st], 1); s_use[st] = 0; }
        fence_proxy_async_smem();
#pragma unroll
        for (int st = 0; st < S; ++st) fill(st, (long long)atomicAdd(p.ticket, 1ull));
    }
    __syncthreads();
    long long t_cur = s_tile[0];
    if (t_cur >= ntiles) return;
    long long tk = 0;
    if (tid == 0) tk = (long long)atomicAdd(p.ticket, 1ull);
    unsigned int cnt_cur, doff_cur, cnt_nxt = 0, doff_nxt = 0;
    rank_tile(0, 0, tk, cnt_cur, doff_cur);

    for (long long it = 0;; ++it) {
        const int nxt_stage = (int)((it + 1) % S);
        const int buf = (int)(it & 1);
        const long long t_nxt = s_tile[nxt_stage];      // written >= 1 barrier ago
        const bool have_next = t_nxt < ntiles;
        if (have_next) {
            long long tk2 = 0;
            if (tid == 0) tk2 = (long long)atomicAdd(p.ticket, 1ull);
            rank_tile(nxt_stage, buf ^ 1, tk2, cnt_nxt, doff_nxt);
        }
        lookback(t_cur, buf, cnt_cur, doff_cur);
        __syncthreads();                               // (4) global bases + staged keys of the current tile visible
        write_out(t_cur, buf);
        if (!have_next) break;
        // The next iteration overwrites staging buffer `buf` and s_gbase[buf] only after its own
        // barriers (1)-(3), which every thread reaches after finishing this write_out.
        t_cur = t_nxt; cnt_cur = cnt_nxt; doff_cur = doff_nxt;
    }
}

}  // namespace b200
