// mgpu_sort.cu — device-driven multi-GPU sample sort for 32-bit keys (keys-only `sort!`), SURVEY §8(e).
//
// One process per GPU; the host code (cuda.jl_b200/multigpu.py, or a Julia task per device) only strings these calls
// together with ONE all-gather — no host-side splitter search, no count matrix on the host, no per-rank `.item()`:
//
//   1. b200_mgpu_hist16      histogram of the top 16 bits of the (order-mapped) keys of this rank's shard
//                            (msd_hist16_kernel, the first kernel of the single-GPU hybrid sort)
//      -- all-gather of the G histograms (256 KB each) --
//   2. b200_mgpu_plan        on every rank, from the SAME G x 65536 table: global histogram, G-1 splitters at
//                            16-bit-bin granularity (bucket b = bins [spl[b], spl[b+1]), sizes within one bin of N/G),
//                            the exact G x G count matrix, hence where this rank's keys start inside every
//                            destination's receive buffer, how much this rank receives, and the 16-bit histogram of
//                            what it will receive (handed to the local sort, which then skips its own histogram pass)
//   3. b200_mgpu_scatter     ONE pass over the shard: bucket of every key (table lookup on its top byte + the few
//                            splitters inside that byte), atomic rank inside the tile, decoupled look-back over the
//                            tiles per destination, keys staged by destination in shared memory and written as full
//                            lines STRAIGHT INTO THE DESTINATION RANK'S RECEIVE BUFFER (NVLink peer stores): the
//                            all-to-all is this kernel; nothing is written to local HBM first and the keys are read once
//   4. b200_sort_keys_from   local hybrid MSD sort of the received keys, src = receive buffer, dst = result
//
// Equal keys always land in one bucket (a 16-bit bin is never split), so the concatenation of the ranks' results is the
// sorted sequence; with heavy duplicates a rank can receive more than N/G keys — `max_recv_total` tells every rank the
// capacity the receive buffers need before anything is sent.  No reference counterpart (the reference has no multi-GPU
// sort).  Pairs / sortperm keep the stable partition of partition.cu.
#include "sort_msd.cuh"

namespace b200 {

constexpr int MG_MAXG = 16;
constexpr int MG_NT = 512, MG_ITEMS = 20, MG_TILE = MG_NT * MG_ITEMS;

using MgpuPlan = b200_mgpu_plan_t;      // public layout (include/b200arrays.h): the host reads recv_total / max_recv_total

struct MgpuPlanParams {
    const unsigned int *hist_all;            // [G][65536]
    MgpuPlan *plan;
    unsigned int *hist_recv;                 // [65536] histogram of the keys this rank will receive
    unsigned long long *ghist;               // [65536] scratch behind the plan: global histogram
    unsigned long long *cnt;                 // [16][16] scratch: exact count matrix [source][bucket]
    int G, rank;
};

// The plan is four small kernels (a single CTA walking the G x 65536 table three times took 0.43 ms at G = 8):
//   A  (64 CTAs, a thread per bin)  global histogram = sum over the ranks, coalesced
//   B  (1 CTA)                      scan of the global histogram -> G - 1 splitters at bin granularity
//   C  (64 CTAs, a thread per bin)  bucket of every bin, the histogram this rank receives, the G x G count matrix
//                                   (warp-reduced: a warp's 32 bins nearly always share their bucket)
//   D  (1 warp)                     offsets / totals for this rank
__global__ void __launch_bounds__(1024) mgpu_plan_sum_kernel(MgpuPlanParams p) {
    const unsigned int bin = blockIdx.x * 1024u + threadIdx.x;
    unsigned long long g = 0;
    for (int s = 0; s < p.G; ++s) g += p.hist_all[(size_t)s * MSD_BINS + bin];
    p.ghist[bin] = g;
    if (blockIdx.x == 0 && threadIdx.x < MG_MAXG * MG_MAXG) p.cnt[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(1024, 1) mgpu_plan_split_kernel(MgpuPlanParams p) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned int s_spl[MG_MAXG + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = p.G;
    if (tid <= MG_MAXG) s_spl[tid] = 0;
    // thread t owns bins [64 t, 64 t + 64)
    const ulonglong2 *h2 = reinterpret_cast<const ulonglong2 *>(p.ghist) + tid * 32;
    unsigned long long mine = 0;
#pragma unroll 4
    for (int k = 0; k < 32; ++k) { const ulonglong2 v = h2[k]; mine += v.x + v.y; }
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    unsigned long long base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) { if (w < warp) base += s_w[w]; total += s_w[w]; }
    // spl[b] = number of bins whose exclusive prefix is below the b-th target = first bin with prefix >= b N / G
    unsigned long long target[MG_MAXG];
#pragma unroll
    for (int b = 0; b < MG_MAXG; ++b) target[b] = b < G ? (total * (unsigned long long)b) / (unsigned long long)G : ~0ull;
    unsigned long long run = base + incl - mine;
    unsigned int below[MG_MAXG];
#pragma unroll
    for (int b = 0; b < MG_MAXG; ++b) below[b] = 0;
    for (int k = 0; k < 32; ++k) {
        const ulonglong2 v = h2[k];
#pragma unroll
        for (int b = 1; b < MG_MAXG; ++b) below[b] += (b < G && run < target[b]) ? 1u : 0u;
        run += v.x;
#pragma unroll
        for (int b = 1; b < MG_MAXG; ++b) below[b] += (b < G && run < target[b]) ? 1u : 0u;
        run += v.y;
    }
#pragma unroll
    for (int b = 1; b < MG_MAXG; ++b)
        if (b < G && below[b]) atomicAdd(&s_spl[b], below[b]);
    __syncthreads();
    if (tid <= MG_MAXG) p.plan->spl[tid] = tid == 0 ? 0u : (tid < G ? s_spl[tid] : (unsigned int)MSD_BINS);
    if (tid == 0) { p.plan->total = total; p.plan->ticket = 0; p.plan->pad = 0; }
}

__global__ void __launch_bounds__(1024) mgpu_plan_count_kernel(MgpuPlanParams p) {
    __shared__ unsigned int s_spl[MG_MAXG + 1];
    const int G = p.G;
    if (threadIdx.x <= MG_MAXG) s_spl[threadIdx.x] = p.plan->spl[threadIdx.x];
    __syncthreads();
    const unsigned int bin = blockIdx.x * 1024u + threadIdx.x;
    unsigned int b = 0;
    for (int j = 1; j < G; ++j) b += (bin >= s_spl[j]) ? 1u : 0u;
    p.hist_recv[bin] = b == (unsigned int)p.rank ? (unsigned int)p.ghist[bin] : 0u;
    const bool same = __all_sync(0xffffffffu, b == __shfl_sync(0xffffffffu, b, 0));
    for (int s = 0; s < G; ++s) {
        const unsigned int v = p.hist_all[(size_t)s * MSD_BINS + bin];
        if (same) {
            const unsigned int lo = __reduce_add_sync(0xffffffffu, v & 0xFFFFu), hi = __reduce_add_sync(0xffffffffu, v >> 16);
            const unsigned long long t = (unsigned long long)lo + ((unsigned long long)hi << 16);
            if ((threadIdx.x & 31) == 0 && t) atomicAdd(&p.cnt[s * MG_MAXG + b], t);
        } else if (v) {
            atomicAdd(&p.cnt[s * MG_MAXG + b], (unsigned long long)v);
        }
    }
}

__global__ void mgpu_plan_finish_kernel(MgpuPlanParams p) {
    const int tid = threadIdx.x, G = p.G;            // one warp
    auto cnt = [&](int s, int b) { return p.cnt[s * MG_MAXG + b]; };
    if (tid < MG_MAXG) {
        const int b = tid;
        unsigned long long before = 0, recv_b = 0;
        for (int s = 0; s < G; ++s) {
            if (b < G) {
                if (s < p.rank) before += cnt(s, b);
                recv_b += cnt(s, b);
            }
        }
        p.plan->dst_base[b] = b < G ? before : 0;
        p.plan->send_count[b] = b < G ? cnt(p.rank, b) : 0;
        p.plan->recv_count[b] = b < G ? cnt(b, p.rank) : 0;   // from source b
        // max over destinations of what they receive
        unsigned long long mx = b < G ? recv_b : 0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0x0000ffffu, mx, o);
            mx = mx > other ? mx : other;
        }
        // longest shard (source b's keys over all buckets): sizes sortperm's 4-byte tagged index
        unsigned long long len = 0;
        for (int d = 0; d < G; ++d) len += b < G ? cnt(b, d) : 0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0x0000ffffu, len, o);
            len = len > other ? len : other;
        }
        if (tid == 0) {
            p.plan->max_shard_len = (uint32_t)(len > 0xFFFFFFFFull ? 0xFFFFFFFFull : len);
            unsigned long long rt = 0;
            for (int s = 0; s < G; ++s) rt += cnt(s, p.rank);
            p.plan->recv_total = rt;
            p.plan->max_recv_total = mx;
        }
    }
}

// ------------------------------------------------------------------ fused partition + exchange
struct MgpuScatterParams {
    const uint32_t *keys;
    unsigned int n;
    MgpuPlan *plan;
    unsigned int *status;          // [tiles][16], zeroed
    uint32_t *dst[MG_MAXG];        // receive buffers (peer-mapped), one per destination rank
    SortXform xf;
    int G;
};

template <bool IDENT>
__global__ void __launch_bounds__(MG_NT, 2) mgpu_scatter_kernel(MgpuScatterParams p) {
    using LB = LookBits<uint32_t>;
    constexpr int NW = MG_NT / 32;
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_raw);                       // [TILE] RAW keys grouped by bucket
    unsigned char *s_bkt = reinterpret_cast<unsigned char *>(s_keys + MG_TILE);   // [TILE] bucket of every staged key
    __shared__ unsigned int s_whist[NW][MG_MAXG];      // per-warp bucket counters -> warp offsets inside the bucket
    __shared__ unsigned int s_boff[MG_MAXG];           // start of every bucket in the staging buffer
    __shared__ unsigned long long s_gbase[MG_MAXG];    // destination index of staged position 0 of the bucket's run
    __shared__ unsigned int s_spl[MG_MAXG + 1];
    __shared__ unsigned short s_lut[256];              // top byte -> first bucket | (splitters inside the byte) << 8
    __shared__ unsigned int s_tk[2];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int G = p.G;
    const unsigned int tiles = (p.n + MG_TILE - 1) / MG_TILE;
    if (tid <= MG_MAXG) s_spl[tid] = p.plan->spl[tid];
    if (tid == 0) s_tk[0] = atomicAdd(&p.plan->ticket, 1u);
    __syncthreads();
    if (tid < 256) {
        // buckets that start inside or before this top byte: first = #{b >= 1 : spl[b] <= 256 tid}, inside = those in (256 tid, 256 tid + 255]
        unsigned int first = 0, inside = 0;
        for (int b = 1; b < G; ++b) {
            if (s_spl[b] <= (unsigned int)tid * 256u) ++first;
            else if (s_spl[b] < (unsigned int)tid * 256u + 256u) ++inside;
        }
        s_lut[tid] = (unsigned short)(first | (inside << 8));
    }
    unsigned int next_t = 0;
    if (tid == 0) next_t = atomicAdd(&p.plan->ticket, 1u);
    __syncthreads();

    // Software-pipelined over tiles like the single-GPU partition passes: once a tile's keys sit in the staging buffer
    // their registers are dead, so the NEXT tile's loads are issued there, before the look-back and the write-out.
    uint32_t keys[MG_ITEMS];
    auto load_tile = [&](unsigned int t) {
        const unsigned int tile_base = t * MG_TILE;
        const int valid = p.n - tile_base > (unsigned int)MG_TILE ? MG_TILE : (int)(p.n - tile_base);
        const uint32_t *kin = p.keys + tile_base + tid;
        if (valid == MG_TILE) {
#pragma unroll
            for (int i = 0; i < MG_ITEMS; ++i) keys[i] = ldg_stream(kin + i * MG_NT);
        } else {
#pragma unroll
            for (int i = 0; i < MG_ITEMS; ++i) keys[i] = (i * MG_NT + tid < valid) ? ldg_stream(kin + i * MG_NT) : 0u;
        }
    };
    if (s_tk[0] < tiles) load_tile(s_tk[0]);
    for (unsigned int it = 0;; ++it) {
        const unsigned int t = s_tk[it & 1];
        if (t >= tiles) break;
        const unsigned int tile_base = t * MG_TILE;
        const int valid = p.n - tile_base > (unsigned int)MG_TILE ? MG_TILE : (int)(p.n - tile_base);
        unsigned int bk[(MG_ITEMS + 3) / 4];           // bucket ids, 8 bits each
        PackedRanks<MG_ITEMS> ranks;
#pragma unroll
        for (int i = 0; i < MG_ITEMS; ++i) {
            const uint32_t e = IDENT ? keys[i] : key_encode<uint32_t>(keys[i], p.xf);
            const unsigned int t16 = e >> 16;
            const unsigned int l = s_lut[t16 >> 8];
            unsigned int b = l & 0xFFu;
            for (unsigned int j = 0; j < (l >> 8); ++j) b += (t16 >= s_spl[(l & 0xFFu) + 1 + j]) ? 1u : 0u;
            if (i * MG_NT + tid >= valid) b = 0xFFu;          // past the end: no bucket
            if ((i & 3) == 0) bk[i >> 2] = b;
            else bk[i >> 2] |= b << (8 * (i & 3));
        }
        // Ranks without atomics (G <= 16 buckets: 32 lanes adding to a handful of shared counters would serialise):
        // a multi-split by ballots.  NB = bits of a bucket id; the NB ballots of an item give every lane the mask of its
        // peers (same bucket) — its rank is the popcount below it — and lane j, standing in for bucket j, the mask of
        // bucket j, which it adds to the warp's running count of that bucket (kept in lane j's register, fetched by
        // shuffle).  ~25 instructions per key whatever G; the lane-private counting sweeps this replaces cost 5 G.
        auto rank_all = [&](auto nb_tag) {
            constexpr int NB = decltype(nb_tag)::value;
            const unsigned int lane = tid & 31, lt = (1u << lane) - 1u;
            unsigned int run = 0;                               // lane j: keys of bucket j seen so far in this warp
#pragma unroll
            for (int i = 0; i < MG_ITEMS; ++i) {
                const unsigned int b = (bk[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                const bool live = b != 0xFFu;
                unsigned int peers = __ballot_sync(0xffffffffu, live), mine = peers;
#pragma unroll
                for (int bit = 0; bit < NB; ++bit) {
                    const unsigned int bal = __ballot_sync(0xffffffffu, (b >> bit) & 1u);
                    peers &= ((b >> bit) & 1u) ? bal : ~bal;
                    mine &= ((lane >> bit) & 1u) ? bal : ~bal;
                }
                const unsigned int base = __shfl_sync(0xffffffffu, run, b & 31u);
                ranks.set(i, live ? base + __popc(peers & lt) : 0u);
                run += (lane < (1u << NB)) ? __popc(mine) : 0u;
            }
            if (lane < (unsigned int)MG_MAXG) s_whist[warp][lane] = lane < (1u << NB) ? run : 0u;
        };
        if (G <= 2) rank_all(std::integral_constant<int, 1>{});
        else if (G <= 4) rank_all(std::integral_constant<int, 2>{});
        else if (G <= 8) rank_all(std::integral_constant<int, 3>{});
        else rank_all(std::integral_constant<int, 4>{});
        __syncthreads();                                                             // (A)
        // bucket threads: warp offsets inside the bucket, tile count, publish, scan over the buckets
        unsigned int cnt = 0;
        if (tid < MG_MAXG) {
#pragma unroll
            for (int w = 0; w < NW; ++w) { const unsigned int c = s_whist[w][tid]; s_whist[w][tid] = cnt; cnt += c; }
            if (tid < G) st_relaxed<uint32_t>(p.status + (size_t)t * MG_MAXG + tid, (t == 0 ? LB::INCLUSIVE : LB::PARTIAL) | cnt);
            unsigned int incl = cnt;
#pragma unroll
            for (int o = 1; o < MG_MAXG; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0x0000ffffu, incl, o);
                if (tid >= o) incl += up;
            }
            s_boff[tid] = incl - cnt;
        }
        if (tid == 0) s_tk[(it + 1) & 1] = next_t;
        __syncthreads();                                                             // (B)
#pragma unroll
        for (int i = 0; i < MG_ITEMS; ++i) {
            if (i * MG_NT + tid < valid) {
                const unsigned int b = (bk[i >> 2] >> (8 * (i & 3))) & 0xFFu;
                const unsigned int pos = s_boff[b] + s_whist[warp][b] + ranks.get(i);
                s_keys[pos] = keys[i];
                s_bkt[pos] = (unsigned char)b;
            }
        }
        {
            const unsigned int tn = s_tk[(it + 1) & 1];        // written before barrier (B)
            if (tn < tiles) {
                if (tid == 0) next_t = atomicAdd(&p.plan->ticket, 1u);
                load_tile(tn);
            }
        }
        // look-back per destination
        if (tid < MG_MAXG) {
            unsigned long long excl = 0;
            if (t > 0 && tid < G) {
                int u = (int)t - 1;
                bool done = false;
                while (!done) {
                    uint32_t w[RX_LOOK];
                    bool all_valid;
                    do {
                        all_valid = true;
#pragma unroll
                        for (int k = 0; k < RX_LOOK; ++k) {
                            w[k] = LB::INCLUSIVE;
                            if (u - k >= 0) w[k] = ld_relaxed<uint32_t>(p.status + (size_t)(u - k) * MG_MAXG + tid);
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
                st_relaxed<uint32_t>(p.status + (size_t)t * MG_MAXG + tid, LB::INCLUSIVE | (uint32_t)(excl + cnt));
            }
            s_gbase[tid] = (tid < G ? p.plan->dst_base[tid] : 0ull) + excl - s_boff[tid];
        }
        __syncthreads();                                                             // (C)
        // write out: consecutive threads -> consecutive addresses of one destination's receive buffer
#pragma unroll
        for (int k = 0; k < MG_ITEMS; ++k) {
            const int i = tid + k * MG_NT;
            if (i < valid) {
                const unsigned int b = s_bkt[i];
                p.dst[b][s_gbase[b] + (unsigned long long)i] = s_keys[i];
            }
        }
        // barrier (A) of the next iteration separates this write-out from the next tile's staging writes; its per-warp
        // counters (s_whist) are rewritten only after barrier (C) above, which every reader has passed
    }
}

}  // namespace b200

extern "C" {

int32_t b200_mgpu_plan_bytes(size_t *bytes) {
    if (!bytes) return B200_INVALID_VALUE;
    // the public struct, then scratch for the plan kernels: global histogram (uint64[65536]) + count matrix (uint64[256])
    *bytes = b200::align_up(sizeof(b200::MgpuPlan), 256) + (size_t)b200::MSD_BINS * 8 + 256 * 8;
    return B200_OK;
}

static bool mgpu_xform(int dtype, int descending, b200::SortXform &xf, int nan_canonical = 0);

int32_t b200_mgpu_hist16(const void *keys, int32_t dtype_key, int64_t n, int32_t descending, int32_t nan_canonical,
                         void *hist16_dev, b200_stream_t stream) {
    using namespace b200;
    SortXform xf;
    if (!mgpu_xform(dtype_key, descending, xf, nan_canonical)) return B200_UNSUPPORTED;
    if (n < 0 || n > (1LL << 30) || !hist16_dev || (n > 0 && !keys)) return B200_INVALID_VALUE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B200_CUDA_TRY(cudaMemsetAsync(hist16_dev, 0, (size_t)MSD_BINS * 4, st));
    if (n == 0) return B200_OK;
    const bool ident = xf.base_enc == 0 && xf.fmn == 0 && xf.mag_mask == 0 && xf.desc == 0;
    const size_t smem = (size_t)MSD_H16_WORDS * 4;
    int64_t grid = devinfo().sm_count;
    const int64_t need = cdiv(n, (int64_t)MSD_H16_THREADS * MSD_H16_VPT * 4);
    if (grid > need) grid = need;
    if (ident) {
        B200_CUDA_TRY(cudaFuncSetAttribute(msd_hist16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        msd_hist16_kernel<true><<<(unsigned int)grid, MSD_H16_THREADS, smem, st>>>(static_cast<const uint32_t *>(keys), n, xf,
                                                                                 static_cast<unsigned int *>(hist16_dev));
    } else {
        B200_CUDA_TRY(cudaFuncSetAttribute(msd_hist16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        msd_hist16_kernel<false><<<(unsigned int)grid, MSD_H16_THREADS, smem, st>>>(static_cast<const uint32_t *>(keys), n, xf,
                                                                                  static_cast<unsigned int *>(hist16_dev));
    }
    B200_CHECK_LAUNCH();
    return B200_OK;
}

int32_t b200_mgpu_plan(const void *hist_all_dev, int32_t nranks, int32_t rank, void *plan_dev, void *hist_recv_dev,
                       b200_stream_t stream) {
    using namespace b200;
    if (!hist_all_dev || !plan_dev || !hist_recv_dev || nranks < 1 || nranks > MG_MAXG || rank < 0 || rank >= nranks)
        return B200_INVALID_VALUE;
    MgpuPlanParams pp{};
    pp.hist_all = static_cast<const unsigned int *>(hist_all_dev);
    pp.plan = static_cast<MgpuPlan *>(plan_dev);
    pp.hist_recv = static_cast<unsigned int *>(hist_recv_dev);
    pp.G = nranks; pp.rank = rank;
    char *scratch = static_cast<char *>(plan_dev) + align_up(sizeof(MgpuPlan), 256);
    pp.ghist = reinterpret_cast<unsigned long long *>(scratch);
    pp.cnt = pp.ghist + MSD_BINS;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    mgpu_plan_sum_kernel<<<MSD_BINS / 1024, 1024, 0, st>>>(pp);
    B200_CHECK_LAUNCH();
    mgpu_plan_split_kernel<<<1, 1024, 0, st>>>(pp);
    B200_CHECK_LAUNCH();
    mgpu_plan_count_kernel<<<MSD_BINS / 1024, 1024, 0, st>>>(pp);
    B200_CHECK_LAUNCH();
    mgpu_plan_finish_kernel<<<1, 32, 0, st>>>(pp);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

int32_t b200_mgpu_scatter_workspace_bytes(int64_t n, size_t *bytes) {
    using namespace b200;
    if (!bytes || n < 0) return B200_INVALID_VALUE;
    *bytes = align_up((size_t)cdiv(n, MG_TILE) * MG_MAXG * 4, 256);
    return B200_OK;
}

int32_t b200_mgpu_scatter(void *workspace, size_t workspace_bytes, const void *keys, int32_t dtype_key, int64_t n,
                          int32_t descending, void *plan_dev, void *const *dst_keys, int32_t nranks,
                          b200_stream_t stream) {
    using namespace b200;
    SortXform xf;
    if (!mgpu_xform(dtype_key, descending, xf)) return B200_UNSUPPORTED;
    if (n < 0 || n > (1LL << 30) || !plan_dev || !dst_keys || nranks < 1 || nranks > MG_MAXG) return B200_INVALID_VALUE;
    if (n == 0) return B200_OK;
    size_t need = 0;
    b200_mgpu_scatter_workspace_bytes(n, &need);
    if (!workspace || !keys) return B200_INVALID_VALUE;
    if (workspace_bytes < need) return B200_WORKSPACE_TOO_SMALL;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B200_CUDA_TRY(cudaMemsetAsync(workspace, 0, need, st));
    MgpuScatterParams sp{};
    sp.keys = static_cast<const uint32_t *>(keys);
    sp.n = (unsigned int)n;
    sp.plan = static_cast<MgpuPlan *>(plan_dev);
    sp.status = static_cast<unsigned int *>(workspace);
    for (int b = 0; b < MG_MAXG; ++b) sp.dst[b] = static_cast<uint32_t *>(b < nranks ? dst_keys[b] : nullptr);
    sp.xf = xf; sp.G = nranks;
    const bool ident = xf.base_enc == 0 && xf.fmn == 0 && xf.mag_mask == 0 && xf.desc == 0;
    const size_t smem = (size_t)MG_TILE * 5;
    int64_t grid = (int64_t)devinfo().sm_count * 2;
    const int64_t tiles = cdiv(n, MG_TILE);
    if (grid > tiles) grid = tiles;
    if (ident) {
        B200_CUDA_TRY(cudaFuncSetAttribute(mgpu_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mgpu_scatter_kernel<true><<<(unsigned int)grid, MG_NT, smem, st>>>(sp);
    } else {
        B200_CUDA_TRY(cudaFuncSetAttribute(mgpu_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mgpu_scatter_kernel<false><<<(unsigned int)grid, MG_NT, smem, st>>>(sp);
    }
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // extern "C"

// Same key transform as api_sort.cu's make_xform for the 32-bit key types.  nan_canonical = 0: keys-only sort! (NaN
// payloads kept, sign dropped); 1: sortperm / pairs (every NaN maps to one key, like partition.cu's part_xform).
static bool mgpu_xform(int dtype, int descending, b200::SortXform &x, int nan_canonical) {
    using namespace b200;
    x = SortXform{};
    const uint64_t all = 0xFFFFFFFFull, sign = 0x80000000ull;
    x.sign = sign;
    x.desc = descending ? all : 0;
    x.inf_bits = all;
    switch (dtype) {
        case B200_U32: return true;
        case B200_I32: x.base_enc = sign; x.base_dec = sign; return true;
        case B200_F32:
            x.inf_bits = 0x7F800000ull;
            x.base_enc = sign; x.base_dec = all; x.fmn = all & ~sign; x.mag_mask = all & ~sign;
            x.nan_or = nan_canonical ? all : 0;
            return true;
        default: return false;
    }
}
