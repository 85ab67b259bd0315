// partition.cu — stable multi-way partition of keys (and an 8-byte payload) by G-1 splitters,
// G <= 16.  First stage of the multi-GPU sample sort: every rank cuts its UNSORTED
// shard into G destination buckets, exchanges them (all-to-all) and sorts once — instead of
// sorting, cutting, exchanging and sorting again.  No reference counterpart (the reference has no
// multi-GPU sort).  Bucket b receives the keys with splitter[b-1] <= key < splitter[b] in the
// order-preserving key space of the sort (isless / rev); within a bucket the input order is kept.
//
// Three small kernels: per-tile bucket counts -> per-bucket exclusive scan over the tiles ->
// ranked scatter (same ballot ranking as the onesweep pass, 4 bits instead of 8).
#include "sort.cuh"

namespace b200 {

constexpr int PT_ITEMS = 16;
constexpr int PT_TILE = RX_THREADS * PT_ITEMS;
constexpr int PT_MAXB = 16;

struct PartParams {
    const void *keys_in;
    void *keys_out;
    const void *vals_in;     // optional 8-byte payload
    uint32_t index_tag;      // VM == 2: the payload is generated, index_tag | (element index in keys_in), 4 bytes
    void *vals_out;
    // scatter destinations: bucket b goes to dst_keys[b] (+ dst_vals[b]) starting at element dst_base[b];
    // local partition: all the same buffer with the packed bucket bases; multi-GPU: peer (NVLink-mapped) buffers
    void *dst_keys[16];
    void *dst_vals[16];
    long long dst_base[16];
    int32_t local_bases;     // 1: dst_base is derived from bucket_counts on the device (packed local output)
    const void *splitters;   // G-1 raw keys (device)
    const b200_mgpu_plan_t *plan;   // non-null (32-bit keys): splitters are the plan's 16-bit bin boundaries and the
                                    // destination bases come from the plan — both stay on the device (mgpu_sort.cu)
    long long *tile_counts;  // [G][tiles]: counts, then exclusive prefixes over tiles
    long long *bucket_counts;// [G]
    int64_t n;
    int64_t tiles;
    int32_t G;
    SortXform xf;
};

template <class K>
__device__ __forceinline__ void load_tile_buckets(const PartParams &p, const K *s_spl, int64_t tile_base, int valid,
                                                  int idx0, K (&keys)[PT_ITEMS], unsigned char (&bid)[PT_ITEMS]) {
    const K *kin = static_cast<const K *>(p.keys_in) + tile_base + idx0;
#pragma unroll
    for (int i = 0; i < PT_ITEMS; ++i) keys[i] = (idx0 + i * 32 < valid) ? ldg_stream(kin + i * 32) : (K)0;
#pragma unroll
    for (int i = 0; i < PT_ITEMS; ++i) {
        const K e = key_encode<K>(keys[i], p.xf);
        int b = 0;
        for (int s = 0; s < p.G - 1; ++s) b += (e >= s_spl[s]) ? 1 : 0;
        bid[i] = (idx0 + i * 32 < valid) ? (unsigned char)b : (unsigned char)0xFF;   // padding matches no bucket
    }
}

template <class K>
__global__ void __launch_bounds__(RX_THREADS) partition_count_kernel(PartParams p) {
    __shared__ K s_spl[PT_MAXB];
    __shared__ unsigned int s_cnt[PT_MAXB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < p.G - 1) s_spl[tid] = p.plan ? (K)((unsigned long long)p.plan->spl[tid + 1] << (8 * sizeof(K) - 16))
                                           : key_encode<K>(static_cast<const K *>(p.splitters)[tid], p.xf);
    if (tid < PT_MAXB) s_cnt[tid] = 0;
    __syncthreads();
    const int64_t t = blockIdx.x, tile_base = t * PT_TILE;
    const int valid = (int)((p.n - tile_base) < PT_TILE ? (p.n - tile_base) : PT_TILE);
    K keys[PT_ITEMS];
    unsigned char bid[PT_ITEMS];
    load_tile_buckets<K>(p, s_spl, tile_base, valid, warp * (32 * PT_ITEMS) + lane, keys, bid);
    // per-thread counts packed 8 bits per bucket (<= 16 items per thread), widened to 16-bit fields
    // before the warp reduction (<= 512 per warp), then one shared atomic per bucket per warp
    unsigned long long pk[2] = {0ull, 0ull};
#pragma unroll
    for (int i = 0; i < PT_ITEMS; ++i) {
        const unsigned int bsel = bid[i];
        if (bsel < PT_MAXB) pk[bsel >> 3] += 1ull << ((bsel & 7) * 8);
    }
    unsigned long long wide[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        wide[2 * h] = pk[h] & 0x00FF00FF00FF00FFull;            // buckets 8h + {0,2,4,6}
        wide[2 * h + 1] = (pk[h] >> 8) & 0x00FF00FF00FF00FFull; // buckets 8h + {1,3,5,7}
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) wide[q] += __shfl_xor_sync(0xffffffffu, wide[q], o);
    }
    if (lane < p.G) {
        const int b = lane;
        const unsigned long long w = wide[2 * (b >> 3) + (b & 1)];
        const unsigned int c = (unsigned int)((w >> (16 * ((b & 7) >> 1))) & 0xFFFFull);
        if (c) atomicAdd(&s_cnt[b], c);
    }
    __syncthreads();
    if (tid < p.G) p.tile_counts[(int64_t)tid * p.tiles + t] = s_cnt[tid];
}

// one CTA per bucket: exclusive scan of its tile counts, total to bucket_counts
__global__ void __launch_bounds__(256) partition_scan_kernel(PartParams p) {
    __shared__ long long s_w[8];
    __shared__ long long s_run;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long *row = p.tile_counts + (int64_t)b * p.tiles;
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int64_t base = 0; base < p.tiles; base += 256) {
        const int64_t i = base + tid;
        const long long v = i < p.tiles ? row[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        long long wb = 0;
        for (int w = 0; w < warp; ++w) wb += s_w[w];
        const long long run = s_run;
        if (i < p.tiles) row[i] = run + wb + incl - v;
        __syncthreads();
        if (tid == 255) s_run = run + wb + incl;
        __syncthreads();
    }
    if (tid == 0) p.bucket_counts[b] = s_run;
}

// Ranked scatter.  Keys (and payload) are first staged in shared memory grouped by bucket, then
// written out with consecutive threads on consecutive addresses of each bucket run (~TILE/G keys,
// i.e. full 128-byte lines): this is what makes stores into PEER memory over NVLink efficient.
// VM: 0 keys only, 1 an 8-byte payload read from vals_in, 2 a generated 4-byte payload (tagged source index)
template <class K, int VM>
__global__ void __launch_bounds__(RX_THREADS) partition_scatter_kernel(PartParams p) {
    using V = typename std::conditional<VM == 2, unsigned int, unsigned long long>::type;
    extern __shared__ __align__(16) unsigned char s_raw[];
    K *s_keys = reinterpret_cast<K *>(s_raw);                                        // [PT_TILE]
    unsigned char *s_bid = reinterpret_cast<unsigned char *>(s_keys + PT_TILE);      // [PT_TILE]
    V *s_vals = reinterpret_cast<V *>(s_bid + PT_TILE);                              // [PT_TILE] (VM != 0)
    __shared__ K s_spl[PT_MAXB];
    __shared__ unsigned int s_wcnt[RX_WARPS][PT_MAXB];
    __shared__ unsigned int s_boff[PT_MAXB];       // start of each bucket run inside the staged tile
    __shared__ long long s_dst[PT_MAXB];           // destination element index of the run's first key
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < p.G - 1) s_spl[tid] = p.plan ? (K)((unsigned long long)p.plan->spl[tid + 1] << (8 * sizeof(K) - 16))
                                           : key_encode<K>(static_cast<const K *>(p.splitters)[tid], p.xf);
    if (tid < RX_WARPS * PT_MAXB) (&s_wcnt[0][0])[tid] = 0;
    __syncthreads();
    const int64_t t = blockIdx.x, tile_base = t * PT_TILE;
    const int valid = (int)((p.n - tile_base) < PT_TILE ? (p.n - tile_base) : PT_TILE);
    const int idx0 = warp * (32 * PT_ITEMS) + lane;
    K keys[PT_ITEMS];
    unsigned char bid[PT_ITEMS];
    load_tile_buckets<K>(p, s_spl, tile_base, valid, idx0, keys, bid);
    // stable rank inside the warp, by (item, lane)
    unsigned short ranks[PT_ITEMS];
    const unsigned int lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < PT_ITEMS; ++i) {
        const unsigned int d = bid[i];
        unsigned int differ = 0;
#pragma unroll
        for (int bit = 0; bit < 8; bit += (bit == 3 ? 4 : 1)) {   // bits 0..3 and bit 7 (the padding marker)
            const bool one = (d >> bit) & 1u;
            const unsigned int bal = __ballot_sync(0xffffffffu, one);
            differ |= bal ^ (one ? 0xffffffffu : 0u);
        }
        const unsigned int peers = ~differ;
        const unsigned int lower = peers & lt_mask;
        const bool real = d < PT_MAXB;
        volatile unsigned int *ctr = &s_wcnt[warp][real ? d : 0];
        const unsigned int old = real ? *ctr : 0u;
        __syncwarp();
        if (real && lower == 0u) *ctr = old + __popc(peers);
        __syncwarp();
        ranks[i] = (unsigned short)(old + __popc(lower));
    }
    __syncthreads();
    if (tid < p.G) {
        unsigned int run = 0;
        for (int w = 0; w < RX_WARPS; ++w) { const unsigned int c = s_wcnt[w][tid]; s_wcnt[w][tid] = run; run += c; }
        s_boff[tid] = run;                         // bucket total for now
        long long base = p.plan ? (long long)p.plan->dst_base[tid] : p.dst_base[tid];
        if (p.local_bases) { base = 0; for (int b = 0; b < tid; ++b) base += p.bucket_counts[b]; }
        s_dst[tid] = base + p.tile_counts[(int64_t)tid * p.tiles + t];
    }
    __syncthreads();
    if (tid == 0) {
        unsigned int run = 0;
        for (int b = 0; b < p.G; ++b) { const unsigned int c = s_boff[b]; s_boff[b] = run; run += c; }
    }
    __syncthreads();
    const unsigned long long *vin = static_cast<const unsigned long long *>(p.vals_in) + tile_base + idx0;
#pragma unroll
    for (int i = 0; i < PT_ITEMS; ++i) {
        if (bid[i] < PT_MAXB) {
            const unsigned int pos = s_boff[bid[i]] + s_wcnt[warp][bid[i]] + ranks[i];
            s_keys[pos] = keys[i];
            s_bid[pos] = bid[i];
            if constexpr (VM == 1) s_vals[pos] = ldg_stream(vin + i * 32);
            if constexpr (VM == 2) s_vals[pos] = p.index_tag | (unsigned int)(tile_base + idx0 + i * 32);
        }
    }
    __syncthreads();
    for (int i = tid; i < valid; i += RX_THREADS) {
        const int b = s_bid[i];
        const long long dst = s_dst[b] + (i - (int)s_boff[b]);
        static_cast<K *>(p.dst_keys[b])[dst] = s_keys[i];
        if constexpr (VM != 0) static_cast<V *>(p.dst_vals[b])[dst] = s_vals[i];
    }
}

static size_t scatter_smem(int kb, int vm) { return (size_t)PT_TILE * (kb + 1) + 16 + (vm == 1 ? (size_t)PT_TILE * 8 : vm == 2 ? (size_t)PT_TILE * 4 : 0); }

template <class K, int VM>
static int32_t launch_scatter(const PartParams &p, cudaStream_t st) {
    const size_t smem = scatter_smem((int)sizeof(K), VM);
    static thread_local bool attr_set[64];
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    auto kern = partition_scatter_kernel<K, VM>;
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    kern<<<(unsigned int)p.tiles, RX_THREADS, smem, st>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

// phase: 1 = count + scan (leaves the per-tile prefixes in the workspace), 2 = scatter, 3 = both
template <class K>
static int32_t run_partition(void *ws, size_t wsb, PartParams p, int vm, int phase, cudaStream_t st, size_t *query) {
    p.tiles = cdiv(p.n, PT_TILE);
    Carver c(ws);
    long long *tile_counts = c.take<long long>((size_t)p.G * p.tiles);
    if (query) { *query = c.bytes(); return B200_OK; }
    if (c.bytes() > wsb) return B200_WORKSPACE_TOO_SMALL;
    if (!ws) return B200_INVALID_VALUE;
    if (p.tiles > 2147483647LL) return B200_INVALID_VALUE;
    p.tile_counts = tile_counts;
    if (phase & 1) {
        partition_count_kernel<K><<<(unsigned int)p.tiles, RX_THREADS, 0, st>>>(p);
        B200_CHECK_LAUNCH();
        partition_scan_kernel<<<p.G, 256, 0, st>>>(p);
        B200_CHECK_LAUNCH();
    }
    if (phase & 2) {
        const int32_t rc = vm == 1 ? launch_scatter<K, 1>(p, st) : vm == 2 ? launch_scatter<K, 2>(p, st) : launch_scatter<K, 0>(p, st);
        if (rc != B200_OK) return rc;
    }
    return B200_OK;
}

template <class... Args> static int32_t run_partition_kb(int kb, Args... args) {
    switch (kb) {
        case 1: return run_partition<uint8_t>(args...);
        case 2: return run_partition<uint16_t>(args...);
        case 4: return run_partition<uint32_t>(args...);
        case 8: return run_partition<uint64_t>(args...);
        default: return B200_UNSUPPORTED;
    }
}

// same transform as the sort (api_sort.cu make_xform), NaNs canonical so that they all go last
static bool part_xform(int dtype, int descending, SortXform &x, int &kb) {
    x = SortXform{};
    kb = dtype_size(dtype);
    if (kb == 0) return false;
    const uint64_t all = kb == 8 ? ~0ull : ((1ull << (8 * kb)) - 1);
    const uint64_t sign = 1ull << (8 * kb - 1);
    x.sign = sign; x.desc = descending ? all : 0; x.inf_bits = all;
    bool is_float = false;
    switch (dtype) {
        case B200_F16: is_float = true; x.inf_bits = 0x7C00ull; break;
        case B200_BF16: is_float = true; x.inf_bits = 0x7F80ull; break;
        case B200_F32: is_float = true; x.inf_bits = 0x7F800000ull; break;
        case B200_F64: is_float = true; x.inf_bits = 0x7FF0000000000000ull; break;
        case B200_I8: case B200_I16: case B200_I32: case B200_I64: x.base_enc = sign; x.base_dec = sign; break;
        case B200_U8: case B200_U16: case B200_U32: case B200_U64: case B200_BOOL: break;
        default: return false;
    }
    if (is_float) { x.base_enc = sign; x.base_dec = all; x.fmn = all & ~sign; x.mag_mask = all & ~sign; x.nan_or = all; }
    return true;
}

// sortperm's last step on every rank: tagged 32-bit source indices -> global 1-based Int64 indices
__global__ void __launch_bounds__(256) mgpu_widen_index_kernel(const unsigned int *__restrict__ in, long long *__restrict__ out,
                                                               int64_t n, int shift, const long long *__restrict__ offsets) {
    __shared__ long long s_off[PT_MAXB];
    if (threadIdx.x < PT_MAXB) s_off[threadIdx.x] = (int)threadIdx.x < (1 << (32 - shift)) && shift < 32 ? offsets[threadIdx.x] : 0;
    __syncthreads();
    const unsigned int low = shift >= 32 ? 0xFFFFFFFFu : ((1u << shift) - 1u);
    const int64_t stride = (int64_t)gridDim.x * 256 * 4;
    for (int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const Vec16 v = ldg_stream16(in + i);
            const unsigned int w[4] = {v.w[0], v.w[1], v.w[2], v.w[3]};
            longlong2 o[2];
            o[0].x = s_off[shift >= 32 ? 0 : (w[0] >> shift)] + (w[0] & low) + 1;
            o[0].y = s_off[shift >= 32 ? 0 : (w[1] >> shift)] + (w[1] & low) + 1;
            o[1].x = s_off[shift >= 32 ? 0 : (w[2] >> shift)] + (w[2] & low) + 1;
            o[1].y = s_off[shift >= 32 ? 0 : (w[3] >> shift)] + (w[3] & low) + 1;
            reinterpret_cast<longlong2 *>(out + i)[0] = o[0];
            reinterpret_cast<longlong2 *>(out + i)[1] = o[1];
        } else {
            for (int64_t j = i; j < n; ++j) {
                const unsigned int w = in[j];
                out[j] = s_off[shift >= 32 ? 0 : (w >> shift)] + (w & low) + 1;
            }
        }
    }
}

}  // namespace b200

extern "C" {

int32_t b200_partition_workspace_bytes(int64_t n, int32_t nbuckets, size_t *bytes) {
    using namespace b200;
    if (!bytes || n < 0 || nbuckets < 1 || nbuckets > PT_MAXB) return B200_INVALID_VALUE;
    *bytes = 0;
    if (n == 0) return B200_OK;
    PartParams p{};
    p.n = n; p.G = nbuckets;
    return run_partition<uint32_t>(nullptr, 0, p, 0, 3, nullptr, bytes);
}

static int32_t partition_common(void *workspace, size_t workspace_bytes, const void *keys_in, const void *payload_in,
                                int32_t dtype_key, int64_t n, const void *splitters, int32_t nbuckets,
                                int32_t descending, void *bucket_counts_dev, int phase, void *const *dst_keys,
                                void *const *dst_payload, const int64_t *dst_base, int local_bases,
                                b200_stream_t stream, const void *plan_dev = nullptr, int gen_index = 0,
                                uint32_t index_tag = 0) {
    using namespace b200;
    if (n < 0 || nbuckets < 1 || nbuckets > PT_MAXB || !bucket_counts_dev) return B200_INVALID_VALUE;
    SortXform xf; int kb;
    if (!part_xform(dtype_key, descending, xf, kb)) return B200_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (phase & 1) B200_CUDA_TRY(cudaMemsetAsync(bucket_counts_dev, 0, sizeof(long long) * nbuckets, st));
        return B200_OK;
    }
    if (!keys_in || (nbuckets > 1 && !splitters && !plan_dev)) return B200_INVALID_VALUE;
    if (plan_dev && kb != 4) return B200_UNSUPPORTED;
    PartParams p{};
    p.plan = static_cast<const b200_mgpu_plan_t *>(plan_dev);
    p.keys_in = keys_in; p.vals_in = payload_in;
    p.splitters = splitters; p.bucket_counts = static_cast<long long *>(bucket_counts_dev);
    p.n = n; p.G = nbuckets; p.xf = xf; p.local_bases = local_bases;
    const bool has_v = payload_in != nullptr || gen_index;
    p.index_tag = index_tag;
    if (phase & 2) {
        if (!dst_keys || (has_v && !dst_payload)) return B200_INVALID_VALUE;
        for (int b = 0; b < nbuckets; ++b) {
            if (!dst_keys[b] || (has_v && !dst_payload[b])) return B200_INVALID_VALUE;
            p.dst_keys[b] = dst_keys[b];
            p.dst_vals[b] = has_v ? dst_payload[b] : nullptr;
            p.dst_base[b] = dst_base ? dst_base[b] : 0;
        }
    }
    return run_partition_kb(kb, workspace, workspace_bytes, p, gen_index ? 2 : (has_v ? 1 : 0), phase, st, (size_t *)nullptr);
}

int32_t b200_partition(void *workspace, size_t workspace_bytes, const void *keys_in, void *keys_out,
                       const void *payload_in, void *payload_out, int32_t dtype_key, int64_t n,
                       const void *splitters, int32_t nbuckets, int32_t descending, void *bucket_counts_dev,
                       b200_stream_t stream) {
    if (n > 0 && (!keys_out || ((payload_in == nullptr) != (payload_out == nullptr)))) return B200_INVALID_VALUE;
    void *dk[16], *dv[16];
    for (int b = 0; b < 16; ++b) { dk[b] = keys_out; dv[b] = payload_out; }
    return partition_common(workspace, workspace_bytes, keys_in, payload_in, dtype_key, n, splitters, nbuckets, descending,
                            bucket_counts_dev, 3, dk, dv, nullptr, 1, stream);
}

int32_t b200_partition_count(void *workspace, size_t workspace_bytes, const void *keys_in, int32_t dtype_key, int64_t n,
                             const void *splitters, int32_t nbuckets, int32_t descending, void *bucket_counts_dev,
                             b200_stream_t stream) {
    return partition_common(workspace, workspace_bytes, keys_in, nullptr, dtype_key, n, splitters, nbuckets, descending,
                            bucket_counts_dev, 1, nullptr, nullptr, nullptr, 0, stream);
}

int32_t b200_partition_scatter(void *workspace, size_t workspace_bytes, const void *keys_in, const void *payload_in,
                               int32_t dtype_key, int64_t n, const void *splitters, int32_t nbuckets,
                               int32_t descending, void *bucket_counts_dev, void *const *dst_keys,
                               void *const *dst_payload, const int64_t *dst_base, b200_stream_t stream) {
    return partition_common(workspace, workspace_bytes, keys_in, payload_in, dtype_key, n, splitters, nbuckets, descending,
                            bucket_counts_dev, 2, dst_keys, dst_payload, dst_base, 0, stream);
}

// Stable partition of (keys, 8-byte payload) by the plan's 16-bit-bin splitters, every bucket stored straight into
// its destination rank's receive buffers at the plan's offsets (sortperm's exchange; keys-only sorts use
// b200_mgpu_scatter).  Workspace: b200_partition_workspace_bytes(n, nranks) + 256.
int32_t b200_mgpu_partition_pairs(void *workspace, size_t workspace_bytes, const void *keys, const void *payload,
                                  int32_t dtype_key, int64_t n, int32_t descending, const void *plan_dev,
                                  void *const *dst_keys, void *const *dst_payload, int32_t nranks, b200_stream_t stream) {
    if (!plan_dev || !workspace || workspace_bytes < 256) return B200_INVALID_VALUE;
    // the first 256 bytes of the workspace hold the bucket totals the count phase writes
    return partition_common(static_cast<char *>(workspace) + 256, workspace_bytes - 256, keys, payload, dtype_key, n, nullptr,
                            nranks, descending, workspace, 3, dst_keys, dst_payload, nullptr, 0, stream, plan_dev);
}

// sortperm's exchange without a payload to read: the payload is GENERATED as (rank << shift) | source index, 4 bytes
// (needs nranks <= 2^(32 - shift) and n <= 2^shift).  dst_index: per-rank UInt32 receive buffers.
int32_t b200_mgpu_partition_index(void *workspace, size_t workspace_bytes, const void *keys, int32_t dtype_key, int64_t n,
                                  int32_t descending, const void *plan_dev, int32_t rank, int32_t shift,
                                  void *const *dst_keys, void *const *dst_index, int32_t nranks, b200_stream_t stream) {
    if (!plan_dev || !workspace || workspace_bytes < 256) return B200_INVALID_VALUE;
    if (shift < 1 || shift > 32 || rank < 0 || rank >= nranks || n > (1ll << shift)) return B200_INVALID_VALUE;
    if (shift < 32 ? ((int64_t)nranks > (1ll << (32 - shift))) : nranks > 1) return B200_INVALID_VALUE;
    const uint32_t tag = shift < 32 ? ((uint32_t)rank << shift) : 0u;
    return partition_common(static_cast<char *>(workspace) + 256, workspace_bytes - 256, keys, nullptr, dtype_key, n, nullptr,
                            nranks, descending, workspace, 3, dst_keys, dst_index, nullptr, 0, stream, plan_dev, 1, tag);
}

// out[i] = offsets[index[i] >> shift] + (index[i] & (2^shift - 1)) + 1   (global 1-based Int64 indices); offsets: Int64[16]
// on the device (entries past the rank count are not read when the tags are valid, but must be readable).
int32_t b200_mgpu_widen_index(const void *index_u32, int64_t n, int32_t shift, const void *offsets_dev, void *out_i64,
                              b200_stream_t stream) {
    using namespace b200;
    if (n < 0 || shift < 1 || shift > 32) return B200_INVALID_VALUE;
    if (n == 0) return B200_OK;
    if (!index_u32 || !offsets_dev || !out_i64) return B200_INVALID_VALUE;
    const int64_t blocks = std::min<int64_t>(cdiv(n, 1024), 148 * 16);
    mgpu_widen_index_kernel<<<(unsigned int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const unsigned int *>(index_u32), static_cast<long long *>(out_i64), n, shift,
        static_cast<const long long *>(offsets_dev));
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // extern "C"
