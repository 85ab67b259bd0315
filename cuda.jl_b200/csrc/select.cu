// select.cu — findall(::CuArray{Bool}): fused scan + compact in ONE pass.
//
// The reference computes `indices = cumsum(bools)`, reads `indices[end]` back to the host,
// allocates the result and launches a scatter kernel (CUDACore/src/indexing.jl:26-66): three
// passes over N plus a device->host sync.  Here every tile counts its flags, resolves its
// output offset with the same decoupled look-back as the scan, and writes its 1-based linear
// indices directly; the total is left in device memory for the caller (who sized `indices`
// for the worst case or slices it afterwards).  SURVEY §8 f1.
#include "scan.cuh"

namespace b200 {

constexpr int SEL_ITEMS = 16;                    // flags per thread = one 16-byte load
constexpr int SEL_TILE = SC_THREADS * SEL_ITEMS;

template <class IndexT>
__global__ void __launch_bounds__(SC_THREADS) select_flagged_kernel(ScanParams p, long long *count_out) {
    using Op = OpAdd<long long>;
    __shared__ long long s_warp[SC_WARPS];
    __shared__ long long s_tile_prefix;
    __shared__ unsigned long long s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1ull);
    __syncthreads();
    const int64_t t = (int64_t)s_ticket;
    const int64_t base = t * SEL_TILE;
    int64_t valid = p.n - base;
    if (valid > SEL_TILE) valid = SEL_TILE;
    const uint8_t *flags = static_cast<const uint8_t *>(p.in) + base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int off = threadIdx.x * SEL_ITEMS;

    uint8_t f[SEL_ITEMS];
    if (off + SEL_ITEMS <= valid && (reinterpret_cast<uintptr_t>(flags + off) % 16) == 0) {
        Vec16 raw = ldg_stream16(flags + off);
        memcpy(f, &raw, 16);
    } else {
#pragma unroll
        for (int k = 0; k < SEL_ITEMS; ++k) f[k] = (off + k < valid) ? ldg_stream(flags + off + k) : (uint8_t)0;
    }
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; ++k) cnt += f[k] != 0;

    // block exclusive scan of the per-thread counts
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = lane < SC_WARPS ? s_warp[lane] : 0;
        long long wi = w;
#pragma unroll
        for (int d = 1; d < SC_WARPS; d <<= 1) {
            const long long up = shfl_up_any(wi, d);
            if (lane >= d) wi += up;
        }
        const long long tile_total = shfl_idx_any(wi, SC_WARPS - 1);
        const long long prefix = tile_lookback<Op>(p, 0, t, lane, tile_total);
        if (lane == 0) {
            s_tile_prefix = prefix;
            if (t == p.tiles_per_seg - 1) *count_out = prefix + tile_total;   // the last tile knows the total
        }
        if (lane < SC_WARPS) s_warp[lane] = wi - w;
    }
    __syncthreads();
    long long dst = s_tile_prefix + s_warp[warp] + (incl - cnt);
    IndexT *out = static_cast<IndexT *>(p.out);
#pragma unroll
    for (int k = 0; k < SEL_ITEMS; ++k) {
        if (f[k] != 0) out[dst++] = (IndexT)(base + off + k + 1);   // 1-based linear index
    }
}

}  // namespace b200

extern "C" {

int32_t b200_select_workspace_bytes(int64_t n, size_t *bytes) {
    using namespace b200;
    if (!bytes || n < 0) return B200_INVALID_VALUE;
    Carver c(nullptr);
    c.take<unsigned long long>(1);
    c.take<uint64_t>((size_t)cdiv(n, SEL_TILE) * StatusWords<long long>::NW);
    *bytes = n == 0 ? 0 : c.bytes();
    return B200_OK;
}

int32_t b200_select_flagged(void *workspace, size_t workspace_bytes, const void *flags_bool, void *indices,
                            int32_t dtype_index, void *count_dev, int64_t n, b200_stream_t stream) {
    using namespace b200;
    if (n < 0 || !count_dev) return B200_INVALID_VALUE;
    if (dtype_index != B200_I32 && dtype_index != B200_I64) return B200_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n == 0) {
        B200_CUDA_TRY(cudaMemsetAsync(count_dev, 0, sizeof(long long), st));
        return B200_OK;
    }
    if (!flags_bool || !indices) return B200_INVALID_VALUE;
    if (dtype_index == B200_I32 && n > 0x7FFFFFFFLL) return B200_INVALID_VALUE;
    const int64_t tiles = cdiv(n, SEL_TILE);
    if (tiles > 2147483647LL) return B200_INVALID_VALUE;
    Carver c(workspace);
    unsigned long long *ticket = c.take<unsigned long long>(1);
    uint64_t *status = c.take<uint64_t>((size_t)tiles * StatusWords<long long>::NW);
    if (c.bytes() > workspace_bytes) return B200_WORKSPACE_TOO_SMALL;
    if (!workspace) return B200_INVALID_VALUE;
    ScanParams p{};
    p.in = flags_bool; p.out = indices; p.status = status; p.ticket = ticket;
    p.pre = 1; p.n = n; p.post = 1; p.tiles_per_seg = tiles;
    B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, c.bytes(), st));
    if (dtype_index == B200_I64)
        select_flagged_kernel<long long><<<(unsigned int)tiles, SC_THREADS, 0, st>>>(p, static_cast<long long *>(count_dev));
    else
        select_flagged_kernel<int><<<(unsigned int)tiles, SC_THREADS, 0, st>>>(p, static_cast<long long *>(count_dev));
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // extern "C"
