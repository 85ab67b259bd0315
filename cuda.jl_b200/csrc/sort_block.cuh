// sort_block.cuh — segments of up to 4096 keys sorted by ONE CTA entirely on chip, and the
// batched transpose used for long strided segments.  SURVEY §8 f2: sort!(A; dims) /
// sortperm(A; dims) for any dims (reference: bitonic_sort! with dims, sorting.jl:909-974,
// tests test/core/sorting.jl:314-315,388-392).
//
// A segment is the set of elements (i, 0..n-1, j) of the (pre, n, post) view: base
// i + pre*n*j, stride pre.  The CTA loads it (warp-striped), runs sizeof(K) stable radix-256
// passes through shared memory with the same ballot ranking as the onesweep kernel, and writes
// the keys (sort!) or the 1-based indices (sortperm; linear indices for N-d input,
// sorting.jl:1067-1070) back to the same strided positions.
#pragma once
#include "sort.cuh"

namespace b200 {

constexpr int RB_ITEMS = 16;
constexpr int RB_TILE = RX_THREADS * RB_ITEMS;   // 4096

struct BlockSortParams {
    const void *keys_in;
    void *keys_out;      // sort!: same array as keys_in
    void *perm_out;      // sortperm
    int64_t pre, n, post;
    SortXform xf;
    int32_t index_dtype;
    int32_t linear;      // sortperm: 1-based linear index into the whole array instead of the segment
};

// ITEMS keys per thread: 16 (segments up to 4096) or 4 (up to 1024: a quarter of the ranking work, e.g. the
// reference's own sort(A; dims=1) benchmark on 512 x 1000, perf/array.jl:151)
template <class K, bool PERM, int ITEMS>
__global__ void __launch_bounds__(RX_THREADS) radix_block_sort_kernel(BlockSortParams p) {
    constexpr int TILE = RX_THREADS * ITEMS;
    constexpr int WARP_ITEMS = 32 * ITEMS;
    constexpr int P = (int)sizeof(K);
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned int *s_whist = reinterpret_cast<unsigned int *>(s_raw);                 // [8][256]
    K *s_keys = reinterpret_cast<K *>(s_whist + RX_WARPS * RX_RADIX);                // [TILE]
    unsigned short *s_vals = reinterpret_cast<unsigned short *>(s_keys + TILE);      // [TILE] (PERM)
    __shared__ unsigned int s_wsum[RX_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t seg = blockIdx.x;
    const int64_t j = seg / p.pre, i = seg - j * p.pre;
    const int64_t base = i + p.pre * p.n * j;
    const int64_t stride = p.pre;
    const int n = (int)p.n;
    const int idx0 = warp * WARP_ITEMS + lane;
    unsigned int *whist = s_whist + warp * RX_RADIX;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const int d = tid;

    const K *kin = static_cast<const K *>(p.keys_in) + base;
    K keys[ITEMS];
    unsigned short vals[PERM ? ITEMS : 1];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int idx = idx0 + it * 32;
        keys[it] = idx < n ? key_encode<K>(kin[(int64_t)idx * stride], p.xf) : (K)~(K)0;
        if constexpr (PERM) vals[it] = (unsigned short)idx;
    }

#pragma unroll 1
    for (int pass = 0; pass < P; ++pass) {
        const int shift = 8 * pass;
#pragma unroll
        for (int k = 0; k < RX_RADIX / 32; ++k) whist[k * 32 + lane] = 0;
        __syncwarp();
        unsigned short ranks[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const unsigned int digit = digit_of<K>(keys[it], shift);
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
        unsigned int incl = tile_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
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
        for (int it = 0; it < ITEMS; ++it) {
            const unsigned int pos = whist[digit_of<K>(keys[it], shift)] + ranks[it];
            s_keys[pos] = keys[it];
            if constexpr (PERM) s_vals[pos] = vals[it];
        }
        __syncthreads();
        if (pass + 1 < P) {
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                keys[it] = s_keys[idx0 + it * 32];
                if constexpr (PERM) vals[it] = s_vals[idx0 + it * 32];
            }
            __syncthreads();   // everyone has reloaded before the next pass rewrites s_whist / s_keys
        }
    }
    // padding (all-ones keys, highest indices) sorted last and stably: positions >= n
    if constexpr (!PERM) {
        K *kout = static_cast<K *>(p.keys_out) + base;
        for (int idx = tid; idx < n; idx += RX_THREADS) kout[(int64_t)idx * stride] = key_decode<K>(s_keys[idx], p.xf);
    } else {
        for (int idx = tid; idx < n; idx += RX_THREADS) {
            const long long r = s_vals[idx];
            const long long ix = p.linear ? 1 + base + r * stride : 1 + r;
            const int64_t dst = base + (int64_t)idx * stride;
            if (p.index_dtype == B200_I64 || p.index_dtype == B200_U64) static_cast<long long *>(p.perm_out)[dst] = ix;
            else static_cast<int *>(p.perm_out)[dst] = (int)ix;
        }
    }
}

// out[y + cols*(x + rows*b)] = in[x + rows*(y + cols*b)]: batched transpose of column-major
// rows x cols matrices through a padded shared-memory tile (coalesced on both sides).
template <class T>
__global__ void __launch_bounds__(256) transpose_batched_kernel(const T *__restrict__ in, T *__restrict__ out,
                                                                int64_t rows, int64_t cols) {
    __shared__ T tile[32][33];
    const int64_t b = blockIdx.z;
    const int64_t x0 = (int64_t)blockIdx.x * 32, y0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const T *inb = in + rows * cols * b;
    T *outb = out + rows * cols * b;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t x = x0 + tx, y = y0 + ty + k * 8;
        if (x < rows && y < cols) tile[ty + k * 8][tx] = inb[x + rows * y];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t y = y0 + tx, x = x0 + ty + k * 8;
        if (x < rows && y < cols) outb[y + cols * x] = tile[tx][ty + k * 8];
    }
}

// After a per-segment sortperm of the transposed keys: tmp[k + n*(i + pre*j)] is the 1-based
// position r+1 (inside the segment) of the k-th smallest element; write the linear (or
// per-segment) index to the strided position (i, k, j) of the output.
template <class IndexT>
__global__ void __launch_bounds__(256) perm_scatter_kernel(const long long *__restrict__ tmp, IndexT *__restrict__ out,
                                                           int64_t pre, int64_t n, int64_t post, int linear) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // index into tmp
    if (e >= pre * n * post) return;
    const int64_t k = e % n, s = e / n;
    const int64_t i = s % pre, j = s / pre;
    const long long r = tmp[e] - 1;
    const long long ix = linear ? 1 + i + pre * (r + n * j) : 1 + r;
    out[i + pre * (k + n * j)] = (IndexT)ix;
}

}  // namespace b200
