// ref_algos.cu — "R" baseline of BASELINE.md §2: the REFERENCE'S ALGORITHMS restated in CUDA C++
// so that they can be timed on the same B200 next to the graft (the reference itself is Julia and
// cannot run in this image).  Restatement of the algorithm and launch structure, nvcc-compiled,
// NOT Julia-JIT code and not a line-by-line port.  BASELINE ONLY: built into
// bench/baselines/libref_algos.so, loaded by bench.py / tools, never linked into libb200arrays.so.
//
//   reduce  partial_mapreduce_grid + reduce_block/reduce_warp + second pass on `partial`
//           (CUDACore/src/mapreduce.jl:7-50, :90-134, :259-298): one block (or reduce_blocks blocks)
//           per output element, scalar loads, thread t takes reduced index t, t+T, ...; shuffle tree;
//           the multi-block case allocates partials and launches the kernel again.
//   scan    partial_scan (Blelloch up/down-sweep, one element per thread, 2*threads smem) +
//           aggregate copy + recursive scan of the block totals + aggregate_partial_scan
//           (CUDACore/src/accumulate.jl:15-131, :166-196).
//   sort    bitonic network: global comparator kernel for spans wider than a block, one
//           shared-memory kernel per stage for the remaining steps
//           (CUDACore/src/sorting.jl:694-714, :860-891, host loop :939-971); n must be a power of two here.
#include <cuda_runtime.h>
#include <stdint.h>

#define EXPORT extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------ reduce (Float32, op = +)
__device__ __forceinline__ float ref_reduce_block(float val) {
    __shared__ float shared[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) val += __shfl_down_sync(0xffffffffu, val, o);   // reduce_warp: offsets 1,2,4,8,16
    if (lane == 0) shared[wid] = val;
    __syncthreads();
    val = (threadIdx.x < (blockDim.x + 31) / 32) ? shared[lane] : 0.0f;
    if (wid == 0) for (int o = 1; o < 32; o <<= 1) val += __shfl_down_sync(0xffffffffu, val, o);
    return val;
}

// (pre, red, post) view; grid = reduce_blocks * (pre*post); partial layout [other][reduce_blocks]
__global__ void ref_partial_mapreduce_grid(const float *A, float *R, int64_t pre, int64_t red, int64_t post,
                                           int reduce_blocks) {
    const int64_t other = pre * post;
    const int64_t blk = blockIdx.x;
    const int64_t iother = blk % other, breduce = blk / other;      // fldmod1(blockIdx, length(Rother))
    const int64_t i = iother % pre, j = iother / pre;
    float val = 0.0f;                                                // op(neutral, neutral)
    int64_t r = threadIdx.x + breduce * blockDim.x;
    while (r < red) {
        val += A[i + pre * (r + red * j)];
        r += (int64_t)blockDim.x * reduce_blocks;
    }
    val = ref_reduce_block(val);
    if (threadIdx.x == 0) R[iother + other * breduce] = val;
}

EXPORT int ref_sum_f32(const float *A, float *R, float *partial, int64_t pre, int64_t red, int64_t post,
                       cudaStream_t st) {
    int dev = 0, sms = 148, bps = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int threads = 1024;                                               // launch_configuration: max block for this kernel
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, ref_partial_mapreduce_grid, threads, 0);
    const int64_t config_blocks = (int64_t)sms * bps;
    const int64_t other = pre * post;
    int64_t want = ((red + 31) / 32) * 32;                            // nextwarp(length(Rreduce))
    if (want < threads) threads = (int)want;
    int64_t reduce_blocks = 1;
    if (other < config_blocks) {
        const int64_t need = (red + threads - 1) / threads, fill = (config_blocks + other - 1) / other;
        reduce_blocks = need < fill ? need : fill;
    }
    if (reduce_blocks == 1) {
        ref_partial_mapreduce_grid<<<(unsigned)other, threads, 0, st>>>(A, R, pre, red, post, 1);
    } else {
        ref_partial_mapreduce_grid<<<(unsigned)(other * reduce_blocks), threads, 0, st>>>(A, partial, pre, red, post, (int)reduce_blocks);
        // second pass: mapreducedim!(identity, op, R, partial) with the reduced dimension = reduce_blocks
        int t2 = (int)(((reduce_blocks + 31) / 32) * 32);
        if (t2 > 1024) t2 = 1024;
        ref_partial_mapreduce_grid<<<(unsigned)other, t2, 0, st>>>(partial, R, other, reduce_blocks, 1, 1);
    }
    return (int)cudaGetLastError();
}
EXPORT int64_t ref_sum_partial_elems(int64_t pre, int64_t post) { return pre * post * 4096; }

// ------------------------------------------------------------------ scan (inclusive +)
template <class T>
__global__ void ref_partial_scan(const T *in, T *out, int64_t n) {
    extern __shared__ unsigned char smem_raw[];
    T *temp = reinterpret_cast<T *>(smem_raw);                        // 2*threads elements in the reference
    const int threads = blockDim.x, thread = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * threads + thread;
    temp[thread] = i < n ? in[i] : T(0);
    // up-sweep
    int offset = 1;
    for (int d = threads >> 1; d > 0; d >>= 1) {
        __syncthreads();
        if (thread < d) { const int ai = offset * (2 * thread + 1) - 1, bi = offset * (2 * thread + 2) - 1; temp[bi] += temp[ai]; }
        offset <<= 1;
    }
    if (thread == 0) temp[threads - 1] = T(0);
    // down-sweep
    for (int d = 1; d < threads; d <<= 1) {
        offset >>= 1;
        __syncthreads();
        if (thread < d) {
            const int ai = offset * (2 * thread + 1) - 1, bi = offset * (2 * thread + 2) - 1;
            const T t = temp[ai]; temp[ai] = temp[bi]; temp[bi] += t;
        }
    }
    __syncthreads();
    if (i < n) out[i] = temp[thread] + in[i];                          // exclusive -> inclusive by re-reading the input
}
template <class T>
__global__ void ref_gather_totals(const T *out, T *agg, int64_t n, int partial, int64_t nblocks) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblocks) { const int64_t last = (b + 1) * partial - 1; agg[b] = last < n ? out[last] : out[n - 1]; }
}
template <class T>
__global__ void ref_aggregate_partial_scan(T *out, const T *agg, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x > 0 && i < n) out[i] += agg[blockIdx.x - 1];
}
template <class T>
static void ref_scan_rec(const T *in, T *out, T *scratch, int64_t n, cudaStream_t st) {
    const int partial = 1024;
    const int64_t blocks = (n + partial - 1) / partial;
    ref_partial_scan<T><<<(unsigned)blocks, partial, 2 * partial * sizeof(T), st>>>(in, out, n);
    if (blocks == 1) return;
    T *agg = scratch;
    ref_gather_totals<T><<<(unsigned)((blocks + 255) / 256), 256, 0, st>>>(out, agg, n, partial, blocks);
    ref_scan_rec<T>(agg, agg, scratch + blocks, blocks, st);          // accumulate!(f, aggregates, aggregates)
    ref_aggregate_partial_scan<T><<<(unsigned)blocks, partial, 0, st>>>(out, agg, n);
}
EXPORT int ref_cumsum_f32(const float *in, float *out, float *scratch, int64_t n, cudaStream_t st) {
    ref_scan_rec<float>(in, out, scratch, n, st);
    return (int)cudaGetLastError();
}
EXPORT int ref_cumsum_i64(const long long *in, long long *out, long long *scratch, int64_t n, cudaStream_t st) {
    ref_scan_rec<long long>(in, out, scratch, n, st);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ bitonic sort (UInt32 keys, n = 2^k)
__global__ void ref_comparator_kernel(uint32_t *a, int64_t n, int64_t kk, int64_t jj) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t l = i ^ jj;
    if (l > i) {
        const uint32_t x = a[i], y = a[l];
        const bool up = (i & kk) == 0;
        if ((x > y) == up) { a[i] = y; a[l] = x; }
    }
}
__global__ void ref_comparator_small_kernel(uint32_t *a, int64_t n, int64_t kk, int64_t jj_start) {
    extern __shared__ uint32_t sm[];
    const int64_t base = (int64_t)blockIdx.x * blockDim.x;
    const int t = threadIdx.x;
    sm[t] = a[base + t];
    __syncthreads();
    for (int64_t jj = jj_start; jj > 0; jj >>= 1) {
        const int64_t i = base + t;
        const int l = t ^ (int)jj;
        if (l > t) {
            const uint32_t x = sm[t], y = sm[l];
            const bool up = (i & kk) == 0;
            if ((x > y) == up) { sm[t] = y; sm[l] = x; }
        }
        __syncthreads();
    }
    a[base + t] = sm[t];
}
EXPORT int ref_bitonic_sort_u32(uint32_t *a, int64_t n, cudaStream_t st) {
    const int block = 1024;
    if (n < block || (n & (n - 1))) return -1;
    for (int64_t kk = 2; kk <= n; kk <<= 1) {
        for (int64_t jj = kk >> 1; jj > 0; jj >>= 1) {
            if (jj < block) {          // the remaining steps of this stage fit a block: one smem kernel
                ref_comparator_small_kernel<<<(unsigned)(n / block), block, block * sizeof(uint32_t), st>>>(a, n, kk, jj);
                break;
            }
            ref_comparator_kernel<<<(unsigned)(n / block), block, 0, st>>>(a, n, kk, jj);
        }
    }
    return (int)cudaGetLastError();
}
