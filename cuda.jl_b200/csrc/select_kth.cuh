// select_kth.cuh — radix select: the k-th smallest key (isless order, `rev` as key complement) without
// sorting.  SURVEY §8 f2: the reference's partialsort (bitonic alg) is "full sort, then c[k]"
// (CUDACore/src/sorting.jl:1015-1020); for a scalar k only the value is needed.
//
// MSD over the key bytes: pass p histograms byte (P-1-p) of the keys that still match the prefix found so
// far (all keys in pass 0), a one-CTA kernel then picks the bin holding rank k, extends the prefix and
// rebases k.  sizeof(K) streaming reads of the keys, nothing written but 256-bin histograms; the input is
// left untouched.  No host round trip: prefix / mask / k live in device memory (graph-capturable).
#pragma once
#include "sort.cuh"

namespace b200 {

struct SelectState {
    unsigned long long k;        // rank still to be located inside the current candidate set
    unsigned long long prefix;   // key bits fixed so far (ordered-key space)
    unsigned long long mask;     // which bits of `prefix` are fixed
};

template <class K>
__global__ void __launch_bounds__(512) select_hist_kernel(const K *__restrict__ keys, int64_t n, SortXform xf, int shift,
                                                          const SelectState *__restrict__ state,
                                                          unsigned int *__restrict__ ghist) {
    constexpr int VN = 16 / (int)sizeof(K);
    __shared__ unsigned int sh[RX_RADIX * RX_HIST_REPL];
    for (int i = threadIdx.x; i < RX_RADIX * RX_HIST_REPL; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const K prefix = (K)state->prefix, mask = (K)state->mask;
    unsigned int *mine = sh + (threadIdx.x & (RX_HIST_REPL - 1));
    auto count = [&](K raw) {
        const K key = key_encode<K>(raw, xf);
        if ((key & mask) == prefix) atomicAdd(mine + (int)((key >> shift) & 0xFF) * RX_HIST_REPL, 1u);
    };
    const bool aligned = (reinterpret_cast<uintptr_t>(keys) % 16) == 0;
    const int64_t nvec = aligned ? n / VN : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; v + stride < nvec; v += 2 * stride) {          // two independent 16-byte loads in flight per thread
        const Vec16 r0 = ldg_stream16(reinterpret_cast<const char *>(keys) + v * 16);
        const Vec16 r1 = ldg_stream16(reinterpret_cast<const char *>(keys) + (v + stride) * 16);
        K e[2 * VN];
        memcpy(e, &r0, 16);
        memcpy(e + VN, &r1, 16);
#pragma unroll
        for (int k = 0; k < 2 * VN; ++k) count(e[k]);
    }
    for (; v < nvec; v += stride) {
        const Vec16 r0 = ldg_stream16(reinterpret_cast<const char *>(keys) + v * 16);
        K e[VN];
        memcpy(e, &r0, 16);
#pragma unroll
        for (int k = 0; k < VN; ++k) count(e[k]);
    }
    for (int64_t i = nvec * VN + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) count(keys[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < RX_RADIX; i += blockDim.x) {
        unsigned int c = 0;
#pragma unroll
        for (int r = 0; r < RX_HIST_REPL; ++r) c += sh[i * RX_HIST_REPL + r];
        if (c) atomicAdd(ghist + i, c);
    }
}

// One CTA of 256 threads: find the bin d with  sum(hist[0..d)) <= k < sum(hist[0..d]),  fix that byte.
// first: seed the state (k0, empty prefix) instead of reading it; last: write the decoded key to `out`.
template <class K>
__global__ void __launch_bounds__(RX_RADIX) select_pick_kernel(const unsigned int *__restrict__ ghist, SelectState *state,
                                                               int shift, int first, unsigned long long k0, int last,
                                                               SortXform xf, K *__restrict__ out) {
    __shared__ unsigned long long s[RX_RADIX];
    const int d = threadIdx.x;
    const unsigned long long c = ghist[d];
    s[d] = c;
    __syncthreads();
    for (int o = 1; o < RX_RADIX; o <<= 1) {
        const unsigned long long v = d >= o ? s[d - o] : 0ull;
        __syncthreads();
        s[d] += v;
        __syncthreads();
    }
    const unsigned long long k = first ? k0 : state->k;
    const unsigned long long prefix = first ? 0ull : state->prefix;
    const unsigned long long mask = first ? 0ull : state->mask;
    __syncthreads();                                         // everyone has read the state before it changes
    const unsigned long long incl = s[d], excl = incl - c;
    if (k >= excl && k < incl) {                             // exactly one thread (k < candidate count)
        const unsigned long long np = prefix | ((unsigned long long)d << shift);
        state->k = k - excl;
        state->prefix = np;
        state->mask = mask | (0xFFull << shift);
        if (last) out[0] = key_decode<K>((K)np, xf);
    }
}

}  // namespace b200
