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

// Sampled fast path (large n): bounds [lo, hi] from a sorted sample bracket the k-th key with overwhelming
// probability; ONE streaming read counts the keys below lo and compacts the keys inside the bracket (a few percent of
// n); the byte passes then run over the candidates only.  `fast` says whether the bracket held rank k (else the
// streaming passes over the whole input run, gated on the same word).
struct SelectFast {
    unsigned long long lo, hi;          // bracket, ordered-key space, inclusive
    unsigned long long below;           // keys < lo
    unsigned long long ncand;           // keys in [lo, hi] (may exceed the buffer: then the fast path is off)
    unsigned long long n_eff;           // min(ncand, cap) once decided
    unsigned int fast;                  // 1: candidates hold rank k
    unsigned int pad;
};

// gate_want: 0 = always run; 1 = run only if fast->fast == 1; 2 = run only if fast->fast == 0.
// n_dev != nullptr: the element count is read from device memory (candidate passes).
template <class K>
__global__ void __launch_bounds__(512) select_hist_kernel(const K *__restrict__ keys, int64_t n, SortXform xf, int shift,
                                                          const SelectState *__restrict__ state,
                                                          unsigned int *__restrict__ ghist,
                                                          const SelectFast *__restrict__ fast = nullptr, int gate_want = 0,
                                                          bool n_from_fast = false) {
    constexpr int VN = 16 / (int)sizeof(K);
    __shared__ unsigned int sh[RX_RADIX * RX_HIST_REPL];
    if (gate_want == 1 && fast->fast != 1u) return;
    if (gate_want == 2 && fast->fast != 0u) return;
    if (n_from_fast) n = (int64_t)fast->n_eff;
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
                                                               SortXform xf, K *__restrict__ out,
                                                               const SelectFast *__restrict__ fast = nullptr, int gate_want = 0) {
    __shared__ unsigned long long s[RX_RADIX];
    if (gate_want == 1 && fast->fast != 1u) return;
    if (gate_want == 2 && fast->fast != 0u) return;
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

// ---- sampled bracket: one CTA gathers S evenly spaced keys, sorts them (bitonic, shared memory) and brackets rank k
constexpr int SEL_SAMPLE = 4096;
template <class K>
__global__ void __launch_bounds__(1024) select_sample_kernel(const K *__restrict__ keys, int64_t n, unsigned long long k,
                                                             SortXform xf, SelectFast *fast) {
    __shared__ K s[SEL_SAMPLE];
    const int tid = threadIdx.x;
    for (int i = tid; i < SEL_SAMPLE; i += 1024) {
        const int64_t idx = (int64_t)(((unsigned __int128)(unsigned long long)i * (unsigned long long)n) / SEL_SAMPLE);
        s[i] = key_encode<K>(keys[idx], xf);
    }
    __syncthreads();
    for (int size = 2; size <= SEL_SAMPLE; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < SEL_SAMPLE / 2; t += 1024) {
                const int lo = 2 * t - (t & (stride - 1));          // index of the lower element of pair t
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const K a = s[lo], b = s[hi];
                if ((a > b) == up) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        // sample rank of the k-th key: q = k S / n, standard deviation sqrt(S p (1 - p)) <= 32; bracket = 5 sigma + slack
        const double p = (double)k / (double)n;
        const double sd = sqrt((double)SEL_SAMPLE * p * (1.0 - p));
        const int q = (int)(p * SEL_SAMPLE);
        const int delta = (int)(4.5 * sd) + 6;
        const int a = q - delta, b = q + delta;
        fast->lo = a <= 0 ? 0ull : (unsigned long long)s[a];
        fast->hi = b >= SEL_SAMPLE - 1 ? ~0ull : (unsigned long long)s[b];
        fast->below = 0; fast->ncand = 0; fast->n_eff = 0; fast->fast = 0; fast->pad = 0;
    }
}

// ONE streaming read: count the keys below the bracket, compact (ordered-key space) the keys inside it.
// Every LANE stages its candidates in its own 64-byte stack in shared memory (a predicated store and an add per key:
// no ballot, no atomic, no barrier); when some lane's stack is half full the warp flushes all 32 stacks with one scan and
// one global atomic.  Measured alternatives at 2^30 Float32: a warp-aggregated global atomic per 32 keys — 22 ms for a
// median bracket (30 M serialised L2 atomics on one word); one atomic per CTA tile behind two barriers — 1.9 ms with an
// empty bracket (barrier-bound); warp-level staging by ballot + popc — 1.35 ms (issue-bound, ~25 instructions per key).
template <class K>
__global__ void __launch_bounds__(512) select_filter_kernel(const K *__restrict__ keys, int64_t n, SortXform xf,
                                                            SelectFast *fast, K *__restrict__ cand, unsigned long long cap) {
    constexpr int VN = 16 / (int)sizeof(K);
    constexpr int NW = 512 / 32;
    constexpr int CAPL = 64 / (int)sizeof(K);                      // entries per lane; one iteration adds at most 2 VN = CAPL / 2
    __shared__ unsigned long long s_below[NW];
    __shared__ K s_stage[CAPL * 512];                              // [slot][thread]: a warp's stores to one slot are consecutive
    const K lo = (K)fast->lo, hi = (K)fast->hi;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    K *stage = s_stage + threadIdx.x;
    unsigned int below = 0;
    unsigned int fill = 0;
    auto flush = [&]() {
        unsigned int incl = fill;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        unsigned long long base = 0;
        if (lane == 31) base = atomicAdd(&fast->ncand, (unsigned long long)incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - fill;
        for (unsigned int i = 0; i < fill; ++i)
            if (base + i < cap) cand[base + i] = stage[i * 512];
        fill = 0;
    };
    auto take = [&](K raw, bool live) {
        const K key = key_encode<K>(raw, xf);
        below += (live && key < lo) ? 1u : 0u;
        if (live && key >= lo && key <= hi) { stage[fill * 512] = key; ++fill; }
    };
    const bool aligned = (reinterpret_cast<uintptr_t>(keys) % 16) == 0;
    const int64_t nvec = aligned ? n / VN : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // every lane of a warp runs the same iterations (the flush is a warp operation): warp-uniform loop bounds
    const int64_t warp_v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane;
    for (int64_t w = warp_v0; w < nvec; w += 2 * stride) {
        const int64_t v = w + lane, v2 = v + stride;
        Vec16 r0 = {}, r1 = {};
        const bool l0 = v < nvec, l1 = v2 < nvec;
        if (l0) r0 = ldg_stream16(reinterpret_cast<const char *>(keys) + v * 16);
        if (l1) r1 = ldg_stream16(reinterpret_cast<const char *>(keys) + v2 * 16);
        K e[2 * VN];
        memcpy(e, &r0, 16);
        memcpy(e + VN, &r1, 16);
#pragma unroll
        for (int q = 0; q < VN; ++q) take(e[q], l0);
#pragma unroll
        for (int q = 0; q < VN; ++q) take(e[VN + q], l1);
        if (__any_sync(0xffffffffu, fill > (unsigned int)(CAPL / 2))) flush();
    }
    for (int64_t w = nvec * VN + warp_v0; w < n; w += stride) {     // unaligned input / tail: one key per thread
        const int64_t i = w + lane;
        const bool live = i < n;
        take(live ? keys[i] : K(0), live);
        if (__any_sync(0xffffffffu, fill > (unsigned int)(CAPL / 2))) flush();
    }
    if (__any_sync(0xffffffffu, fill != 0)) flush();
    unsigned long long below64 = below;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below64 += __shfl_xor_sync(0xffffffffu, below64, o);
    if (lane == 0) s_below[warp] = below64;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < NW; ++w) t += s_below[w];
        if (t) atomicAdd(&fast->below, t);
    }
}

// Did the bracket hold rank k?  Seeds the candidate passes' state.
static __global__ void select_decide_kernel(SelectFast *fast, SelectState *state, unsigned long long k, unsigned long long cap) {
    const unsigned long long below = fast->below, nc = fast->ncand;
    const bool ok = nc <= cap && k >= below && k < below + nc;
    fast->fast = ok ? 1u : 0u;
    fast->n_eff = ok ? nc : 0ull;
    state->k = ok ? k - below : k;
    state->prefix = 0;
    state->mask = 0;
}

}  // namespace b200
