/*
 * oracle_cpu.c — C restatement of the CPU path the reference's tests compare against
 * (Base Julia on `Array`), for the grafted reduce / scan / sort primitives.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: linked by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs; never by the product library.
 *
 * PARITY UNPINNED against Julia itself: there is no julia binary in this image, and the
 * algorithms below live in Julia Base (external to /root/reference; any Julia 1.10-1.13,
 * .buildkite/pipeline.yml:8-66).  They are restated from Base's published algorithms and
 * anchored on the reference's call sites:
 *   sum / maximum / count(A; dims)   GPUArrays.mapreducedim! hook  CUDACore/src/mapreduce.jl:168
 *                                    compared via testf() against Base  test/helpers.jl:13-14
 *   cumsum / accumulate              Base._accumulate! hooks       CUDACore/src/accumulate.jl:204-216
 *   sort! / sortperm                 Base.sort!/sortperm! hooks    CUDACore/src/sorting.jl:1003-1065
 * Base algorithms restated:
 *   - mapreduce over all elements: pairwise, 1024-element sequential base case
 *     (Base.mapreduce_impl, pairwise_blocksize = 1024)
 *   - reductions along dims: Base._mapreducedim! — reducing the first dimension keeps a
 *     scalar accumulator per column and runs down it; otherwise the output vector is
 *     updated column by column (R[i] = op(R[i], A[i,j]))
 *   - cumsum of a float vector: Base.accumulate_pairwise! (blocksize 128); integers: serial
 *   - sort! of UInt32/Float32 vectors: Julia >= 1.9 dispatches large bit-type inputs to an
 *     LSD radix sort after mapping floats to ordered unsigned keys (isless order, NaNs last)
 *   - sortperm: stable merge sort of the index vector under isless (ties by index)
 * Validated against oracle/oracle.py (numpy) in tests/test_oracle_golden.py::test_c_restatement_matches_numpy_oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ helpers */
static inline float bf16_to_f32(uint16_t h) {
    uint32_t w = (uint32_t)h << 16;
    float f;
    memcpy(&f, &w, 4);
    return f;
}
static inline uint16_t f32_to_bf16(float f) { /* round to nearest even, NaN preserved */
    uint32_t w;
    memcpy(&w, &f, 4);
    if ((w & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((w >> 16) | 0x0040u);
    w += 0x7fffu + ((w >> 16) & 1u);
    return (uint16_t)(w >> 16);
}
/* Julia's max/min: NaN-propagating, -0.0 < +0.0 */
static inline float jl_maxf(float a, float b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return signbit(a) ? b : a;
    return a > b ? a : b;
}

/* ------------------------------------------------------------------ full reductions */
static float pairwise_sum_f32(const float *a, int64_t lo, int64_t hi) { /* [lo, hi) */
    if (hi - lo <= 1024) {
        float s = a[lo];
        for (int64_t i = lo + 1; i < hi; ++i) s += a[i];
        return s;
    }
    int64_t mid = lo + ((hi - lo) >> 1);
    return pairwise_sum_f32(a, lo, mid) + pairwise_sum_f32(a, mid, hi);
}
float ref_sum_f32(const float *a, int64_t n) { return n == 0 ? 0.0f : pairwise_sum_f32(a, 0, n); }

float ref_maximum_f32(const float *a, int64_t n) {
    float m = a[0];
    for (int64_t i = 1; i < n; ++i) m = jl_maxf(m, a[i]);
    return m;
}
float ref_maximum_bf16(const uint16_t *a, int64_t n) {
    float m = bf16_to_f32(a[0]);
    for (int64_t i = 1; i < n; ++i) m = jl_maxf(m, bf16_to_f32(a[i]));
    return m;
}
int64_t ref_count_bool(const uint8_t *a, int64_t n) {
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) c += a[i] != 0;
    return c;
}

/* ------------------------------------------------------------------ dims reductions (column-major m x n) */
void ref_sum_dims1_f32(const float *a, int64_t m, int64_t n, float *out) { /* out[j] = sum_i a[i,j] */
    for (int64_t j = 0; j < n; ++j) {
        const float *col = a + j * m;
        float r = 0.0f;
        for (int64_t i = 0; i < m; ++i) r += col[i];
        out[j] = r;
    }
}
void ref_sum_dims2_f32(const float *a, int64_t m, int64_t n, float *out) { /* out[i] = sum_j a[i,j] */
    for (int64_t i = 0; i < m; ++i) out[i] = 0.0f;
    for (int64_t j = 0; j < n; ++j) {
        const float *col = a + j * m;
        for (int64_t i = 0; i < m; ++i) out[i] += col[i];
    }
}
/* BFloat16 arithmetic rounds after every add (BFloat16s.jl: +(x,y) = BFloat16(Float32(x)+Float32(y))) */
void ref_sum_dims1_bf16(const uint16_t *a, int64_t m, int64_t n, uint16_t *out) {
    for (int64_t j = 0; j < n; ++j) {
        const uint16_t *col = a + j * m;
        uint16_t r = 0;
        for (int64_t i = 0; i < m; ++i) r = f32_to_bf16(bf16_to_f32(r) + bf16_to_f32(col[i]));
        out[j] = r;
    }
}
void ref_sum_dims2_bf16(const uint16_t *a, int64_t m, int64_t n, uint16_t *out) {
    for (int64_t i = 0; i < m; ++i) out[i] = 0;
    for (int64_t j = 0; j < n; ++j) {
        const uint16_t *col = a + j * m;
        for (int64_t i = 0; i < m; ++i) out[i] = f32_to_bf16(bf16_to_f32(out[i]) + bf16_to_f32(col[i]));
    }
}

/* ------------------------------------------------------------------ scans */
static float accumulate_pairwise_f32(const float *v, float *c, float s, int64_t i1, int64_t n) {
    /* Base.accumulate_pairwise: blocksize 128; returns the block total */
    if (n < 128) {
        float acc = v[i1];
        c[i1] = s + acc; /* op(s, v[i1]) */
        for (int64_t i = i1 + 1; i < i1 + n; ++i) {
            acc += v[i];
            c[i] = s + acc;
        }
        return acc;
    }
    int64_t n2 = n >> 1;
    float s1 = accumulate_pairwise_f32(v, c, s, i1, n2);
    float s2 = accumulate_pairwise_f32(v, c, s + s1, i1 + n2, n - n2);
    return s1 + s2;
}
void ref_cumsum_f32(const float *v, float *out, int64_t n) {
    if (n == 0) return;
    /* first element seeds the running sum, as in accumulate_pairwise! */
    out[0] = v[0];
    if (n > 1) accumulate_pairwise_f32(v, out, v[0], 1, n - 1);
}
void ref_cumsum_i64(const int64_t *v, int64_t *out, int64_t n) {
    uint64_t acc = 0;
    for (int64_t i = 0; i < n; ++i) {
        acc += (uint64_t)v[i];
        out[i] = (int64_t)acc;
    }
}

/* ------------------------------------------------------------------ sorts */
static inline uint32_t f32_key(uint32_t b) { /* isless order; all NaNs last */
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
static void radix_sort_u32(uint32_t *a, uint32_t *tmp, int64_t n) {
    for (int pass = 0; pass < 4; ++pass) {
        int64_t cnt[257] = {0};
        const int sh = 8 * pass;
        for (int64_t i = 0; i < n; ++i) cnt[((a[i] >> sh) & 0xff) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; ++i) tmp[cnt[(a[i] >> sh) & 0xff]++] = a[i];
        uint32_t *t = a; a = tmp; tmp = t;
    }
}
void ref_sort_u32(uint32_t *a, int64_t n) {
    uint32_t *tmp = (uint32_t *)malloc((size_t)n * 4);
    radix_sort_u32(a, tmp, n);
    free(tmp);
}
void ref_sort_f32(float *a, int64_t n) { /* NaN payloads are kept: keys carry |bits| for NaNs */
    uint32_t *k = (uint32_t *)a, *tmp = (uint32_t *)malloc((size_t)n * 4);
    for (int64_t i = 0; i < n; ++i) {
        uint32_t b = k[i];
        k[i] = ((b & 0x7fffffffu) > 0x7f800000u) ? (b | 0x80000000u) : ((b & 0x80000000u) ? ~b : (b | 0x80000000u));
    }
    radix_sort_u32(k, tmp, n);
    for (int64_t i = 0; i < n; ++i) {
        uint32_t key = k[i];
        k[i] = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
    }
    free(tmp);
}
/* stable merge sort of 1-based indices under isless(v[i], v[j]) */
static void msort_perm(const uint32_t *key, int64_t *ix, int64_t *tmp, int64_t lo, int64_t hi) {
    if (hi - lo <= 20) { /* insertion sort, stable */
        for (int64_t i = lo + 1; i < hi; ++i) {
            int64_t x = ix[i];
            int64_t j = i;
            while (j > lo && key[x - 1] < key[ix[j - 1] - 1]) { ix[j] = ix[j - 1]; --j; }
            ix[j] = x;
        }
        return;
    }
    int64_t mid = lo + ((hi - lo) >> 1);
    msort_perm(key, ix, tmp, lo, mid);
    msort_perm(key, ix, tmp, mid, hi);
    memcpy(tmp + lo, ix + lo, (size_t)(mid - lo) * sizeof(int64_t));
    int64_t i = lo, j = mid, k = lo;
    while (i < mid && j < hi) {
        if (key[ix[j] - 1] < key[tmp[i] - 1]) ix[k++] = ix[j++];
        else ix[k++] = tmp[i++];
    }
    while (i < mid) ix[k++] = tmp[i++];
}
void ref_sortperm_f32(const float *v, int64_t *perm, int64_t n) {
    uint32_t *key = (uint32_t *)malloc((size_t)n * 4);
    int64_t *tmp = (int64_t *)malloc((size_t)n * 8);
    const uint32_t *b = (const uint32_t *)v;
    for (int64_t i = 0; i < n; ++i) { key[i] = f32_key(b[i]); perm[i] = i + 1; }
    msort_perm(key, perm, tmp, 0, n);
    free(key);
    free(tmp);
}
void ref_sortperm_u32(const uint32_t *v, int64_t *perm, int64_t n) {
    int64_t *tmp = (int64_t *)malloc((size_t)n * 8);
    for (int64_t i = 0; i < n; ++i) perm[i] = i + 1;
    msort_perm(v, perm, tmp, 0, n);
    free(tmp);
}

/* Comparison-sort variant (SURVEY section 8d asks for BOTH a radix and a comparison baseline, because Julia >= 1.9
 * picks radix sort for large bit-type vectors while older versions used quick/merge sort). */
static int cmp_u32(const void *a, const void *b) {
    const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}
void ref_sort_f32_cmp(float *a, int64_t n) {
    uint32_t *k = (uint32_t *)a;
    for (int64_t i = 0; i < n; ++i) k[i] = f32_key(k[i]);
    qsort(k, (size_t)n, sizeof(uint32_t), cmp_u32);
    for (int64_t i = 0; i < n; ++i) { uint32_t e = k[i]; k[i] = (e & 0x80000000u) ? (e ^ 0x80000000u) : ~e; }
}

