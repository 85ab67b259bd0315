// api_sort.cu — extern "C" entry points for sort / sortperm / pairs (include/b200arrays.h).
#include "sort_host.h"

namespace b200 {

static bool make_xform(int dtype, int descending, int nan_canonical, SortXform &x, int &kb) {
    x = SortXform{};
    kb = dtype_size(dtype);
    if (kb == 0) return false;
    const uint64_t all = kb == 8 ? ~0ull : ((1ull << (8 * kb)) - 1);
    const uint64_t sign = 1ull << (8 * kb - 1);
    x.sign = sign;
    x.desc = descending ? all : 0;
    x.inf_bits = all;   // never exceeded unless a float format narrows it below
    bool is_float = false;
    switch (dtype) {
        case B200_F16: is_float = true; x.inf_bits = 0x7C00ull; break;
        case B200_BF16: is_float = true; x.inf_bits = 0x7F80ull; break;
        case B200_F32: is_float = true; x.inf_bits = 0x7F800000ull; break;
        case B200_F64: is_float = true; x.inf_bits = 0x7FF0000000000000ull; break;
        case B200_I8: case B200_I16: case B200_I32: case B200_I64:
            x.base_enc = sign; x.base_dec = sign; break;
        case B200_U8: case B200_U16: case B200_U32: case B200_U64: case B200_BOOL: break;
        default: return false;
    }
    if (is_float) {
        x.base_enc = sign; x.base_dec = all; x.fmn = all & ~sign; x.mag_mask = all & ~sign;
        x.nan_or = nan_canonical ? all : 0;
    }
    return true;
}

static int32_t run(int kb, const SortCall &c) {
    switch (kb) {
        case 1: return sort_run_k1(c);
        case 2: return sort_run_k2(c);
        case 4: return sort_run_k4(c);
        case 8: return sort_run_k8(c);
        default: return B200_UNSUPPORTED;
    }
}

static bool index_dtype_ok(int d) { return d == B200_I32 || d == B200_I64 || d == B200_U32 || d == B200_U64; }

}  // namespace b200

extern "C" {

int32_t b200_sort_workspace_bytes(int32_t dtype_key, int32_t dtype_payload, int64_t n, int64_t nseg, size_t *bytes) {
    using namespace b200;
    if (!bytes || n < 0 || nseg < 0) return B200_INVALID_VALUE;
    *bytes = 0;
    SortXform xf; int kb;
    if (!make_xform(dtype_key, 0, 0, xf, kb)) return B200_UNSUPPORTED;
    if (n == 0 || nseg == 0) return B200_OK;
    SortCall c{};
    c.n = n; c.xf = xf; c.query = bytes;
    if (dtype_payload < 0) c.mode = SORT_KEYS;
    else if (dtype_payload == B200_SORTPERM_INDEX) c.mode = SORT_PERM;
    else {
        c.mode = SORT_PAIRS;
        c.payload_bytes = dtype_size(dtype_payload);
        if (c.payload_bytes != 4 && c.payload_bytes != 8) return B200_UNSUPPORTED;
    }
    return run(kb, c);   // segments are sorted one after the other and share the workspace
}

int32_t b200_sort_keys(void *workspace, size_t workspace_bytes, void *keys, int32_t dtype_key, int64_t n, int64_t nseg,
                       int32_t descending, b200_stream_t stream) {
    using namespace b200;
    if (n < 0 || nseg < 0) return B200_INVALID_VALUE;
    SortXform xf; int kb;
    if (!make_xform(dtype_key, descending, 0, xf, kb)) return B200_UNSUPPORTED;
    if (n <= 1 || nseg == 0) return B200_OK;
    if (!keys) return B200_INVALID_VALUE;
    for (int64_t s = 0; s < nseg; ++s) {
        SortCall c{};
        c.ws = workspace; c.ws_bytes = workspace_bytes;
        c.keys = static_cast<char *>(keys) + (size_t)s * n * kb;
        c.mode = SORT_KEYS; c.n = n; c.xf = xf;
        c.stream = reinterpret_cast<cudaStream_t>(stream);
        int32_t rc = run(kb, c);
        if (rc != B200_OK) return rc;
    }
    return B200_OK;
}

int32_t b200_sortperm(void *workspace, size_t workspace_bytes, const void *keys, void *perm, int32_t dtype_key,
                      int32_t dtype_index, int64_t n, int64_t nseg, int32_t descending, int32_t linear,
                      b200_stream_t stream) {
    using namespace b200;
    if (n < 0 || nseg < 0) return B200_INVALID_VALUE;
    SortXform xf; int kb;
    if (!make_xform(dtype_key, descending, 1, xf, kb)) return B200_UNSUPPORTED;
    if (!index_dtype_ok(dtype_index)) return B200_UNSUPPORTED;
    if (n == 0 || nseg == 0) return B200_OK;
    if (!keys || !perm) return B200_INVALID_VALUE;
    const int ib = dtype_size(dtype_index);
    if (ib == 4 && (linear ? n * nseg : n) > 0x7FFFFFFFLL) return B200_INVALID_VALUE;
    for (int64_t s = 0; s < nseg; ++s) {
        SortCall c{};
        c.ws = workspace; c.ws_bytes = workspace_bytes;
        c.keys = const_cast<char *>(static_cast<const char *>(keys)) + (size_t)s * n * kb;
        c.payload = static_cast<char *>(perm) + (size_t)s * n * ib;
        c.mode = SORT_PERM; c.n = n; c.xf = xf;
        c.index_dtype = dtype_index;
        c.index_base = 1 + (linear ? s * n : 0);
        c.stream = reinterpret_cast<cudaStream_t>(stream);
        int32_t rc = run(kb, c);
        if (rc != B200_OK) return rc;
    }
    return B200_OK;
}

int32_t b200_sort_pairs(void *workspace, size_t workspace_bytes, void *keys, void *payload, int32_t dtype_key,
                        int32_t dtype_payload, int64_t n, int32_t descending, b200_stream_t stream) {
    using namespace b200;
    if (n < 0) return B200_INVALID_VALUE;
    SortXform xf; int kb;
    if (!make_xform(dtype_key, descending, 0, xf, kb)) return B200_UNSUPPORTED;
    const int pb = dtype_size(dtype_payload);
    if (pb != 4 && pb != 8) return B200_UNSUPPORTED;
    if (n <= 1) return B200_OK;
    if (!keys || !payload) return B200_INVALID_VALUE;
    SortCall c{};
    c.ws = workspace; c.ws_bytes = workspace_bytes;
    c.keys = keys; c.payload = payload;
    c.mode = SORT_PAIRS; c.payload_bytes = pb; c.n = n; c.xf = xf;
    c.stream = reinterpret_cast<cudaStream_t>(stream);
    return run(kb, c);
}

int32_t b200_search_sorted(const void *sorted_keys, int64_t n, const void *needles, int64_t m, int32_t dtype_key,
                           int32_t descending, void *lower_bounds_dev, b200_stream_t stream) {
    using namespace b200;
    if (n < 0 || m < 0) return B200_INVALID_VALUE;
    SortXform xf; int kb;
    if (!make_xform(dtype_key, descending, 1, xf, kb)) return B200_UNSUPPORTED;
    if (m == 0) return B200_OK;
    if (!needles || !lower_bounds_dev || (n > 0 && !sorted_keys)) return B200_INVALID_VALUE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long *out = static_cast<long long *>(lower_bounds_dev);
    switch (kb) {
        case 1: return lower_bound_k1(sorted_keys, n, needles, m, xf, out, st);
        case 2: return lower_bound_k2(sorted_keys, n, needles, m, xf, out, st);
        case 4: return lower_bound_k4(sorted_keys, n, needles, m, xf, out, st);
        case 8: return lower_bound_k8(sorted_keys, n, needles, m, xf, out, st);
        default: return B200_UNSUPPORTED;
    }
}

}  // extern "C"
