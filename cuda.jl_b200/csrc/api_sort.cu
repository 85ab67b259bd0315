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

// ---- general (pre, n, post) machinery ------------------------------------------------
}  // extern "C"  (helpers below are C++)

namespace b200 {

static int32_t run_block(int kb, const BlockSortParams &bp, int perm, int64_t nseg, cudaStream_t st) {
    switch (kb) {
        case 1: return block_sort_k1(bp, perm, nseg, st);
        case 2: return block_sort_k2(bp, perm, nseg, st);
        case 4: return block_sort_k4(bp, perm, nseg, st);
        case 8: return block_sort_k8(bp, perm, nseg, st);
        default: return B200_UNSUPPORTED;
    }
}

template <class T>
static int32_t transpose_T(const void *in, void *out, int64_t rows, int64_t cols, int64_t batch, cudaStream_t st) {
    if (cdiv(cols, 32) > 65535) return B200_INVALID_VALUE;
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {          // gridDim.z is limited to 65535
        const int64_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        dim3 grid((unsigned int)cdiv(rows, 32), (unsigned int)cdiv(cols, 32), (unsigned int)nb);
        transpose_batched_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T *>(in) + b0 * rows * cols,
                                                          static_cast<T *>(out) + b0 * rows * cols, rows, cols);
        B200_CHECK_LAUNCH();
    }
    return B200_OK;
}
static int32_t transpose_bytes(int elem, const void *in, void *out, int64_t rows, int64_t cols, int64_t batch, cudaStream_t st) {
    switch (elem) {
        case 1: return transpose_T<uint8_t>(in, out, rows, cols, batch, st);
        case 2: return transpose_T<uint16_t>(in, out, rows, cols, batch, st);
        case 4: return transpose_T<uint32_t>(in, out, rows, cols, batch, st);
        case 8: return transpose_T<uint64_t>(in, out, rows, cols, batch, st);
        default: return B200_UNSUPPORTED;
    }
}

// mode: SORT_KEYS or SORT_PERM.  Strategy: n <= 4096 -> one CTA per segment (any stride);
// contiguous long segments -> device-wide onesweep per segment; strided long segments ->
// transpose, sort the now contiguous segments, transpose / scatter back.
static int32_t sort_dims(void *ws, size_t wsb, void *keys, void *perm, int dtype_key, int dtype_index, int mode,
                         int64_t pre, int64_t n, int64_t post, int descending, int linear, cudaStream_t st,
                         size_t *query) {
    SortXform xf; int kb;
    if (!make_xform(dtype_key, descending, mode == SORT_PERM ? 1 : 0, xf, kb)) return B200_UNSUPPORTED;
    if (mode == SORT_PERM && !index_dtype_ok(dtype_index)) return B200_UNSUPPORTED;
    if (query) *query = 0;
    const int64_t nseg = pre * post;
    if (n == 0 || nseg == 0) return B200_OK;
    if (mode == SORT_KEYS && n == 1) return B200_OK;
    const int ib = mode == SORT_PERM ? dtype_size(dtype_index) : 0;
    if (mode == SORT_PERM && ib == 4 && (linear ? n * nseg : n) > 0x7FFFFFFFLL) return B200_INVALID_VALUE;

    if (n <= RB_TILE) {
        if (query) return B200_OK;
        BlockSortParams bp{};
        bp.keys_in = keys; bp.keys_out = keys; bp.perm_out = perm;
        bp.pre = pre; bp.n = n; bp.post = post; bp.xf = xf; bp.index_dtype = dtype_index; bp.linear = linear;
        return run_block(kb, bp, mode == SORT_PERM, nseg, st);
    }
    // inner device-wide sort of one contiguous segment
    SortCall inner{};
    inner.mode = mode; inner.n = n; inner.xf = xf; inner.stream = st;
    inner.index_dtype = (pre == 1) ? dtype_index : B200_I64;
    size_t inner_bytes = 0;
    inner.query = &inner_bytes;
    int32_t rc = run(kb, inner);
    if (rc != B200_OK) return rc;
    inner.query = nullptr;

    Carver c(ws);
    char *inner_ws = c.take<char>(inner_bytes);
    char *tkeys = pre > 1 ? c.take<char>((size_t)n * nseg * kb) : nullptr;
    long long *tperm = (pre > 1 && mode == SORT_PERM) ? c.take<long long>((size_t)n * nseg) : nullptr;
    if (query) { *query = c.bytes(); return B200_OK; }
    if (c.bytes() > wsb) return B200_WORKSPACE_TOO_SMALL;
    if (ws == nullptr) return B200_INVALID_VALUE;
    inner.ws = inner_ws; inner.ws_bytes = inner_bytes;

    char *kbase = static_cast<char *>(keys);
    if (pre > 1) {   // (pre, n, post) -> (n, pre, post): segments become contiguous
        rc = transpose_bytes(kb, keys, tkeys, pre, n, post, st);
        if (rc != B200_OK) return rc;
        kbase = tkeys;
    }
    for (int64_t s = 0; s < nseg; ++s) {
        inner.keys = kbase + (size_t)s * n * kb;
        if (mode == SORT_PERM) {
            if (pre == 1) {
                inner.payload = static_cast<char *>(perm) + (size_t)s * n * ib;
                inner.index_base = 1 + (linear ? s * n : 0);
            } else {
                inner.payload = tperm + (size_t)s * n;
                inner.index_base = 1;
            }
        }
        rc = run(kb, inner);
        if (rc != B200_OK) return rc;
    }
    if (pre > 1) {
        if (mode == SORT_KEYS) return transpose_bytes(kb, tkeys, keys, n, pre, post, st);
        const int64_t total = n * nseg;
        if (cdiv(total, 256) > 2147483647LL) return B200_INVALID_VALUE;
        if (ib == 8) perm_scatter_kernel<long long><<<(unsigned int)cdiv(total, 256), 256, 0, st>>>(tperm, static_cast<long long *>(perm), pre, n, post, linear);
        else perm_scatter_kernel<int><<<(unsigned int)cdiv(total, 256), 256, 0, st>>>(tperm, static_cast<int *>(perm), pre, n, post, linear);
        B200_CHECK_LAUNCH();
    }
    return B200_OK;
}

}  // namespace b200

extern "C" {

int32_t b200_sort_dims_workspace_bytes(int32_t dtype_key, int32_t sortperm, int64_t pre, int64_t n, int64_t post,
                                       size_t *bytes) {
    using namespace b200;
    if (!bytes || pre < 0 || n < 0 || post < 0) return B200_INVALID_VALUE;
    *bytes = 0;
    return sort_dims(nullptr, 0, nullptr, nullptr, dtype_key, B200_I64, sortperm ? SORT_PERM : SORT_KEYS, pre, n, post,
                     0, 0, nullptr, bytes);
}

int32_t b200_sort_keys_dims(void *workspace, size_t workspace_bytes, void *keys, int32_t dtype_key, int64_t pre,
                            int64_t n, int64_t post, int32_t descending, b200_stream_t stream) {
    using namespace b200;
    if (pre < 0 || n < 0 || post < 0) return B200_INVALID_VALUE;
    if (pre * n * post > 0 && n > 1 && !keys) return B200_INVALID_VALUE;
    return sort_dims(workspace, workspace_bytes, keys, nullptr, dtype_key, B200_I64, SORT_KEYS, pre, n, post, descending,
                     0, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int32_t b200_sortperm_dims(void *workspace, size_t workspace_bytes, const void *keys, void *perm, int32_t dtype_key,
                           int32_t dtype_index, int64_t pre, int64_t n, int64_t post, int32_t descending,
                           int32_t linear, b200_stream_t stream) {
    using namespace b200;
    if (pre < 0 || n < 0 || post < 0) return B200_INVALID_VALUE;
    if (!index_dtype_ok(dtype_index) || dtype_size(dtype_key) == 0) return B200_UNSUPPORTED;
    if (pre * n * post > 0 && (!keys || !perm)) return B200_INVALID_VALUE;
    return sort_dims(workspace, workspace_bytes, const_cast<void *>(keys), perm, dtype_key, dtype_index, SORT_PERM, pre, n,
                     post, descending, linear, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int32_t b200_sort_workspace_bytes(int32_t dtype_key, int32_t dtype_payload, int64_t n, int64_t nseg, size_t *bytes) {
    using namespace b200;
    if (!bytes || n < 0 || nseg < 0) return B200_INVALID_VALUE;
    *bytes = 0;
    if (dtype_payload < 0) return b200_sort_dims_workspace_bytes(dtype_key, 0, 1, n, nseg, bytes);
    if (dtype_payload == B200_SORTPERM_INDEX) return b200_sort_dims_workspace_bytes(dtype_key, 1, 1, n, nseg, bytes);
    SortXform xf; int kb;
    if (!make_xform(dtype_key, 0, 0, xf, kb)) return B200_UNSUPPORTED;
    if (n == 0) return B200_OK;
    SortCall c{};
    c.n = n; c.xf = xf; c.query = bytes;
    c.mode = SORT_PAIRS;
    c.payload_bytes = dtype_size(dtype_payload);
    if (c.payload_bytes != 4 && c.payload_bytes != 8) return B200_UNSUPPORTED;
    return run(kb, c);
}

int32_t b200_sort_keys_from(void *workspace, size_t workspace_bytes, const void *src, void *dst, int32_t dtype_key,
                            int64_t n, int32_t descending, const void *hist16_dev, b200_stream_t stream) {
    using namespace b200;
    if (n < 0) return B200_INVALID_VALUE;
    SortXform xf; int kb;
    if (!make_xform(dtype_key, descending, 0, xf, kb)) return B200_UNSUPPORTED;
    if (n == 0) return B200_OK;
    if (!src || !dst) return B200_INVALID_VALUE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n == 1 || kb != 4 || (kb & 1) || n <= RB_TILE) {
        // short inputs and the odd-pass (1-byte) layout: copy, then the in-place entry point
        if (src != dst) B200_CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)n * kb, cudaMemcpyDeviceToDevice, st));
        return b200_sort_keys(workspace, workspace_bytes, dst, dtype_key, n, 1, descending, stream);
    }
    SortCall c{};
    c.ws = workspace; c.ws_bytes = workspace_bytes;
    c.keys = dst; c.keys_src = src == dst ? nullptr : src; c.hist16_in = hist16_dev;
    c.mode = SORT_KEYS; c.n = n; c.xf = xf; c.stream = st;
    return run(kb, c);
}

int32_t b200_sort_keys(void *workspace, size_t workspace_bytes, void *keys, int32_t dtype_key, int64_t n, int64_t nseg,
                       int32_t descending, b200_stream_t stream) {
    return b200_sort_keys_dims(workspace, workspace_bytes, keys, dtype_key, 1, n, nseg, descending, stream);
}

int32_t b200_sortperm(void *workspace, size_t workspace_bytes, const void *keys, void *perm, int32_t dtype_key,
                      int32_t dtype_index, int64_t n, int64_t nseg, int32_t descending, int32_t linear,
                      b200_stream_t stream) {
    return b200_sortperm_dims(workspace, workspace_bytes, keys, perm, dtype_key, dtype_index, 1, n, nseg, descending,
                              linear, stream);
}

int32_t b200_sort_pairs(void *workspace, size_t workspace_bytes, void *keys, void *payload, int32_t dtype_key,
                        int32_t dtype_payload, int64_t n, int32_t descending, b200_stream_t stream) {
    using namespace b200;
    if (n < 0) return B200_INVALID_VALUE;
    SortXform xf; int kb;
    // pairs are sortperm's building block: every NaN is ONE key (isless), so NaNs keep their input order
    if (!make_xform(dtype_key, descending, 1, xf, kb)) return B200_UNSUPPORTED;
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

static int32_t select_kth_dispatch(void *ws, size_t wsb, const void *keys, int32_t dtype_key, int64_t n, int64_t k,
                                   int32_t descending, void *out, b200_stream_t stream, size_t *query) {
    using namespace b200;
    SortXform xf; int kb;
    if (!make_xform(dtype_key, descending, 0, xf, kb)) return B200_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (kb) {
        case 1: return select_kth_k1(ws, wsb, keys, n, k, xf, out, st, query);
        case 2: return select_kth_k2(ws, wsb, keys, n, k, xf, out, st, query);
        case 4: return select_kth_k4(ws, wsb, keys, n, k, xf, out, st, query);
        case 8: return select_kth_k8(ws, wsb, keys, n, k, xf, out, st, query);
        default: return B200_UNSUPPORTED;
    }
}

int32_t b200_select_kth_workspace_bytes(int32_t dtype_key, int64_t n, size_t *bytes) {
    if (!bytes || n < 0) return B200_INVALID_VALUE;
    *bytes = 0;
    if (n > 0xFFFFFFFFLL) return B200_UNSUPPORTED;      // 32-bit histogram bins: the query doubles as the capability probe
    return select_kth_dispatch(nullptr, 0, nullptr, dtype_key, n, 0, 0, nullptr, nullptr, bytes);
}

int32_t b200_select_kth(void *workspace, size_t workspace_bytes, const void *keys, int32_t dtype_key, int64_t n,
                        int64_t k, int32_t descending, void *out_value, b200_stream_t stream) {
    if (b200::dtype_size(dtype_key) == 0) return B200_UNSUPPORTED;
    if (n <= 0 || k < 0 || k >= n || !keys || !out_value) return B200_INVALID_VALUE;
    if (n > 0xFFFFFFFFLL) return B200_UNSUPPORTED;      // the per-digit histograms are 32-bit bins (the Julia hook then falls back)
    return select_kth_dispatch(workspace, workspace_bytes, keys, dtype_key, n, k, descending, out_value, stream, nullptr);
}

}  // extern "C"
