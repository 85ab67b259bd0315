// api_scan.cu — extern "C" entry points for scan (include/b200arrays.h).
#include "common.cuh"

namespace b200 {
#define X(I, O)                                                                                               \
    int32_t scan_pair_##I##_##O(void *, size_t, const void *, void *, int, int64_t, int64_t, int64_t, int,    \
                                const void *, const void *, int, cudaStream_t, size_t *);
#include "scan_pairs.def"
#undef X

using ScanFn = int32_t (*)(void *, size_t, const void *, void *, int, int64_t, int64_t, int64_t, int, const void *,
                           const void *, int, cudaStream_t, size_t *);

static ScanFn find_scan(int din, int dout) {
#define X(I, O) \
    if (din == B200_##I && dout == B200_##O) return scan_pair_##I##_##O;
#include "scan_pairs.def"
#undef X
    return nullptr;
}

static int32_t scan_common(void *ws, size_t wsb, const void *in, void *out, int din, int dout, int op, int excl,
                           int64_t pre, int64_t n, int64_t post, const void *init, const void *carry, int ncarry,
                           cudaStream_t st, size_t *q) {
    if (pre < 0 || n < 0 || post < 0) return B200_INVALID_VALUE;
    if (op < 0 || op >= B200_OP_EXTREMA) return op == B200_OP_EXTREMA ? B200_UNSUPPORTED : B200_INVALID_VALUE;
    ScanFn fn = find_scan(din, dout);
    if (!fn) return B200_UNSUPPORTED;
    if (pre == 0 || n == 0 || post == 0) {  // isempty(inds_t[dims]) && return output (accumulate.jl:141)
        if (q) *q = 0;
        return B200_OK;
    }
    if (!q && (in == nullptr || out == nullptr)) return B200_INVALID_VALUE;
    if (ncarry < 0 || (ncarry > 0 && carry == nullptr)) return B200_INVALID_VALUE;
    return fn(ws, wsb, in, out, op, pre, n, post, excl ? 1 : 0, init, carry, ncarry, st, q);
}
}  // namespace b200

extern "C" {

int32_t b200_scan_workspace_bytes(int32_t dtype_in, int32_t dtype_out, int32_t op, int64_t pre, int64_t n,
                                  int64_t post, size_t *bytes) {
    if (!bytes) return B200_INVALID_VALUE;
    *bytes = 0;
    return b200::scan_common(nullptr, 0, nullptr, nullptr, dtype_in, dtype_out, op, 0, pre, n, post, nullptr, nullptr, 0,
                             nullptr, bytes);
}

int32_t b200_scan(void *workspace, size_t workspace_bytes, const void *in, void *out, int32_t dtype_in,
                  int32_t dtype_out, int32_t op, int32_t exclusive, int64_t pre, int64_t n, int64_t post,
                  const void *init_host, b200_stream_t stream) {
    return b200::scan_common(workspace, workspace_bytes, in, out, dtype_in, dtype_out, op, exclusive, pre, n, post,
                             init_host, nullptr, 0, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

int32_t b200_scan_carry(void *workspace, size_t workspace_bytes, const void *in, void *out, int32_t dtype_in,
                        int32_t dtype_out, int32_t op, int32_t exclusive, int64_t pre, int64_t n, int64_t post,
                        const void *init_host, const void *carry_dev, int32_t carry_count, b200_stream_t stream) {
    return b200::scan_common(workspace, workspace_bytes, in, out, dtype_in, dtype_out, op, exclusive, pre, n, post,
                             init_host, carry_dev, carry_count, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

}  // extern "C"
