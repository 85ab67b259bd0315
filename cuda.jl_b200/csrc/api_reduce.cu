// api_reduce.cu — extern "C" entry points for reduce (include/b200arrays.h).
#include "common.cuh"

namespace b200 {
#define X(I, O)                                                                                                   \
    int32_t reduce_pair_##I##_##O(void *, size_t, const void *, void *, int, int, int64_t, int64_t, int64_t, int, \
                                  cudaStream_t, size_t *);
#include "reduce_pairs.def"
#undef X

using ReduceFn = int32_t (*)(void *, size_t, const void *, void *, int, int, int64_t, int64_t, int64_t, int,
                             cudaStream_t, size_t *);

static ReduceFn find_reduce(int din, int dout) {
#define X(I, O) \
    if (din == B200_##I && dout == B200_##O) return reduce_pair_##I##_##O;
#include "reduce_pairs.def"
#undef X
    return nullptr;
}

static int32_t reduce_common(void *ws, size_t wsb, const void *in, void *out, int din, int dout, int op, int map,
                             int64_t pre, int64_t red, int64_t post, int init_mode, cudaStream_t st, size_t *q) {
    if (pre < 0 || red < 0 || post < 0) return B200_INVALID_VALUE;
    if (op < 0 || op >= B200_OP_COUNT || map < 0 || map >= B200_MAP_COUNT) return B200_INVALID_VALUE;
    if (init_mode != B200_INIT_NEUTRAL && init_mode != B200_INIT_FROM_OUT) return B200_INVALID_VALUE;
    ReduceFn fn = find_reduce(din, dout);
    if (!fn) return B200_UNSUPPORTED;
    if (pre == 0 || red == 0 || post == 0) {  // length(A) == 0 && return R  (mapreduce.jl:175)
        if (q) *q = 0;
        return B200_OK;
    }
    if (!q && (in == nullptr || out == nullptr)) return B200_INVALID_VALUE;
    return fn(ws, wsb, in, out, op, map, pre, red, post, init_mode, st, q);
}
}  // namespace b200

extern "C" {

int32_t b200_reduce_workspace_bytes(int32_t dtype_in, int32_t dtype_out, int32_t op, int32_t map, int64_t pre,
                                    int64_t red, int64_t post, size_t *bytes) {
    if (!bytes) return B200_INVALID_VALUE;
    *bytes = 0;
    return b200::reduce_common(nullptr, 0, nullptr, nullptr, dtype_in, dtype_out, op, map, pre, red, post,
                               B200_INIT_NEUTRAL, nullptr, bytes);
}

int32_t b200_reduce(void *workspace, size_t workspace_bytes, const void *in, void *out, int32_t dtype_in,
                    int32_t dtype_out, int32_t op, int32_t map, int64_t pre, int64_t red, int64_t post,
                    int32_t init_mode, b200_stream_t stream) {
    return b200::reduce_common(workspace, workspace_bytes, in, out, dtype_in, dtype_out, op, map, pre, red, post,
                               init_mode, reinterpret_cast<cudaStream_t>(stream), nullptr);
}

}  // extern "C"
