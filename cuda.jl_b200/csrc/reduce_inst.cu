// reduce_inst.cu — one translation unit per (dtype_in -> dtype_out) pair, selected
// with -DB200_PAIR_IN=<dtype> -DB200_PAIR_OUT=<dtype> -DB200_PAIR_NAME=<symbol>.
// Keeps nvcc's per-TU work small so `make -j` builds the operator matrix in parallel.
#include "reduce.cuh"

namespace b200 {

using Tin = CType<B200_PAIR_IN>::type;
using Tout = CType<B200_PAIR_OUT>::type;
using A = AccOf<Tout>::type;

constexpr bool kInFloat = is_float_like<Tin>::value;
constexpr bool kBool = (B200_PAIR_IN == B200_BOOL);
constexpr bool kBoolOut = (B200_PAIR_OUT == B200_BOOL);

template <class Op, class Map>
static int32_t go(void *ws, size_t wsb, const void *in, void *out, int64_t pre, int64_t red, int64_t post,
                  int init_mode, cudaStream_t st, size_t *q) {
    return launch_reduce<Tin, Tout, Op, Map>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
}

int32_t B200_PAIR_NAME(void *ws, size_t wsb, const void *in, void *out, int op, int map, int64_t pre, int64_t red,
                       int64_t post, int init_mode, cudaStream_t st, size_t *q) {
    if (map == B200_MAP_ABS2) {
        if (op != B200_OP_ADD) return B200_UNSUPPORTED;
        if constexpr (kBool) return go<OpAdd<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
        else return go<OpAdd<A>, MapAbs2>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
    }
    if (map != B200_MAP_IDENTITY) return B200_UNSUPPORTED;
    if constexpr (kBool && kBoolOut) {
        // Bool -> Bool: any/all and friends on 0/1 bytes
        switch (op) {
            case B200_OP_AND: case B200_OP_MIN: case B200_OP_MUL:
                return go<OpAnd<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
            case B200_OP_OR: case B200_OP_MAX:
                return go<OpOr<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
            default: return B200_UNSUPPORTED;
        }
    } else if constexpr (kBool) {
        // Bool -> Int64: count / sum
        if (op == B200_OP_ADD) return go<OpAdd<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
        return B200_UNSUPPORTED;
    } else {
        switch (op) {
            case B200_OP_ADD: return go<OpAdd<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
            case B200_OP_MUL: return go<OpMul<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
            case B200_OP_MAX: return go<OpMax<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
            case B200_OP_MIN: return go<OpMin<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
            case B200_OP_EXTREMA:
                return go<OpExtrema<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
            case B200_OP_AND:
                if constexpr (!kInFloat) return go<OpAnd<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
                else return B200_UNSUPPORTED;
            case B200_OP_OR:
                if constexpr (!kInFloat) return go<OpOr<A>, MapIdentity>(ws, wsb, in, out, pre, red, post, init_mode, st, q);
                else return B200_UNSUPPORTED;
            default: return B200_UNSUPPORTED;
        }
    }
}

}  // namespace b200
