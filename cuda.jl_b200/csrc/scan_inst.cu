// scan_inst.cu — one translation unit per (dtype_in -> dtype_out) pair (see Makefile).
#include "scan.cuh"

namespace b200 {

using Tin = CType<B200_PAIR_IN>::type;
using Tout = CType<B200_PAIR_OUT>::type;
using A = AccOf<Tout>::type;

constexpr bool kInFloat = is_float_like<Tin>::value;
constexpr bool kBool = (B200_PAIR_IN == B200_BOOL);
constexpr bool kBoolOut = (B200_PAIR_OUT == B200_BOOL);

template <class Op>
static int32_t go(void *ws, size_t wsb, const void *in, void *out, int64_t pre, int64_t n, int64_t post, int excl,
                  const void *init, const void *carry, int ncarry, cudaStream_t st, size_t *q) {
    return launch_scan<Tin, Tout, Op>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
}

int32_t B200_PAIR_NAME(void *ws, size_t wsb, const void *in, void *out, int op, int64_t pre, int64_t n, int64_t post,
                       int excl, const void *init, const void *carry, int ncarry, cudaStream_t st, size_t *q) {
    if constexpr (kBool && kBoolOut) {
        switch (op) {
            case B200_OP_AND: case B200_OP_MIN: case B200_OP_MUL: return go<OpAnd<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
            case B200_OP_OR: case B200_OP_MAX: return go<OpOr<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
            default: return B200_UNSUPPORTED;
        }
    } else if constexpr (kBool) {
        if (op == B200_OP_ADD) return go<OpAdd<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
        return B200_UNSUPPORTED;
    } else {
        switch (op) {
            case B200_OP_ADD: return go<OpAdd<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
            case B200_OP_MUL: return go<OpMul<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
            case B200_OP_MAX: return go<OpMax<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
            case B200_OP_MIN: return go<OpMin<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
            case B200_OP_AND:
                if constexpr (!kInFloat) return go<OpAnd<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
                else return B200_UNSUPPORTED;
            case B200_OP_OR:
                if constexpr (!kInFloat) return go<OpOr<A>>(ws, wsb, in, out, pre, n, post, excl, init, carry, ncarry, st, q);
                else return B200_UNSUPPORTED;
            default: return B200_UNSUPPORTED;
        }
    }
}

}  // namespace b200
