// ops.cuh — operator / map functors shared by the reduce and scan kernels.
// Semantics follow SURVEY §8 a1' (Julia Base + GPUArrays, restated):
//   ADD/MUL  wrap for integers, plain IEEE for floats;
//   MAX/MIN  IEEE-754-2019 maximum/minimum for floats: NaN-propagating, -0.0 < +0.0
//            (CUDACore/src/device/intrinsics/math.jl:458-487);
//   AND/OR   bitwise (logical for Bool bytes holding 0/1);
//   ABS2     integers: x*x wrapping in the INPUT type, then converted; floats: x*x
//            evaluated in the accumulator type.
#pragma once
#include "common.cuh"
#include <limits>
#include <math_constants.h>

namespace b200 {

template <class A> struct NumLim {
    __host__ __device__ static constexpr A lowest() { return std::numeric_limits<A>::lowest(); }
    __host__ __device__ static constexpr A highest() { return std::numeric_limits<A>::max(); }
};
template <> struct NumLim<float> {
    __device__ static float lowest() { return -CUDART_INF_F; }
    __device__ static float highest() { return CUDART_INF_F; }
};
template <> struct NumLim<double> {
    __device__ static double lowest() { return -CUDART_INF; }
    __device__ static double highest() { return CUDART_INF; }
};

template <class A> __device__ __forceinline__ A ieee_max(A a, A b) {
    if constexpr (std::is_same<A, float>::value) {
        // fmaxf drops NaNs; Julia's max propagates them.  max.NaN.f32 is exactly
        // IEEE-754-2019 maximum (NaN-propagating, -0.0 < +0.0): one FMNMX.NAN.
        float r;
        asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
    } else if constexpr (std::is_same<A, double>::value) {
        if (a != a) return a;
        if (b != b) return b;
        if (a == b) return __longlong_as_double(__double_as_longlong(a) & __double_as_longlong(b));
        return a > b ? a : b;
    } else {
        return a > b ? a : b;
    }
}
template <class A> __device__ __forceinline__ A ieee_min(A a, A b) {
    if constexpr (std::is_same<A, float>::value) {
        float r;
        asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
        return r;
    } else if constexpr (std::is_same<A, double>::value) {
        if (a != a) return a;
        if (b != b) return b;
        if (a == b) return __longlong_as_double(__double_as_longlong(a) | __double_as_longlong(b));
        return a < b ? a : b;
    } else {
        return a < b ? a : b;
    }
}

// Binary operators on the accumulator type A.
template <class A> struct OpAdd {
    using acc_t = A;
    using val_t = A;
    __device__ static A lift(A v) { return v; }
    __device__ static A identity() { return A(0); }
    __device__ static A apply(A a, A b) { return a + b; }
};
template <class A> struct OpMul {
    using acc_t = A;
    using val_t = A;
    __device__ static A lift(A v) { return v; }
    __device__ static A identity() { return A(1); }
    __device__ static A apply(A a, A b) { return a * b; }
};
template <class A> struct OpMax {
    using acc_t = A;
    using val_t = A;
    __device__ static A lift(A v) { return v; }
    __device__ static A identity() { return NumLim<A>::lowest(); }
    __device__ static A apply(A a, A b) { return ieee_max(a, b); }
};
template <class A> struct OpMin {
    using acc_t = A;
    using val_t = A;
    __device__ static A lift(A v) { return v; }
    __device__ static A identity() { return NumLim<A>::highest(); }
    __device__ static A apply(A a, A b) { return ieee_min(a, b); }
};
template <class A> struct OpAnd {
    using acc_t = A;
    using val_t = A;
    __device__ static A lift(A v) { return v; }
    __device__ static A identity() { return std::is_same<A, uint8_t>::value ? A(1) : A(~A(0)); }
    __device__ static A apply(A a, A b) { return a & b; }
};
template <class A> struct OpOr {
    using acc_t = A;
    using val_t = A;
    __device__ static A lift(A v) { return v; }
    __device__ static A identity() { return A(0); }
    __device__ static A apply(A a, A b) { return a | b; }
};
// Bool MAX/MIN/MUL are OR/AND on 0/1 bytes; unsigned wrap is harmless there.

// (min, max) pair for extrema.
template <class A> struct MinMax { A mn, mx; };
template <class A> struct OpExtrema {
    using acc_t = MinMax<A>;
    using val_t = A;
    __device__ static acc_t lift(A v) { return {v, v}; }
    __device__ static acc_t identity() { return {NumLim<A>::highest(), NumLim<A>::lowest()}; }
    __device__ static acc_t apply(acc_t a, acc_t b) { return {ieee_min(a.mn, b.mn), ieee_max(a.mx, b.mx)}; }
};

// Element maps: input storage type -> accumulator value.
struct MapIdentity {
    template <class A, class Tin> __device__ static A apply(Tin x) { return cvt<A>(x); }
};
struct MapAbs2 {
    template <class A, class Tin> __device__ static A apply(Tin x) {
        if constexpr (std::is_integral<Tin>::value) {
            using U = typename std::make_unsigned<Tin>::type;
            U u = (U)x;
            Tin sq = (Tin)(U)(u * u);  // wraps in the input type, like Julia's abs2(::Int32)
            return cvt<A>(sq);
        } else {
            A a = cvt<A>(x);
            return a * a;
        }
    }
};

}  // namespace b200
