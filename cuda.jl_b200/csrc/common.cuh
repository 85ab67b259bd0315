// common.cuh — shared device/host helpers for libb200arrays (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>
#include <type_traits>
#ifdef B200_TUNING
#include <cstdlib>
#endif
#include "../../include/b200arrays.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200arrays is written for sm_100a (B200) only"
#endif

namespace b200 {

// ---------------------------------------------------------------- host side
struct DevInfo {
    int sm_count;
    int cc_major;
};
const DevInfo &devinfo();            // attributes of the CURRENT device (cached per device)
void note_launch(int n = 1);         // host-side launch counter (b200_launch_count)
int32_t fail_cuda(cudaError_t e);    // records e, returns B200_LAUNCH_FAILURE

#define B200_CUDA_TRY(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ::b200::fail_cuda(_e);  \
    } while (0)

#define B200_CHECK_LAUNCH()                                   \
    do {                                                      \
        ::b200::note_launch();                                \
        cudaError_t _e = cudaGetLastError();                  \
        if (_e != cudaSuccess) return ::b200::fail_cuda(_e);  \
    } while (0)

// A/B measurement knobs: read from the environment only in -DB200_TUNING builds (tools/build_variant.sh);
// the product library has no environment dependence and always takes the default.
#ifdef B200_TUNING
static inline int tune_int(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }
#else
static inline int tune_int(const char *, int dflt) { return dflt; }
#endif

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Carves 256-byte aligned sub-buffers out of the caller's workspace.
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *p) : base(static_cast<char *>(p)) {}
    template <class T> T *take(size_t count) {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += align_up(count * sizeof(T), 256);
        return p;
    }
    size_t bytes() const { return off; }
};

// ---------------------------------------------------------------- dtypes
template <int D> struct CType;
template <> struct CType<B200_F16> { using type = __half; };
template <> struct CType<B200_BF16> { using type = __nv_bfloat16; };
template <> struct CType<B200_F32> { using type = float; };
template <> struct CType<B200_F64> { using type = double; };
template <> struct CType<B200_I32> { using type = int32_t; };
template <> struct CType<B200_I64> { using type = int64_t; };
template <> struct CType<B200_U32> { using type = uint32_t; };
template <> struct CType<B200_U64> { using type = uint64_t; };
template <> struct CType<B200_BOOL> { using type = uint8_t; };
template <> struct CType<B200_I8> { using type = int8_t; };
template <> struct CType<B200_U8> { using type = uint8_t; };
template <> struct CType<B200_I16> { using type = int16_t; };
template <> struct CType<B200_U16> { using type = uint16_t; };

static inline int dtype_size(int d) {
    switch (d) {
        case B200_F16: case B200_BF16: case B200_I16: case B200_U16: return 2;
        case B200_F32: case B200_I32: case B200_U32: return 4;
        case B200_F64: case B200_I64: case B200_U64: return 8;
        case B200_BOOL: case B200_I8: case B200_U8: return 1;
        default: return 0;
    }
}

// Accumulator type for an output element type: 16-bit floats accumulate in
// Float32 and are rounded once on store.
template <class T> struct AccOf { using type = T; };
template <> struct AccOf<__half> { using type = float; };
template <> struct AccOf<__nv_bfloat16> { using type = float; };

template <class T> struct is_float_like : std::false_type {};
template <> struct is_float_like<float> : std::true_type {};
template <> struct is_float_like<double> : std::true_type {};
template <> struct is_float_like<__half> : std::true_type {};
template <> struct is_float_like<__nv_bfloat16> : std::true_type {};

// Conversions between storage types and accumulator types.
template <class To, class From> __host__ __device__ __forceinline__ To cvt(From x) { return static_cast<To>(x); }
template <> __host__ __device__ __forceinline__ float cvt<float, __half>(__half x) { return __half2float(x); }
template <> __host__ __device__ __forceinline__ float cvt<float, __nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <> __host__ __device__ __forceinline__ __half cvt<__half, float>(float x) { return __float2half_rn(x); }
template <> __host__ __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16, float>(float x) { return __float2bfloat16_rn(x); }
template <> __host__ __device__ __forceinline__ __half cvt<__half, __half>(__half x) { return x; }
template <> __host__ __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16, __nv_bfloat16>(__nv_bfloat16 x) { return x; }
template <> __host__ __device__ __forceinline__ double cvt<double, __half>(__half x) { return (double)__half2float(x); }
template <> __host__ __device__ __forceinline__ double cvt<double, __nv_bfloat16>(__nv_bfloat16 x) { return (double)__bfloat162float(x); }

// ---------------------------------------------------------------- device side
#ifdef __CUDACC__

// Streaming (read-once) global loads: non-coherent path, no L1 allocation.
// 16-byte and 32-byte forms; the 32-byte form is sm_100's LDG.E.256.
struct alignas(16) Vec16 { uint32_t w[4]; };
struct alignas(32) Vec32 { uint32_t w[8]; };

__device__ __forceinline__ Vec16 ldg_stream16(const void *p) {
    Vec16 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p));
    return v;
}
__device__ __forceinline__ Vec32 ldg_stream32(const void *p) {
    Vec32 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]),
                   "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7]) : "l"(p));
    return v;
}
template <class T> __device__ __forceinline__ T ldg_stream(const T *p) {
    if constexpr (sizeof(T) == 1) {
        uint16_t r; asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=h"(r) : "l"(p));
        uint8_t b = (uint8_t)r; T t; memcpy(&t, &b, 1); return t;
    } else if constexpr (sizeof(T) == 2) {
        uint16_t r; asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
        T t; memcpy(&t, &r, 2); return t;
    } else if constexpr (sizeof(T) == 4) {
        uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
        T t; memcpy(&t, &r, 4); return t;
    } else {
        static_assert(sizeof(T) == 8, "unsupported element size");
        uint64_t r; asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
        T t; memcpy(&t, &r, 8); return t;
    }
}

// Coherent streaming loads (no .nc): used where input and output may alias exactly
// (accumulate!(op, x, x)), each address being read before the same CTA overwrites it.
__device__ __forceinline__ Vec16 ld_coh16(const void *p) {
    Vec16 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p) : "memory");
    return v;
}
template <class T> __device__ __forceinline__ T ld_coh(const T *p) {
    if constexpr (sizeof(T) == 1) {
        uint16_t r; asm volatile("ld.global.L1::no_allocate.u8 %0, [%1];" : "=h"(r) : "l"(p) : "memory");
        uint8_t b = (uint8_t)r; T t; memcpy(&t, &b, 1); return t;
    } else if constexpr (sizeof(T) == 2) {
        uint16_t r; asm volatile("ld.global.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p) : "memory");
        T t; memcpy(&t, &r, 2); return t;
    } else if constexpr (sizeof(T) == 4) {
        uint32_t r; asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
        T t; memcpy(&t, &r, 4); return t;
    } else {
        static_assert(sizeof(T) == 8, "unsupported element size");
        uint64_t r; asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
        T t; memcpy(&t, &r, 8); return t;
    }
}

__device__ __forceinline__ void stg_stream16(void *p, const Vec16 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]) : "memory");
}

// Generic shuffle for trivially-copyable T of 4/8/16 bytes (and small ones).
template <class T> __device__ __forceinline__ T shfl_xor_any(T v, int mask, int width = 32) {
    if constexpr (sizeof(T) < 4) {
        uint32_t w = 0; memcpy(&w, &v, sizeof(T));
        w = __shfl_xor_sync(0xffffffffu, w, mask, width);
        T r; memcpy(&r, &w, sizeof(T)); return r;
    } else {
        static_assert(sizeof(T) % 4 == 0, "shuffle needs 4-byte multiples");
        uint32_t w[sizeof(T) / 4];
        memcpy(w, &v, sizeof(T));
#pragma unroll
        for (int i = 0; i < (int)(sizeof(T) / 4); ++i) w[i] = __shfl_xor_sync(0xffffffffu, w[i], mask, width);
        T r; memcpy(&r, w, sizeof(T)); return r;
    }
}
template <class T> __device__ __forceinline__ T shfl_up_any(T v, int delta) {
    if constexpr (sizeof(T) < 4) {
        uint32_t w = 0; memcpy(&w, &v, sizeof(T));
        w = __shfl_up_sync(0xffffffffu, w, delta);
        T r; memcpy(&r, &w, sizeof(T)); return r;
    } else {
        uint32_t w[sizeof(T) / 4];
        memcpy(w, &v, sizeof(T));
#pragma unroll
        for (int i = 0; i < (int)(sizeof(T) / 4); ++i) w[i] = __shfl_up_sync(0xffffffffu, w[i], delta);
        T r; memcpy(&r, w, sizeof(T)); return r;
    }
}
template <class T> __device__ __forceinline__ T shfl_idx_any(T v, int src) {
    if constexpr (sizeof(T) < 4) {
        uint32_t w = 0; memcpy(&w, &v, sizeof(T));
        w = __shfl_sync(0xffffffffu, w, src);
        T r; memcpy(&r, &w, sizeof(T)); return r;
    } else {
        uint32_t w[sizeof(T) / 4];
        memcpy(w, &v, sizeof(T));
#pragma unroll
        for (int i = 0; i < (int)(sizeof(T) / 4); ++i) w[i] = __shfl_sync(0xffffffffu, w[i], src);
        T r; memcpy(&r, w, sizeof(T)); return r;
    }
}

// ---------------------------------------------------------------- TMA (cp.async.bulk) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

#endif  // __CUDACC__
}  // namespace b200
