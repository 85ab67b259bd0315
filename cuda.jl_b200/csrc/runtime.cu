// runtime.cu — library-level entry points and host helpers (no kernels).
#include "common.cuh"
#include <mutex>

namespace b200 {

static thread_local int32_t tl_last_cuda_error = 0;
static thread_local int64_t tl_launches = 0;

int32_t fail_cuda(cudaError_t e) {
    tl_last_cuda_error = (int32_t)e;
    return B200_LAUNCH_FAILURE;
}
void note_launch(int n) { tl_launches += n; }

const DevInfo &devinfo() {
    // Per-device attribute cache.  Only queries, never cudaSetDevice.
    static DevInfo cache[64];
    static bool have[64];
    static std::mutex mu;
    static DevInfo fallback{148, 10};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        (void)cudaGetLastError();
        return fallback;
    }
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        DevInfo d{148, 10};
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) d.sm_count = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess) d.cc_major = v;
        (void)cudaGetLastError();
        cache[dev] = d;
        have[dev] = true;
    }
    return cache[dev];
}

}  // namespace b200

extern "C" {

int32_t b200_version(void) { return 100; /* 0.1.0 */ }

const char *b200_status_string(int32_t status) {
    switch (status) {
        case B200_OK: return "B200_OK";
        case B200_INVALID_VALUE: return "B200_INVALID_VALUE: bad pointer, size or enum";
        case B200_UNSUPPORTED: return "B200_UNSUPPORTED: (dtype, op) combination outside the graft";
        case B200_WORKSPACE_TOO_SMALL: return "B200_WORKSPACE_TOO_SMALL";
        case B200_LAUNCH_FAILURE: return "B200_LAUNCH_FAILURE: CUDA error, see b200_last_cuda_error()";
        case B200_WRONG_DEVICE: return "B200_WRONG_DEVICE: libb200arrays is built for sm_100a only";
        default: return "B200: unknown status";
    }
}

int32_t b200_last_cuda_error(void) { return b200::tl_last_cuda_error; }

int32_t b200_check_device(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return B200_WRONG_DEVICE; }
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return B200_WRONG_DEVICE;
    }
    return major == 10 ? B200_OK : B200_WRONG_DEVICE;
}

int64_t b200_launch_count(void) { return b200::tl_launches; }
void b200_reset_launch_count(void) { b200::tl_launches = 0; }

}  // extern "C"
