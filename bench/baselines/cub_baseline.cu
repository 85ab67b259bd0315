// cub_baseline.cu — "L" comparator of SURVEY §8(d): CUB's device-wide primitives (the state of practice) called
// directly, on the same inputs as the graft.  Baseline only: built into its own shared library under
// bench/baselines/, loaded by bench.py's comparators, never linked into libb200arrays.so.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#define EXPORT extern "C" __attribute__((visibility("default")))

// Two-phase like CUB itself: temp == nullptr -> *temp_bytes is set.
EXPORT int cub_sort_keys_u32(void *temp, size_t *temp_bytes, const uint32_t *in, uint32_t *out, int64_t n, void *stream) {
    return (int)cub::DeviceRadixSort::SortKeys(temp, *temp_bytes, in, out, n, 0, 32, (cudaStream_t)stream);
}
EXPORT int cub_sort_keys_f32(void *temp, size_t *temp_bytes, const float *in, float *out, int64_t n, void *stream) {
    return (int)cub::DeviceRadixSort::SortKeys(temp, *temp_bytes, in, out, n, 0, 32, (cudaStream_t)stream);
}
EXPORT int cub_sort_pairs_f32_u32(void *temp, size_t *temp_bytes, const float *kin, float *kout, const uint32_t *vin,
                                  uint32_t *vout, int64_t n, void *stream) {
    return (int)cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, kin, kout, vin, vout, n, 0, 32, (cudaStream_t)stream);
}
EXPORT int cub_inclusive_sum_f32(void *temp, size_t *temp_bytes, const float *in, float *out, int64_t n, void *stream) {
    return (int)cub::DeviceScan::InclusiveSum(temp, *temp_bytes, in, out, n, (cudaStream_t)stream);
}
EXPORT int cub_inclusive_sum_i64(void *temp, size_t *temp_bytes, const long long *in, long long *out, int64_t n, void *stream) {
    return (int)cub::DeviceScan::InclusiveSum(temp, *temp_bytes, in, out, n, (cudaStream_t)stream);
}
EXPORT int cub_sum_f32(void *temp, size_t *temp_bytes, const float *in, float *out, int64_t n, void *stream) {
    return (int)cub::DeviceReduce::Sum(temp, *temp_bytes, in, out, n, (cudaStream_t)stream);
}
