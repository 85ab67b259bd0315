// Temporary: entry points declared in include/b200arrays.h that are not implemented yet.
#include "common.cuh"
extern "C" {
int32_t b200_sort_workspace_bytes(int32_t, int32_t, int64_t, int64_t, size_t *) { return B200_UNSUPPORTED; }
int32_t b200_sort_keys(void *, size_t, void *, int32_t, int64_t, int64_t, int32_t, b200_stream_t) { return B200_UNSUPPORTED; }
int32_t b200_sortperm(void *, size_t, const void *, void *, int32_t, int32_t, int64_t, int64_t, int32_t, int32_t, b200_stream_t) { return B200_UNSUPPORTED; }
int32_t b200_sort_pairs(void *, size_t, void *, void *, int32_t, int32_t, int64_t, int32_t, b200_stream_t) { return B200_UNSUPPORTED; }
int32_t b200_select_workspace_bytes(int64_t, size_t *) { return B200_UNSUPPORTED; }
int32_t b200_select_flagged(void *, size_t, const void *, void *, int32_t, void *, int64_t, b200_stream_t) { return B200_UNSUPPORTED; }
}
