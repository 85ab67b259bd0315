// Temporary: entry points declared in include/b200arrays.h that are not implemented yet.
#include "common.cuh"
extern "C" {
int32_t b200_select_workspace_bytes(int64_t, size_t *) { return B200_UNSUPPORTED; }
int32_t b200_select_flagged(void *, size_t, const void *, void *, int32_t, void *, int64_t, b200_stream_t) { return B200_UNSUPPORTED; }
}
