// sort_host.h — host-side call descriptor shared by api_sort.cu and the per-key-width units.
#pragma once
#include "common.cuh"
#include "sort.cuh"
#include "sort_block.cuh"

namespace b200 {

enum SortMode { SORT_KEYS = 0, SORT_PERM = 1, SORT_PAIRS = 2 };

struct SortCall {
    void *ws;
    size_t ws_bytes;
    void *keys;            // KEYS/PAIRS: sorted in place; PERM: read only
    void *payload;         // PERM: output permutation; PAIRS: permuted in place
    int mode;
    int payload_bytes;     // PAIRS: 4 or 8
    int index_dtype;       // PERM: B200_I32/I64/U32/U64
    int64_t index_base;    // PERM: added to the 0-based element index
    int64_t n;
    SortXform xf;
    cudaStream_t stream;
    size_t *query;         // non-null: only compute the workspace size
    const void *keys_src;  // KEYS: where the unsorted keys are read from (nullptr: `keys` itself, in place).  The sorted
                           // result always lands in `keys`; `keys_src` is only read (multi-GPU: the receive buffer).
    const void *hist16_in; // hybrid MSD path: device histogram of the top 16 bits of the order-mapped keys, computed
                           // by the caller (b200_mgpu_plan) — the sort then skips its own histogram pass
};

int32_t sort_run_k1(const SortCall &);
int32_t sort_run_k2(const SortCall &);
int32_t sort_run_k4(const SortCall &);
int32_t sort_run_k8(const SortCall &);
int32_t block_sort_k1(const BlockSortParams &, int, int64_t, cudaStream_t);
int32_t block_sort_k2(const BlockSortParams &, int, int64_t, cudaStream_t);
int32_t block_sort_k4(const BlockSortParams &, int, int64_t, cudaStream_t);
int32_t block_sort_k8(const BlockSortParams &, int, int64_t, cudaStream_t);
int32_t lower_bound_k1(const void *, int64_t, const void *, int64_t, const SortXform &, long long *, cudaStream_t);
int32_t lower_bound_k2(const void *, int64_t, const void *, int64_t, const SortXform &, long long *, cudaStream_t);
int32_t lower_bound_k4(const void *, int64_t, const void *, int64_t, const SortXform &, long long *, cudaStream_t);
int32_t lower_bound_k8(const void *, int64_t, const void *, int64_t, const SortXform &, long long *, cudaStream_t);
int32_t select_kth_k1(void *, size_t, const void *, int64_t, int64_t, const SortXform &, void *, cudaStream_t, size_t *);
int32_t select_kth_k2(void *, size_t, const void *, int64_t, int64_t, const SortXform &, void *, cudaStream_t, size_t *);
int32_t select_kth_k4(void *, size_t, const void *, int64_t, int64_t, const SortXform &, void *, cudaStream_t, size_t *);
int32_t select_kth_k8(void *, size_t, const void *, int64_t, int64_t, const SortXform &, void *, cudaStream_t, size_t *);

}  // namespace b200
