/*
 * b200arrays.h — C ABI of libb200arrays.so: B200 (sm_100a) reduce / scan / sort
 * primitives for CUDA.jl's CuArray hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces ONE reference
 * interface (paths relative to the JuliaGPU/CUDA.jl v6.3.0 tree):
 *
 *   b200_reduce            GPUArrays.mapreducedim!(f, op, R::AnyCuArray, A; init)
 *                          CUDACore/src/mapreduce.jl:168-301 (kernels :90-157)
 *   b200_scan              CUDACore.scan!(f, output, input; dims, init, neutral)
 *                          CUDACore/src/accumulate.jl:135-199 (kernels :15-131)
 *   b200_sort_keys         Base.sort!(c::AnyCuArray; rev, dims)  -> bitonic_sort!(c)
 *                          CUDACore/src/sorting.jl:1003-1009, :909-974
 *   b200_sortperm          Base.sortperm!(ix, A; rev) / Base.sortperm(c)
 *                          CUDACore/src/sorting.jl:1051-1070 (bitonic_sort!((A, ix)))
 *   b200_sort_pairs        (key, payload) variant used by the multi-GPU sample sort
 *   b200_select_flagged    findall(::CuArray{Bool}) = cumsum + scatter
 *                          CUDACore/src/indexing.jl:26-66   ("next")
 *
 * Conventions (mirroring how lib/curand binds libcurand,
 * lib/curand/src/libcurand.jl:166-215):
 *   - plain C symbols, no C++ types, no exceptions; int32 status return, 0 = OK.
 *   - device pointers are raw `void*` (CuPtr{T} is ABI-identical to Ptr,
 *     CUDACore/src/pointer.jl:13-28); the caller's stream (CUstream ==
 *     cudaStream_t, CUDACore/lib/cudadrv/stream.jl:78) is explicit in every call.
 *   - the library allocates NOTHING on the device and keeps no device state:
 *     temporaries come from the caller (CUDA.jl's stream-ordered pool through
 *     with_workspace, CUDACore/src/utils/call.jl:23-103).  Two-phase protocol:
 *     `*_workspace_bytes` first, then the call proper with a buffer of at least
 *     that many bytes (256-byte aligned).  The workspace may be uninitialised;
 *     all control words are (re)initialised on `stream` inside the call.
 *   - no cudaSetDevice / cudaDeviceSynchronize / cudaStreamSynchronize inside:
 *     calls are asynchronous on `stream`, re-entrant, thread-safe and
 *     CUDA-graph capturable.
 *   - arrays are dense and column-major.  An N-d array reduced / scanned /
 *     sorted along one dimension (or one contiguous group of dimensions) is
 *     passed as its (pre, n, post) factorisation:
 *         element (i, r, j) lives at  i + pre * (r + n * j),
 *     i in [0,pre), r in [0,n) along the processed dimension, j in [0,post).
 */
#ifndef B200ARRAYS_H
#define B200ARRAYS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

/* Opaque here so that the header needs no CUDA headers; this is CUstream. */
typedef struct CUstream_st *b200_stream_t;

typedef enum {
    B200_OK = 0,
    B200_INVALID_VALUE = 1,        /* bad pointer / negative size / bad enum        */
    B200_UNSUPPORTED = 2,          /* (dtype, op) combination outside the graft     */
    B200_WORKSPACE_TOO_SMALL = 3,
    B200_LAUNCH_FAILURE = 4,       /* cudaError of the failing call: b200_last_cuda_error() */
    B200_WRONG_DEVICE = 5          /* not an sm_100 device                           */
} b200_status_t;

/* Element types (Julia names in comments). Bool is one byte holding 0 or 1. */
typedef enum {
    B200_F16 = 0,   /* Float16  */
    B200_BF16 = 1,  /* BFloat16 (BFloat16s.jl) */
    B200_F32 = 2,   /* Float32  */
    B200_F64 = 3,   /* Float64  */
    B200_I32 = 4,   /* Int32    */
    B200_I64 = 5,   /* Int64    */
    B200_U32 = 6,   /* UInt32   */
    B200_U64 = 7,   /* UInt64   */
    B200_BOOL = 8,  /* Bool     */
    B200_I8 = 9,    /* Int8   (sort only) */
    B200_U8 = 10,   /* UInt8  (sort only) */
    B200_I16 = 11,  /* Int16  (sort only) */
    B200_U16 = 12,  /* UInt16 (sort only) */
    B200_DTYPE_COUNT = 13
} b200_dtype_t;

/* Binary reduction / scan operators.  ADD covers both `+` and Base.add_sum, MUL
 * both `*` and Base.mul_prod: the widening those imply (Int32 -> Int64, Bool ->
 * Int64 ...) is expressed by the caller through dtype_out, exactly as GPUArrays
 * expresses it through eltype(R).  MAX / MIN follow IEEE-754-2019 maximum /
 * minimum for floats (NaN-propagating, -0.0 < +0.0), as the reference's device
 * overrides do (CUDACore/src/device/intrinsics/math.jl:458-487).
 * EXTREMA is Base.ExtremaMap / Base._extrema_rf: the output holds (min, max)
 * pairs of dtype_out, i.e. 2 elements per reduced slice. */
typedef enum {
    B200_OP_ADD = 0,
    B200_OP_MUL = 1,
    B200_OP_MAX = 2,
    B200_OP_MIN = 3,
    B200_OP_AND = 4,
    B200_OP_OR = 5,
    B200_OP_EXTREMA = 6,   /* reduce only */
    B200_OP_COUNT = 7
} b200_op_t;

/* Element map applied before the operator (the `f` of mapreduce). */
typedef enum {
    B200_MAP_IDENTITY = 0,
    B200_MAP_ABS2 = 1,     /* integers: x*x wrapping in the input type; floats: x*x */
    B200_MAP_COUNT = 2
} b200_map_t;

/* How the prior contents of the output are treated (the `init` kwarg of
 * mapreducedim!, mapreduce.jl:108-112,142-146). */
typedef enum {
    B200_INIT_NEUTRAL = 0,   /* init === neutral_element(op, T): out is overwritten      */
    B200_INIT_FROM_OUT = 1   /* init === nothing: out = op(out_before, reduction), folded once */
} b200_init_t;

/* ---- library ----------------------------------------------------------- */
int32_t b200_version(void);                       /* major*10000 + minor*100 + patch */
const char *b200_status_string(int32_t status);
int32_t b200_last_cuda_error(void);               /* thread-local cudaError_t of the last LAUNCH_FAILURE */
/* 0 when the current device is compute capability 10.x, B200_WRONG_DEVICE otherwise. */
int32_t b200_check_device(void);
/* Number of kernels launched by this thread through the library since the last
 * reset (bench.py's gpu_launches).  Host-side counter only. */
int64_t b200_launch_count(void);
void b200_reset_launch_count(void);

/* ---- reduce: GPUArrays.mapreducedim! (mapreduce.jl:168) ----------------- */
/* out[i, 0, j] = op over r of map(in[i, r, j]);  in: pre*red*post elements of
 * dtype_in, out: pre*post elements of dtype_out (2x for EXTREMA).
 * Accumulation happens in dtype_out (Float16/BFloat16: in Float32, rounded once).
 * When pre*post >= 1024 * #SM and pre > 1 (the reference's serial-kernel
 * regime, mapreduce.jl:162-166,203) each output is accumulated sequentially in
 * index order, which the reference's own tests pin (test/core/array.jl:771-773).
 * red == 0 or pre*post == 0: returns OK without touching out (mapreduce.jl:175). */
int32_t b200_reduce_workspace_bytes(int32_t dtype_in, int32_t dtype_out, int32_t op, int32_t map,
                                    int64_t pre, int64_t red, int64_t post, size_t *bytes);
int32_t b200_reduce(void *workspace, size_t workspace_bytes,
                    const void *in, void *out,
                    int32_t dtype_in, int32_t dtype_out, int32_t op, int32_t map,
                    int64_t pre, int64_t red, int64_t post,
                    int32_t init_mode, b200_stream_t stream);

/* findmax / findmin / argmax / argmin (reference: GPUArrays findminmax through
 * mapreducedim!, tests test/core/array.jl:767-769,782-784).  out_vals[i,0,j] = extremal value of the
 * slice (dtype as the input), out_idx = its 1-based LINEAR index into the input (Int64).  Julia's
 * rules: first extremal element wins; NaN is the maximum for findmax and the minimum for findmin;
 * -0.0 < +0.0. */
int32_t b200_findminmax_workspace_bytes(int32_t dtype, int64_t pre, int64_t red, int64_t post, size_t *bytes);
int32_t b200_findminmax(void *workspace, size_t workspace_bytes, const void *in, void *out_vals,
                        void *out_idx, int32_t dtype, int32_t is_max,
                        int64_t pre, int64_t red, int64_t post, b200_stream_t stream);

/* ---- scan: CUDACore.scan! (accumulate.jl:135) --------------------------- */
/* Inclusive (exclusive != 0: exclusive) prefix-op along n.  `init` is a HOST
 * pointer to one dtype_out value (the x of Some(x)) folded into every output,
 * or NULL (accumulate.jl:90-92,123-125).  Exclusive scans start from init (or
 * the operator's neutral element).  in == out (exact aliasing) is allowed.
 * Ops: ADD, MUL, MAX, MIN, AND, OR. */
int32_t b200_scan_workspace_bytes(int32_t dtype_in, int32_t dtype_out, int32_t op,
                                  int64_t pre, int64_t n, int64_t post, size_t *bytes);
int32_t b200_scan(void *workspace, size_t workspace_bytes,
                  const void *in, void *out,
                  int32_t dtype_in, int32_t dtype_out, int32_t op, int32_t exclusive,
                  int64_t pre, int64_t n, int64_t post,
                  const void *init_host, b200_stream_t stream);
/* Same, plus a DEVICE-side carry: `carry_dev` points at `carry_count` values of the ACCUMULATOR type of dtype_out
 * (Float32 for Float16 / BFloat16, dtype_out otherwise) that are folded, in order, into the seed of every chain
 * after `init_host`.  This is the sharded scan of SURVEY section 8(e): rank r passes the all-gathered shard totals and
 * carry_count = r, so the carry never visits the host (no sync, graph-capturable).  No reference counterpart
 * (the reference has no multi-GPU scan); single-GPU semantics are those of b200_scan with
 * init = op(init_host, carry[0], ..., carry[carry_count-1]). */
int32_t b200_scan_carry(void *workspace, size_t workspace_bytes,
                        const void *in, void *out,
                        int32_t dtype_in, int32_t dtype_out, int32_t op, int32_t exclusive,
                        int64_t pre, int64_t n, int64_t post,
                        const void *init_host, const void *carry_dev, int32_t carry_count,
                        b200_stream_t stream);

/* ---- multi-GPU keys-only sort (no reference counterpart) -------------------------------------
 * One process (Julia task) per GPU strings four calls together around ONE all-gather of 256 KB per rank; splitters,
 * counts and destination offsets never visit the host (csrc/mgpu_sort.cu):
 *   b200_mgpu_hist16   -> hist16_dev[65536] (uint32): histogram of the top 16 bits of the order-mapped keys
 *   (all-gather the G histograms into hist_all_dev[G][65536])
 *   b200_mgpu_plan     -> *plan_dev (b200_mgpu_plan_t) and hist_recv_dev[65536]; identical decisions on every rank
 *   b200_mgpu_scatter  -> one pass over the shard; bucket b is stored straight into dst_keys[b] (rank b's receive
 *                         buffer, peer-mapped: NVLink stores) starting at plan.dst_base[b]; every receive buffer must
 *                         hold plan.max_recv_total keys; the caller synchronises the ranks before and after
 *   b200_sort_keys_from(src = receive buffer, dst, hist16_dev = hist_recv_dev) -> the rank's sorted slice
 * Key types: UInt32, Int32, Float32; n <= 2^30 per rank; at most 16 ranks. */
typedef struct b200_mgpu_plan {
    uint32_t spl[17];             /* bucket b = 16-bit bins [spl[b], spl[b+1]) */
    uint32_t ticket;
    uint32_t max_shard_len;       /* keys held by the longest shard (saturating) */
    uint32_t pad;
    uint64_t dst_base[16];        /* element index of this rank's first key inside destination b's receive buffer */
    uint64_t send_count[16];      /* keys this rank sends to destination b */
    uint64_t recv_count[16];      /* keys this rank receives from source s */
    uint64_t recv_total;          /* keys this rank receives in total (its slice of the result) */
    uint64_t max_recv_total;      /* maximum of recv_total over all ranks */
    uint64_t total;               /* N */
} b200_mgpu_plan_t;
int32_t b200_mgpu_plan_bytes(size_t *bytes);
int32_t b200_mgpu_hist16(const void *keys, int32_t dtype_key, int64_t n, int32_t descending, int32_t nan_canonical,
                         void *hist16_dev, b200_stream_t stream);   /* nan_canonical: 0 for sort!, 1 for sortperm / pairs */
int32_t b200_mgpu_plan(const void *hist_all_dev, int32_t nranks, int32_t rank, void *plan_dev, void *hist_recv_dev,
                       b200_stream_t stream);
int32_t b200_mgpu_scatter_workspace_bytes(int64_t n, size_t *bytes);
int32_t b200_mgpu_scatter(void *workspace, size_t workspace_bytes, const void *keys, int32_t dtype_key, int64_t n,
                          int32_t descending, void *plan_dev, void *const *dst_keys, int32_t nranks,
                          b200_stream_t stream);
/* sortperm's exchange: STABLE partition of (keys, 8-byte payload) by the plan's splitters, bucket b stored straight into
 * dst_keys[b] / dst_payload[b] at plan.dst_base[b].  Workspace: b200_partition_workspace_bytes(n, nranks) + 256. */
int32_t b200_mgpu_partition_pairs(void *workspace, size_t workspace_bytes, const void *keys, const void *payload,
                                  int32_t dtype_key, int64_t n, int32_t descending, const void *plan_dev,
                                  void *const *dst_keys, void *const *dst_payload, int32_t nranks, b200_stream_t stream);
/* The same exchange with a GENERATED 4-byte payload instead of one read from memory: (rank << shift) | source index
 * (needs n <= 2^shift and nranks <= 2^(32 - shift)); dst_index[b]: UInt32 receive buffers.  After the local pairs sort,
 * b200_mgpu_widen_index turns the tagged indices into global 1-based Int64 indices:
 * out[i] = offsets[index[i] >> shift] + (index[i] & (2^shift - 1)) + 1, offsets = Int64[16] on the device (the number of
 * keys held by the ranks before each source rank). */
int32_t b200_mgpu_partition_index(void *workspace, size_t workspace_bytes, const void *keys, int32_t dtype_key, int64_t n,
                                  int32_t descending, const void *plan_dev, int32_t rank, int32_t shift,
                                  void *const *dst_keys, void *const *dst_index, int32_t nranks, b200_stream_t stream);
int32_t b200_mgpu_widen_index(const void *index_u32, int64_t n, int32_t shift, const void *offsets_dev, void *out_i64,
                              b200_stream_t stream);
/* sort! of `n` keys read from `src` into `dst` (src is only read; src == dst sorts in place; workspace as for
 * b200_sort_workspace_bytes(dtype_key, -1, n, 1)).  hist16_dev, if not NULL, is the exact histogram of the top 16 bits
 * of the order-mapped keys of `src` (b200_mgpu_plan's hist_recv_dev): the sort then skips its own histogram pass. */
int32_t b200_sort_keys_from(void *workspace, size_t workspace_bytes, const void *src, void *dst, int32_t dtype_key,
                            int64_t n, int32_t descending, const void *hist16_dev, b200_stream_t stream);

/* ---- sort: Base.sort! / sortperm! hooks (sorting.jl:1003-1070) ---------- */
/* Order is Julia's `isless`: -Inf < ... < -0.0 < +0.0 < ... < +Inf < NaN (all
 * NaNs equal, last), two's-complement order for signed ints; `descending` is
 * `rev=true` (stable: ties keep ascending index, sorting.jl:582-592).
 * Stable LSD radix sort.  Segments: `nseg` independent contiguous vectors of
 * `n` keys each (sort!(A; dims=1) of an n x nseg matrix); nseg = 1 for vectors. */
/* dtype_payload: -1 = keys only (b200_sort_keys), B200_SORTPERM_INDEX = b200_sortperm,
 * otherwise the payload dtype of b200_sort_pairs. */
#define B200_SORTPERM_INDEX 100
int32_t b200_sort_workspace_bytes(int32_t dtype_key, int32_t dtype_payload,
                                  int64_t n, int64_t nseg, size_t *bytes);
/* In place. */
int32_t b200_sort_keys(void *workspace, size_t workspace_bytes,
                       void *keys, int32_t dtype_key, int64_t n, int64_t nseg,
                       int32_t descending, b200_stream_t stream);
/* perm[k] = 1-based index (within its segment when `linear` == 0, else the
 * 1-based linear index into the whole array, sorting.jl:1067-1070) of the k-th
 * smallest key; keys are not modified (sorting.jl:857-858).  dtype_index is any
 * of I32/I64/U32/U64 (test/core/sorting.jl:395-403). */
int32_t b200_sortperm(void *workspace, size_t workspace_bytes,
                      const void *keys, void *perm,
                      int32_t dtype_key, int32_t dtype_index, int64_t n, int64_t nseg,
                      int32_t descending, int32_t linear, b200_stream_t stream);
/* General (pre, n, post) forms: sort!(A; dims) / sortperm(A; dims) along ANY dimension
 * (sorting.jl:909-974 with dims; tests test/core/sorting.jl:314-315,388-392).  Segments of up
 * to 4096 keys are sorted by one CTA each, whatever their stride; longer strided segments are
 * transposed, sorted and transposed back.  b200_sort_keys / b200_sortperm are the pre == 1 case. */
int32_t b200_sort_dims_workspace_bytes(int32_t dtype_key, int32_t sortperm, int64_t pre, int64_t n,
                                       int64_t post, size_t *bytes);
int32_t b200_sort_keys_dims(void *workspace, size_t workspace_bytes, void *keys, int32_t dtype_key,
                            int64_t pre, int64_t n, int64_t post, int32_t descending,
                            b200_stream_t stream);
int32_t b200_sortperm_dims(void *workspace, size_t workspace_bytes, const void *keys, void *perm,
                           int32_t dtype_key, int32_t dtype_index, int64_t pre, int64_t n,
                           int64_t post, int32_t descending, int32_t linear, b200_stream_t stream);
/* Sorts keys in place and permutes `payload` (4- or 8-byte elements, dtype
 * I32/U32/I64/U64/F32/F64 treated as raw bits) along with them. */
int32_t b200_sort_pairs(void *workspace, size_t workspace_bytes,
                        void *keys, void *payload,
                        int32_t dtype_key, int32_t dtype_payload, int64_t n,
                        int32_t descending, b200_stream_t stream);

/* lower_bounds[j] = number of keys in `sorted_keys` (sorted by b200_sort_keys with the same
 * `descending`) that order strictly before needles[j]; int64 on the device.  The multi-GPU
 * sample sort cuts each locally sorted shard at the splitters with it (no
 * reference counterpart: the reference has no multi-GPU sort). */
int32_t b200_search_sorted(const void *sorted_keys, int64_t n, const void *needles, int64_t m,
                           int32_t dtype_key, int32_t descending, void *lower_bounds_dev,
                           b200_stream_t stream);

/* Stable partition of keys (and an optional 8-byte payload) into `nbuckets` <= 16 buckets by
 * nbuckets-1 splitter keys (device): bucket b = keys with splitter[b-1] <= key < splitter[b] in
 * the sort's order (isless / descending).  bucket_counts_dev: int64[nbuckets].  First stage of the
 * multi-GPU sample sort: partition the unsorted shard, all-to-all, sort once. */
int32_t b200_partition_workspace_bytes(int64_t n, int32_t nbuckets, size_t *bytes);
int32_t b200_partition(void *workspace, size_t workspace_bytes, const void *keys_in, void *keys_out,
                       const void *payload_in, void *payload_out, int32_t dtype_key, int64_t n,
                       const void *splitters, int32_t nbuckets, int32_t descending,
                       void *bucket_counts_dev, b200_stream_t stream);

/* Two-phase form for the fused partition + exchange of the multi-GPU sort: `count` leaves the
 * per-tile prefixes in the workspace and the bucket totals in bucket_counts_dev; after the ranks
 * have exchanged their totals, `scatter` (same workspace, same inputs) writes bucket b straight to
 * dst_keys[b] (+ dst_payload[b]) starting at element dst_base[b] — the destinations are the
 * ranks' receive buffers mapped over NVLink (symmetric memory), so the all-to-all is the
 * kernel's own coalesced peer stores.  dst_* are HOST arrays of nbuckets entries. */
int32_t b200_partition_count(void *workspace, size_t workspace_bytes, const void *keys_in, int32_t dtype_key,
                             int64_t n, const void *splitters, int32_t nbuckets, int32_t descending,
                             void *bucket_counts_dev, b200_stream_t stream);
int32_t b200_partition_scatter(void *workspace, size_t workspace_bytes, const void *keys_in,
                               const void *payload_in, int32_t dtype_key, int64_t n, const void *splitters,
                               int32_t nbuckets, int32_t descending, void *bucket_counts_dev,
                               void *const *dst_keys, void *const *dst_payload, const int64_t *dst_base,
                               b200_stream_t stream);

/* ---- radix select: partialsort(c, k) for a scalar k (sorting.jl:1015-1020,1046-1048) - */
/* The reference's bitonic partialsort is "full sort, then c[k]".  b200_select_kth writes the
 * element that a stable sort (isless order, `descending` as in b200_sort_keys) would put at
 * 0-based position k to out_value[0] (device, element type of the keys) WITHOUT sorting or
 * modifying `keys`.  n >= 2^22: a sorted sample brackets rank k, ONE streaming read compacts the bracket (n / 8 keys of
 * workspace) and the MSD byte passes run over the candidates; a bracket that misses falls back, on the device, to
 * sizeof(key) streaming reads of the input (the path taken for smaller n).  n >= 2^32: B200_UNSUPPORTED (32-bit bins).
 * Keys-only sort! on floats (b200_sort_keys*) keeps NaN payloads but drops the NaN SIGN bit; select returns the same key. */
int32_t b200_select_kth_workspace_bytes(int32_t dtype_key, int64_t n, size_t *bytes);
int32_t b200_select_kth(void *workspace, size_t workspace_bytes, const void *keys, int32_t dtype_key,
                        int64_t n, int64_t k, int32_t descending, void *out_value, b200_stream_t stream);

/* ---- select: findall(::CuArray{Bool}) (indexing.jl:26-66) --------------- */
/* Writes the 1-based linear indices (dtype_index I32/I64) of the non-zero flags
 * to `indices` in order and the number found to *count_dev (device int64).
 * One pass: scan and scatter are fused.  `indices` must hold n entries. */
int32_t b200_select_workspace_bytes(int64_t n, size_t *bytes);
int32_t b200_select_flagged(void *workspace, size_t workspace_bytes,
                            const void *flags_bool, void *indices, int32_t dtype_index,
                            void *count_dev, int64_t n, b200_stream_t stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* B200ARRAYS_H */
