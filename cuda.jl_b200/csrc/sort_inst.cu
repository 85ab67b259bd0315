// sort_inst.cu — one translation unit per key width (-DB200_KEY_BYTES=1|2|4|8).
// Host orchestration of the onesweep passes for keys / sortperm / pairs.
#include "sort_pipe.cuh"
#include "sort_ctr.cuh"
#include "select_kth.cuh"
#include "sort_block.cuh"
#include "sort_host.h"
#include <cstdlib>

namespace b200 {

using K = KeyOf<B200_KEY_BYTES>::type;
constexpr int P = B200_KEY_BYTES;                       // 8-bit digits: one pass per key byte
// Keys per thread.  Measured on B200 (2^30 UInt32): keys-only 19.9 ms at 16, 17.4 ms at 20, 17.8 ms at 24;
// pairs 24.9 ms at 16, 25.4 ms at 20, 28.5 ms at 24 (shared-memory staging of key + payload costs occupancy).
#ifndef B200_RX_ITEMS_KEYS
#define B200_RX_ITEMS_KEYS 20
#endif
template <class V> struct ItemsFor {
    static constexpr int value = B200_KEY_BYTES == 8 ? 12 : (PayTraits<V>::HAS ? 16 : B200_RX_ITEMS_KEYS);
};
template <class V> constexpr int tile_for() { return RX_THREADS * ItemsFor<V>::value; }

template <class V> static size_t onesweep_smem() {
    using VT = typename PayTraits<V>::type;
    constexpr int TILE = tile_for<V>();
    size_t s = (size_t)RX_WARPS * RX_RADIX * 4 + RX_RADIX * 4 + RX_RADIX * 8 + ((size_t)TILE * sizeof(K) + 15) / 16 * 16;
    if (PayTraits<V>::HAS) s += (size_t)TILE * sizeof(VT);
    return s;
}

template <class V> static size_t pipe_smem() {
    using VT = typename PayTraits<V>::type;
    constexpr int TILE = tile_for<V>();
    size_t s = (size_t)(RXP_STAGES + 2) * TILE * sizeof(K) + (size_t)RX_WARPS * RX_RADIX * 4 + 2 * RX_RADIX * 8;
    if (PayTraits<V>::HAS) s += 2 * (size_t)TILE * sizeof(VT);
    return s;
}

// Persistent pipelined pass: needs 16-byte aligned keys (TMA) and at least one full tile.
template <class V, class LookT> static int32_t launch_pipe(const OnesweepParams &p, int64_t tiles, cudaStream_t st) {
    auto kern = radix_onesweep_pipe_kernel<K, V, LookT, ItemsFor<V>::value>;
    const size_t smem = pipe_smem<V>();
    static thread_local bool attr_set[64];
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    int64_t per_sm = (int64_t)(220 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 3) per_sm = 3;
    int64_t grid = (int64_t)devinfo().sm_count * per_sm;
    if (grid > tiles) grid = tiles;
    kern<<<(unsigned int)grid, RX_THREADS, smem, st>>>(p, (long long)tiles);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

// 512-thread tiles for keys-only sorts of 32-bit keys (experiment, B200_SORT_NT=512; see radix_onesweep_kernel's NT).
static bool wide_cta() {
#if B200_KEY_BYTES == 4
    static const bool on = [] { const char *e = getenv("B200_SORT_NT"); return e && atoi(e) == 512; }();
    return on;
#else
    return false;
#endif
}
constexpr int WIDE_NT = 512;
constexpr int WIDE_TILE = WIDE_NT * ItemsFor<NoPayload>::value;

// Counter-ranked pass (sort_ctr.cuh) for keys-only sorts of 32-bit keys: 0 = off, 1 = 128 threads x 64 keys,
// 2 = 64 threads x 128 keys.
static int ctr_mode() {
#if B200_KEY_BYTES == 4
    static const int mode = [] { const char *e = getenv("B200_SORT_CTR"); return e ? atoi(e) : 0; }();
    return mode;
#else
    return 0;
#endif
}
constexpr int CTR_TILE = 8192;

#if B200_KEY_BYTES == 4
template <class LookT, int NT, int ITEMS> static int32_t launch_ctr(const OnesweepParams &p, int64_t tiles, cudaStream_t st) {
    auto kern = radix_onesweep_ctr_kernel<LookT, NT, ITEMS>;
    constexpr size_t smem = ctr_smem_bytes<NT, ITEMS>();
    static thread_local bool attr_set[64];
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    kern<<<(unsigned int)tiles, NT, smem, st>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}
#endif

template <class V, class LookT> static int32_t launch_pass(const OnesweepParams &p, int64_t tiles, cudaStream_t st) {
#if B200_KEY_BYTES == 4
    if constexpr (!PayTraits<V>::HAS) {
        if (ctr_mode() == 1) return launch_ctr<LookT, 128, 64>(p, tiles, st);
        if (ctr_mode() == 2) return launch_ctr<LookT, 64, 128>(p, tiles, st);
    }
#endif
    static const bool use_pipe = getenv("B200_SORT_PIPE") != nullptr;   // persistent variant: measured slower (fewer warps per SM)
    if (use_pipe && tiles >= 2 && (reinterpret_cast<uintptr_t>(p.keys_in) % 16) == 0 && pipe_smem<V>() <= 200 * 1024)
        return launch_pipe<V, LookT>(p, tiles, st);
    const size_t smem = onesweep_smem<V>();
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
#if B200_KEY_BYTES == 4
    if constexpr (!PayTraits<V>::HAS) {
        if (wide_cta()) {
            constexpr size_t wsmem = (size_t)(WIDE_NT / 32) * RX_RADIX * 4 + RX_RADIX * 4 + RX_RADIX * 8 + (size_t)WIDE_TILE * sizeof(K);
            auto k0 = radix_onesweep_kernel<K, V, LookT, ItemsFor<V>::value, false, WIDE_NT>;
            auto k1 = radix_onesweep_kernel<K, V, LookT, ItemsFor<V>::value, true, WIDE_NT>;
            static thread_local bool wattr_set[64];
            if (dev >= 0 && dev < 64 && !wattr_set[dev]) {
                B200_CUDA_TRY(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
                B200_CUDA_TRY(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
                wattr_set[dev] = true;
            }
            if (p.unstable_ok) k1<<<(unsigned int)tiles, WIDE_NT, wsmem, st>>>(p);
            else k0<<<(unsigned int)tiles, WIDE_NT, wsmem, st>>>(p);
            B200_CHECK_LAUNCH();
            return B200_OK;
        }
    }
#endif
    if constexpr (!PayTraits<V>::HAS) {
        if (p.unstable_ok) {
            auto akern = radix_onesweep_kernel<K, V, LookT, ItemsFor<V>::value, true>;
            static thread_local bool aattr_set[64];
            if (dev >= 0 && dev < 64 && !aattr_set[dev]) {
                B200_CUDA_TRY(cudaFuncSetAttribute(akern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                aattr_set[dev] = true;
            }
            akern<<<(unsigned int)tiles, RX_THREADS, smem, st>>>(p);
            B200_CHECK_LAUNCH();
            return B200_OK;
        }
    }
    auto kern = radix_onesweep_kernel<K, V, LookT, ItemsFor<V>::value>;
    static thread_local bool attr_set[64];
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    kern<<<(unsigned int)tiles, RX_THREADS, smem, st>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

template <class LookT> static int32_t dispatch_pass(const OnesweepParams &p, int pay_bytes, int64_t tiles, cudaStream_t st) {
    if (pay_bytes == 0) return launch_pass<NoPayload, LookT>(p, tiles, st);
    if (pay_bytes == 4) return launch_pass<uint32_t, LookT>(p, tiles, st);
    return launch_pass<uint64_t, LookT>(p, tiles, st);
}

#define B200_CAT2(a, b) a##b
#define B200_CAT(a, b) B200_CAT2(a, b)

int32_t B200_CAT(sort_run_k, B200_KEY_BYTES)(const SortCall &c) {
    const int64_t n = c.n;
    const int64_t max_tiles = cdiv(n, c.mode == SORT_KEYS ? tile_for<NoPayload>() : tile_for<uint32_t>());
    const int64_t tiles = (c.mode == SORT_KEYS && ctr_mode() && !c.query) ? cdiv(n, CTR_TILE)
                        : (c.mode == SORT_KEYS && wide_cta() && !c.query) ? cdiv(n, WIDE_TILE) : max_tiles;
    // 32-bit status words carry 30-bit counts.  Every prefix that is ever READ belongs to a tile before the last
    // one and is therefore < n, so n == 2^30 still fits (only the last tile's own inclusive value can reach 2^30,
    // and nobody looks back at the last tile).
    const bool wide = n > (1LL << 30);
    const size_t look_bytes = wide ? 8 : 4;
    // payload width carried through the passes
    int pay = 0;
    if (c.mode == SORT_PERM) pay = n > 0xFFFFFFFFLL ? 8 : 4;
    else if (c.mode == SORT_PAIRS) pay = c.payload_bytes;

    Carver w(c.ws);
    unsigned long long *ghist = w.take<unsigned long long>(P * RX_RADIX);
    unsigned long long *bins = w.take<unsigned long long>(P * RX_RADIX);
    unsigned long long *ticket = w.take<unsigned long long>(1);
    char *status = w.take<char>((size_t)max_tiles * RX_RADIX * look_bytes);
    char *kbuf0 = w.take<char>((size_t)n * sizeof(K));
    char *kbuf1 = c.mode == SORT_PERM ? w.take<char>((size_t)n * sizeof(K)) : nullptr;
    char *vbuf0 = pay ? w.take<char>((size_t)n * pay) : nullptr;
    char *vbuf1 = c.mode == SORT_PERM ? w.take<char>((size_t)n * pay) : nullptr;
    if (c.query) { *c.query = w.bytes(); return B200_OK; }
    if (w.bytes() > c.ws_bytes) return B200_WORKSPACE_TOO_SMALL;
    if (c.ws == nullptr) return B200_INVALID_VALUE;
    if (tiles > 2147483647LL) return B200_INVALID_VALUE;
    cudaStream_t st = c.stream;

    // ---- digit histograms of all passes in one read, then their exclusive scans
    B200_CUDA_TRY(cudaMemsetAsync(ghist, 0, (size_t)P * RX_RADIX * sizeof(unsigned long long), st));
    {
        const DevInfo &dev = devinfo();
        int64_t grid = (int64_t)dev.sm_count * 4;
        const int64_t need = cdiv(n, 512 * (16 / (int64_t)sizeof(K)));
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        constexpr size_t hist_smem = (size_t)P * RX_RADIX * RX_HIST_REPL * sizeof(unsigned int);
        static thread_local bool hist_attr[64];
        int hdev = 0;
        B200_CUDA_TRY(cudaGetDevice(&hdev));
        if (hdev >= 0 && hdev < 64 && !hist_attr[hdev]) {
            B200_CUDA_TRY(cudaFuncSetAttribute(radix_hist_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
            hist_attr[hdev] = true;
        }
        radix_hist_kernel<K><<<(unsigned int)grid, 512, hist_smem, st>>>(static_cast<const K *>(c.keys), n, c.xf, ghist);
        B200_CHECK_LAUNCH();
        radix_bins_kernel<<<P, RX_RADIX, 0, st>>>(ghist, bins);
        B200_CHECK_LAUNCH();
    }

    const size_t clear = (size_t)((status + (size_t)tiles * RX_RADIX * look_bytes) - reinterpret_cast<char *>(ticket));
    for (int pass = 0; pass < P; ++pass) {
        OnesweepParams p{};
        p.n = n;
        p.shift = 8 * pass;
        p.bins = bins + pass * RX_RADIX;
        p.status = status;
        p.ticket = ticket;
        p.xf = c.xf;
        p.encode_in = pass == 0;
        static const bool no_atomic_rank = getenv("B200_SORT_STABLE_FIRST") != nullptr;   // A/B switch
        p.unstable_ok = pass == 0 && c.mode == SORT_KEYS && !no_atomic_rank;
        p.decode_out = (pass == P - 1) && c.mode != SORT_PERM;
        p.store_keys = !(pass == P - 1 && c.mode == SORT_PERM);
        p.gen_payload = (pass == 0 && c.mode == SORT_PERM);
        p.index_dtype = (pass == P - 1 && c.mode == SORT_PERM) ? c.index_dtype : -1;
        p.index_base = c.index_base;
        if (c.mode == SORT_PERM) {
            char *kb[2] = {kbuf0, kbuf1};
            char *vb[2] = {vbuf0, vbuf1};
            p.keys_in = pass == 0 ? c.keys : kb[(pass - 1) & 1];
            p.keys_out = kb[pass & 1];
            p.vals_in = pass == 0 ? nullptr : vb[(pass - 1) & 1];
            p.vals_out = pass == P - 1 ? c.payload : vb[pass & 1];
        } else {
            // ping-pong between the caller's buffer and the workspace copy
            const bool from_user = (pass & 1) == 0;
            p.keys_in = from_user ? c.keys : kbuf0;
            p.keys_out = from_user ? kbuf0 : c.keys;
            if (pay) {
                p.vals_in = from_user ? c.payload : vbuf0;
                p.vals_out = from_user ? vbuf0 : c.payload;
            }
        }
        B200_CUDA_TRY(cudaMemsetAsync(ticket, 0, clear, st));
        int32_t rc = wide ? dispatch_pass<uint64_t>(p, pay, tiles, st) : dispatch_pass<uint32_t>(p, pay, tiles, st);
        if (rc != B200_OK) return rc;
    }
    if (c.mode != SORT_PERM && (P & 1)) {
        // odd number of passes: the result sits in the workspace copy
        B200_CUDA_TRY(cudaMemcpyAsync(c.keys, kbuf0, (size_t)n * sizeof(K), cudaMemcpyDeviceToDevice, st));
        if (pay) B200_CUDA_TRY(cudaMemcpyAsync(c.payload, vbuf0, (size_t)n * pay, cudaMemcpyDeviceToDevice, st));
    }
    return B200_OK;
}


template <bool PERM, int ITEMS> static int32_t launch_block_sort(const BlockSortParams &bp, int64_t nseg, cudaStream_t st) {
    constexpr int TILE = RX_THREADS * ITEMS;
    const size_t smem = (size_t)RX_WARPS * RX_RADIX * 4 + (size_t)TILE * sizeof(K) + (PERM ? (size_t)TILE * 2 : 0);
    auto kern = radix_block_sort_kernel<K, PERM, ITEMS>;
    static thread_local bool attr_set[64];
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[dev] = true;
    }
    kern<<<(unsigned int)nseg, RX_THREADS, smem, st>>>(bp);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

int32_t B200_CAT(block_sort_k, B200_KEY_BYTES)(const BlockSortParams &bp, int perm, int64_t nseg, cudaStream_t st) {
    if (nseg == 0) return B200_OK;
    if (nseg > 2147483647LL) return B200_INVALID_VALUE;
    const bool small = bp.n <= RX_THREADS * 4;
    if (perm) return small ? launch_block_sort<true, 4>(bp, nseg, st) : launch_block_sort<true, RB_ITEMS>(bp, nseg, st);
    return small ? launch_block_sort<false, 4>(bp, nseg, st) : launch_block_sort<false, RB_ITEMS>(bp, nseg, st);
}

// k-th smallest key of `keys` (0-based k) -> out[0]; workspace = SelectState + one 256-bin histogram per pass.
int32_t B200_CAT(select_kth_k, B200_KEY_BYTES)(void *ws, size_t wsb, const void *keys, int64_t n, int64_t k,
                                               const SortXform &xf, void *out, cudaStream_t st, size_t *query) {
    Carver w(ws);
    SelectState *state = w.take<SelectState>(1);
    unsigned int *hist = w.take<unsigned int>(P * RX_RADIX);
    if (query) { *query = w.bytes(); return B200_OK; }
    if (w.bytes() > wsb) return B200_WORKSPACE_TOO_SMALL;
    if (ws == nullptr) return B200_INVALID_VALUE;
    B200_CUDA_TRY(cudaMemsetAsync(ws, 0, w.bytes(), st));
    const DevInfo &dev = devinfo();
    int64_t grid = (int64_t)dev.sm_count * 4;
    const int64_t need = cdiv(n, 512 * (16 / (int64_t)sizeof(K)));
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    for (int pass = 0; pass < P; ++pass) {
        const int shift = 8 * (P - 1 - pass);
        unsigned int *h = hist + pass * RX_RADIX;
        select_hist_kernel<K><<<(unsigned int)grid, 512, 0, st>>>(static_cast<const K *>(keys), n, xf, shift, state, h);
        B200_CHECK_LAUNCH();
        select_pick_kernel<K><<<1, RX_RADIX, 0, st>>>(h, state, shift, pass == 0, (unsigned long long)k, pass == P - 1, xf,
                                                      static_cast<K *>(out));
        B200_CHECK_LAUNCH();
    }
    return B200_OK;
}

int32_t B200_CAT(lower_bound_k, B200_KEY_BYTES)(const void *sorted, int64_t n, const void *needles, int64_t m,
                                               const SortXform &xf, long long *out, cudaStream_t st) {
    if (m == 0) return B200_OK;
    radix_lower_bound_kernel<K><<<(unsigned int)cdiv(m, 128), 128, 0, st>>>(static_cast<const K *>(sorted), n,
                                                                           static_cast<const K *>(needles), m, xf, out);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // namespace b200
