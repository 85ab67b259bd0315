// sort_inst.cu — one translation unit per key width (-DB200_KEY_BYTES=1|2|4|8).
// Host orchestration: LSD onesweep passes for keys / sortperm / pairs, and (32-bit keys-only) the hybrid MSD
// path of sort_msd.cuh with the LSD passes as its device-gated fallback.
#include "select_kth.cuh"
#include "sort_block.cuh"
#include "sort_host.h"
#if B200_KEY_BYTES == 4
#include "sort_msd.cuh"
#endif

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

// One-time (per kernel, per device) setup: opt in to the dynamic shared memory and ask how many CTAs fit on an SM —
// the passes are persistent (tiles by ticket), so the grid is resident CTAs, not tiles.
struct KernelSetup {
    bool done[64] = {};
    int ctas_per_sm[64] = {};
};
template <class Kern> static int32_t setup_kernel(KernelSetup &ks, Kern kern, int threads, size_t smem, int *ctas_per_sm) {
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return B200_INVALID_VALUE;
    if (!ks.done[dev]) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem));
        ks.ctas_per_sm[dev] = nb < 1 ? 1 : nb;
        ks.done[dev] = true;
    }
    *ctas_per_sm = ks.ctas_per_sm[dev];
    return B200_OK;
}

template <class Kern> static int32_t launch_pass_kernel(KernelSetup &ks, Kern kern, int threads, size_t smem, bool persistent,
                                                        const OnesweepParams &p, cudaStream_t st) {
    int per_sm = 1;
    const int32_t rc = setup_kernel(ks, kern, threads, smem, &per_sm);
    if (rc != B200_OK) return rc;
    int64_t grid = p.tiles;
    if (persistent && grid > (int64_t)devinfo().sm_count * per_sm) grid = (int64_t)devinfo().sm_count * per_sm;
    kern<<<(unsigned int)grid, threads, smem, st>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

template <class V, class LookT> static int32_t launch_pass(const OnesweepParams &p, cudaStream_t st) {
    const size_t smem = onesweep_smem<V>();
    constexpr int IT = ItemsFor<V>::value;
#if B200_KEY_BYTES == 4
    if constexpr (!PayTraits<V>::HAS && sizeof(LookT) == 4) {
        if (p.gate) {       // fallback of the hybrid MSD path: persistent CTAs (an all-returning grid must be small)
            static thread_local KernelSetup ks0, ks1;
            if (p.unstable_ok)
                return launch_pass_kernel(ks1, radix_onesweep_kernel<K, V, LookT, IT, true, RX_THREADS, true>, RX_THREADS, smem, true, p, st);
            return launch_pass_kernel(ks0, radix_onesweep_kernel<K, V, LookT, IT, false, RX_THREADS, true>, RX_THREADS, smem, true, p, st);
        }
    }
#endif
    if constexpr (!PayTraits<V>::HAS) {
        if (p.unstable_ok) {
            static thread_local KernelSetup ks;
            return launch_pass_kernel(ks, radix_onesweep_kernel<K, V, LookT, IT, true>, RX_THREADS, smem, false, p, st);
        }
    }
    static thread_local KernelSetup ks;
    return launch_pass_kernel(ks, radix_onesweep_kernel<K, V, LookT, IT>, RX_THREADS, smem, false, p, st);
}

template <class LookT> static int32_t dispatch_pass(const OnesweepParams &p, int pay_bytes, cudaStream_t st) {
    if (pay_bytes == 0) return launch_pass<NoPayload, LookT>(p, st);
    if (pay_bytes == 4) return launch_pass<uint32_t, LookT>(p, st);
    return launch_pass<uint64_t, LookT>(p, st);
}

#define B200_CAT2(a, b) a##b
#define B200_CAT(a, b) B200_CAT2(a, b)

#if B200_KEY_BYTES == 4
// zero-fill that only runs when the hybrid path declined the call (count is in 16-byte units; 256-byte aligned buffers)
static __global__ void gated_zero_kernel(uint4 *p, size_t count, const unsigned int *gate) {
    if (ld_gate(gate) != 0) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0, 0, 0, 0);
}

// ---- hybrid MSD path (sort_msd.cuh): dense 32-bit keys-only sorts -------------------------------------------
#ifdef B200_TUNING
constexpr int MSD_TILE_MIN = 4096;            // smallest P1 / P2 tile of any tuning configuration (sizes the look-back arrays)
#else
constexpr int MSD_TILE_MIN = 512 * 20;        // the P1 / P2 tile
#endif
constexpr int LEAF_NT = 512;
constexpr int LEAF16_NT = 1024;              // the 128 KB table leaves one CTA per SM: make it a full one
constexpr unsigned int LEAF8_CAP = 1u << 18;  // segments up to this many keys (4 per value on average; a bin holds 255) go to the 8-bit-bin leaf first
constexpr unsigned int LEAF_MAX_SEG = 65535; // always admitted; beyond it a segment may hold at most n / 64 keys (one CTA sorts it)
constexpr unsigned int MSD_MIN_AVG = 768;    // hybrid path: at least this many keys per non-empty 16-bit segment on average
constexpr unsigned int MSD_DENSE_AVG = 6144; // from this average on: counting leaves only (they walk 65536 bins per segment whatever its
                                             // length); below it the short segments are sorted on chip (msd_leaf_block_kernel)
constexpr int64_t MSD_MAX_N = 1LL << 30;     // 32-bit look-back words carry 30-bit counts

#ifdef B200_TUNING
// P1 / P2 tile shape: 0 = 512 threads x 20 keys (default, the only one in the product build); measured on B200, 2^30
// uniform UInt32: 512x20 12.2 ms, 512x16 13.2 ms, 1024x10 14.3 ms, 256x20 19.1 ms, 256x16 25.4 ms
static int msd_cfg() {
    static const int cfg = tune_int("B200_MSD_CFG", 0);
    return cfg;
}
#endif

static unsigned int leaf8_cap() {
    static const unsigned int cap = (unsigned int)tune_int("B200_MSD_CAP8", (int)LEAF8_CAP);   // tuning builds: small caps exercise the 16-bit leaf on small inputs
    return cap;
}

// Tile of the on-chip leaf for a sort of n keys (0: not used): the smallest instantiation that holds 1.25 x the average
// segment of keys spread over all 65536 segments.  Segments longer than the tile go to the counting leaves.
static int block_leaf_tile(int64_t n) {
    const int64_t want = (n >> 16) + (n >> 18);
    if (want <= RX_THREADS * 6) return RX_THREADS * 6;
    if (want <= RX_THREADS * 12) return RX_THREADS * 12;
    if (want <= RX_THREADS * 20) return RX_THREADS * 20;
    return 0;
}

static bool xform_is_identity(const SortXform &x) { return x.base_enc == 0 && x.fmn == 0 && x.mag_mask == 0 && x.desc == 0; }

struct MsdBuffers {
    MsdCtl *ctl;
    unsigned int *hist16, *status1, *status2, *list16, *spill, *seg_start, *seg_list, *l1_tile_start;
};

// MSD_ITEMS: keys per thread of the heavy-digit instantiation of P1 / P2 (its ranking sweeps need the registers);
// MSD_ITEMS_PLAIN: of the plain one — fewer, larger tiles mean fewer look-backs and barriers per key.  Measured at 2^30
// uniform keys: 20 keys per thread 10.50 ms, 24: 10.21, 28: 9.94 (56 KB of staging per CTA, two CTAs per SM, 24 bytes of
// spills in P2), 32: 9.91 but presorted input 13.2 ms (spills in the row-uniform path); the heavy instantiation lost 4 %
// already at 24 (Float32 rand 11.0 -> 11.5 ms).
template <bool IDENT, int MSD_NT = 512, int MSD_ITEMS = 20, int MSD_ITEMS_PLAIN = MSD_ITEMS == 20 ? 28 : MSD_ITEMS>
static int32_t msd_launch(const SortCall &c, const MsdBuffers &b, uint32_t *alt, cudaStream_t st) {
    constexpr int MSD_TILE = MSD_NT * MSD_ITEMS;
    constexpr int MSD_TILE_PLAIN = MSD_NT * MSD_ITEMS_PLAIN;
    const int64_t n = c.n;
    const DevInfo &dev = devinfo();
    const void *src = c.keys_src ? c.keys_src : c.keys;
    if (!c.hist16_in) {   // H: histogram of the top 16 bits
        static thread_local KernelSetup ks;
        int per_sm = 1;
        const size_t smem = (size_t)MSD_H16_WORDS * 4;
        int32_t rc = setup_kernel(ks, msd_hist16_kernel<IDENT>, MSD_H16_THREADS, smem, &per_sm);
        if (rc != B200_OK) return rc;
        int64_t grid = dev.sm_count;
        const int64_t need = cdiv(n, (int64_t)MSD_H16_THREADS * MSD_H16_VPT * 4);
        if (grid > need) grid = need;
        msd_hist16_kernel<IDENT><<<(unsigned int)grid, MSD_H16_THREADS, smem, st>>>(static_cast<const uint32_t *>(src), n, c.xf, b.hist16);
        B200_CHECK_LAUNCH();
    }
    {   // PL
        MsdPlanParams pp{};
        pp.hist16 = c.hist16_in ? static_cast<const unsigned int *>(c.hist16_in) : b.hist16; pp.seg_start = b.seg_start; pp.seg_list = b.seg_list; pp.l1_tile_start = b.l1_tile_start;
        pp.list16 = b.list16;
        pp.ctl = b.ctl; pp.n = (unsigned int)n; pp.tile2 = MSD_TILE; pp.tile2_plain = MSD_TILE_PLAIN;
        pp.cap8 = leaf8_cap();
        pp.cap_a = (unsigned int)(n >> 6) > LEAF_MAX_SEG ? (unsigned int)(n >> 6) : LEAF_MAX_SEG;
        static const unsigned int min_avg = (unsigned int)tune_int("B200_MSD_MIN_AVG", (int)MSD_MIN_AVG);
        static const unsigned int dense_avg = (unsigned int)tune_int("B200_MSD_DENSE_AVG", (int)MSD_DENSE_AVG);
        pp.min_avg = min_avg;
        pp.dense_avg = dense_avg;
        pp.small_cap = (unsigned int)block_leaf_tile(n);
        msd_plan_kernel<<<1, 1024, 0, st>>>(pp);
        B200_CHECK_LAUNCH();
    }
    MsdScatterParams sp{};
    sp.n = (unsigned int)n; sp.seg_start = b.seg_start; sp.l1_tile_start = b.l1_tile_start; sp.ctl = b.ctl; sp.xf = c.xf;
    // P1: keys -> alt by byte 3; P2: alt -> keys by byte 2 inside every level-1 bucket.  Each level is launched in both
    // instantiations; the plan's ctl->heavy bit lets exactly one of them run (the other returns at once).
    auto launch_level = [&](auto kern, KernelSetup &ks, int64_t tiles, size_t ssmem) -> int32_t {
        int per_sm = 1;
        int32_t rc = setup_kernel(ks, kern, MSD_NT, ssmem, &per_sm);
        if (rc != B200_OK) return rc;
        int64_t grid = (int64_t)dev.sm_count * per_sm;
        if (grid > tiles) grid = tiles;
        kern<<<(unsigned int)grid, MSD_NT, ssmem, st>>>(sp);
        B200_CHECK_LAUNCH();
        return B200_OK;
    };
    constexpr size_t SS_H = (size_t)MSD_TILE * 4, SS_P = (size_t)MSD_TILE_PLAIN * 4;
    {
        static thread_local KernelSetup ks_a, ks_b;
        sp.keys_in = static_cast<const uint32_t *>(src); sp.keys_out = alt; sp.status = b.status1;
        int32_t rc = launch_level(msd_scatter_kernel<1, MSD_NT, MSD_ITEMS_PLAIN, IDENT, false>, ks_a, cdiv(n, MSD_TILE_PLAIN), SS_P);
        if (rc != B200_OK) return rc;
        rc = launch_level(msd_scatter_kernel<1, MSD_NT, MSD_ITEMS, IDENT, true>, ks_b, cdiv(n, MSD_TILE), SS_H);
        if (rc != B200_OK) return rc;
    }
    {
        static thread_local KernelSetup ks_a, ks_b;
        sp.keys_in = alt; sp.keys_out = static_cast<uint32_t *>(c.keys); sp.status = b.status2;
        int32_t rc = launch_level(msd_scatter_kernel<2, MSD_NT, MSD_ITEMS_PLAIN, IDENT, false>, ks_a, cdiv(n, MSD_TILE_PLAIN) + 256, SS_P);
        if (rc != B200_OK) return rc;
        rc = launch_level(msd_scatter_kernel<2, MSD_NT, MSD_ITEMS, IDENT, true>, ks_b, cdiv(n, MSD_TILE) + 256, SS_H);
        if (rc != B200_OK) return rc;
    }
    MsdLeafParams lp{};
    lp.keys = static_cast<uint32_t *>(c.keys); lp.seg_start = b.seg_start; lp.seg_list = b.seg_list; lp.ctl = b.ctl; lp.xf = c.xf;
    lp.list16 = b.list16; lp.spill = b.spill; lp.cap8 = leaf8_cap();
    if (const int bt = block_leaf_tile(n)) {   // short segments of inputs that are sparse at 16 bits (gated on ctl->small_cap)
        auto go = [&](auto kern, KernelSetup &ks) -> int32_t {
            const size_t smem = (size_t)RX_WARPS * RX_RADIX * 4 + (size_t)bt * 4;
            int per_sm = 1;
            int32_t rc = setup_kernel(ks, kern, RX_THREADS, smem, &per_sm);
            if (rc != B200_OK) return rc;
            kern<<<(unsigned int)(dev.sm_count * per_sm), RX_THREADS, smem, st>>>(lp);
            B200_CHECK_LAUNCH();
            return B200_OK;
        };
        static thread_local KernelSetup ks6, ks12, ks20;
        int32_t rc = bt == RX_THREADS * 6 ? go(msd_leaf_block_kernel<6, IDENT>, ks6)
                   : bt == RX_THREADS * 12 ? go(msd_leaf_block_kernel<12, IDENT>, ks12) : go(msd_leaf_block_kernel<20, IDENT>, ks20);
        if (rc != B200_OK) return rc;
    }
    {   // 8-bit bins: segments up to LEAF8_CAP keys without heavy duplicates
        static thread_local KernelSetup ks;
        int per_sm = 1;
        auto kern = msd_leaf_kernel<LEAF_NT, IDENT>;
        const size_t smem = (size_t)MSD_BINS;
        int32_t rc = setup_kernel(ks, kern, LEAF_NT, smem, &per_sm);
        if (rc != B200_OK) return rc;
        kern<<<(unsigned int)(dev.sm_count * per_sm), LEAF_NT, smem, st>>>(lp);
        B200_CHECK_LAUNCH();
    }
    {   // 16-bit bins: what the 8-bit kernel skipped or gave up on
        static thread_local KernelSetup ks;
        int per_sm = 1;
        auto kern = msd_leaf16_kernel<LEAF16_NT, IDENT>;
        const size_t smem = (size_t)MSD_BINS * 2;
        int32_t rc = setup_kernel(ks, kern, LEAF16_NT, smem, &per_sm);
        if (rc != B200_OK) return rc;
        kern<<<(unsigned int)dev.sm_count, LEAF16_NT, smem, st>>>(lp);      // one spill table per CTA: grid == sm_count
        B200_CHECK_LAUNCH();
    }
    return B200_OK;
}

// The workspace LAYOUT carries the hybrid path's buffers for every keys-only call it could ever be asked to run
// (so that the size query does not depend on whether the caller will pass a histogram); whether the path is attempted
// is decided per call.
static bool msd_layout(const SortCall &c) {
    static const int on = tune_int("B200_SORT_HYBRID", 1);
    static const int min_log2 = tune_int("B200_MSD_MIN_LOG2N", 27);
    const int64_t lay_min = min_log2 < 24 ? (1LL << min_log2) : (1LL << 24);
    return on && c.mode == SORT_KEYS && c.n >= lay_min && c.n <= MSD_MAX_N;
}
static bool msd_eligible(const SortCall &c) {
    static const int min_log2 = tune_int("B200_MSD_MIN_LOG2N", 27);
    // with a caller-supplied histogram (multi-GPU slice: dense inside a narrow key range) the plan kernel's density
    // test decides alone, and the attempt costs no extra pass
    const int64_t min_n = c.hist16_in ? (1LL << 24) : (1LL << min_log2);
    return msd_layout(c) && c.n >= min_n;
}
#else
static bool msd_layout(const SortCall &) { return false; }
static bool msd_eligible(const SortCall &) { return false; }
#endif

int32_t B200_CAT(sort_run_k, B200_KEY_BYTES)(const SortCall &c) {
    const int64_t n = c.n;
    const int64_t tiles = cdiv(n, c.mode == SORT_KEYS ? tile_for<NoPayload>() : tile_for<uint32_t>());
    // 32-bit status words carry 30-bit counts.  Every prefix that is ever READ belongs to a tile before the last
    // one and is therefore < n, so n == 2^30 still fits (only the last tile's own inclusive value can reach 2^30,
    // and nobody looks back at the last tile).
    const bool wide = n > (1LL << 30);
    const size_t look_bytes = wide ? 8 : 4;
    // payload width carried through the passes
    int pay = 0;
    if (c.mode == SORT_PERM) pay = n > 0xFFFFFFFFLL ? 8 : 4;
    else if (c.mode == SORT_PAIRS) pay = c.payload_bytes;
    const bool hybrid_layout = msd_layout(c);
    const bool hybrid = msd_eligible(c);

    // workspace: [ctl | ghist | (hybrid: hist16 | list16 | status2) | status] — everything up to the part of `status`
    // that the first kernels need is zeroed by ONE memset per call; when the hybrid path is attempted, the rest of
    // `status` (only the LSD fallback reads it) is zeroed by a kernel gated on the fallback
    Carver w(c.ws);
    char *ctl_raw = w.take<char>(256);
    unsigned long long *ghist = w.take<unsigned long long>(P * RX_RADIX);
#if B200_KEY_BYTES == 4
    MsdBuffers mb{};
    mb.ctl = reinterpret_cast<MsdCtl *>(ctl_raw);
    static_assert(sizeof(MsdCtl) <= 256, "control block");
    if (hybrid_layout) {
        mb.hist16 = w.take<unsigned int>(MSD_BINS);
        mb.list16 = w.take<unsigned int>(MSD_BINS);
        mb.status2 = w.take<unsigned int>((size_t)(cdiv(n, MSD_TILE_MIN) + 256) * RX_RADIX);
    }
#endif
    // P1 of the hybrid path shares the head of the LSD look-back array (only one of the two paths runs)
    size_t status_bytes = (size_t)tiles * RX_RADIX * look_bytes;
    size_t status_head = status_bytes;
#if B200_KEY_BYTES == 4
    if (hybrid_layout) {
        status_head = (size_t)cdiv(n, MSD_TILE_MIN) * RX_RADIX * sizeof(unsigned int);
        if (status_bytes < status_head) status_bytes = status_head;
    }
#endif
    const size_t status_off = w.bytes();
    char *status = w.take<char>(status_bytes);
    const size_t clear_bytes = hybrid ? status_off + status_head : w.bytes();
#if B200_KEY_BYTES == 4
    if (hybrid_layout) {
        mb.status1 = reinterpret_cast<unsigned int *>(status);
        mb.seg_start = w.take<unsigned int>(MSD_BINS + 1);
        mb.seg_list = w.take<unsigned int>(MSD_BINS);
        mb.l1_tile_start = w.take<unsigned int>(257);
        mb.spill = w.take<unsigned int>((size_t)devinfo().sm_count * MSD_BINS);   // zeroed by the 16-bit leaf itself
    }
#endif
    unsigned long long *bins = w.take<unsigned long long>(P * RX_RADIX);
    char *kbuf0 = w.take<char>((size_t)n * sizeof(K));
    char *kbuf1 = c.mode == SORT_PERM ? w.take<char>((size_t)n * sizeof(K)) : nullptr;
    char *vbuf0 = pay ? w.take<char>((size_t)n * pay) : nullptr;
    char *vbuf1 = c.mode == SORT_PERM ? w.take<char>((size_t)n * pay) : nullptr;
    if (c.query) { *c.query = w.bytes(); return B200_OK; }
    if (w.bytes() > c.ws_bytes) return B200_WORKSPACE_TOO_SMALL;
    if (c.ws == nullptr) return B200_INVALID_VALUE;
    if (tiles > 2147483647LL) return B200_INVALID_VALUE;
    cudaStream_t st = c.stream;
    SortCtl *ctl = reinterpret_cast<SortCtl *>(ctl_raw);
    static_assert(P <= 8 && sizeof(SortCtl) <= 256, "one LSD ticket per pass; control block is the first 256 bytes");

    B200_CUDA_TRY(cudaMemsetAsync(c.ws, 0, clear_bytes, st));
    const unsigned int *gate = nullptr;
#if B200_KEY_BYTES == 4
    if (hybrid) {
        uint32_t *alt = reinterpret_cast<uint32_t *>(kbuf0);
        int32_t rc;
#ifdef B200_TUNING
        if (msd_cfg() == 1) rc = xform_is_identity(c.xf) ? msd_launch<true, 256, 20>(c, mb, alt, st) : msd_launch<false, 256, 20>(c, mb, alt, st);
        else if (msd_cfg() == 2) rc = xform_is_identity(c.xf) ? msd_launch<true, 256, 16>(c, mb, alt, st) : msd_launch<false, 256, 16>(c, mb, alt, st);
        else if (msd_cfg() == 3) rc = xform_is_identity(c.xf) ? msd_launch<true, 1024, 10>(c, mb, alt, st) : msd_launch<false, 1024, 10>(c, mb, alt, st);
        else if (msd_cfg() == 4) rc = xform_is_identity(c.xf) ? msd_launch<true, 512, 16>(c, mb, alt, st) : msd_launch<false, 512, 16>(c, mb, alt, st);
        else if (msd_cfg() == 5) rc = xform_is_identity(c.xf) ? msd_launch<true, 512, 12>(c, mb, alt, st) : msd_launch<false, 512, 12>(c, mb, alt, st);
        else if (msd_cfg() == 6) rc = xform_is_identity(c.xf) ? msd_launch<true, 512, 24>(c, mb, alt, st) : msd_launch<false, 512, 24>(c, mb, alt, st);
        else
#endif
        rc = xform_is_identity(c.xf) ? msd_launch<true>(c, mb, alt, st) : msd_launch<false>(c, mb, alt, st);
        if (rc != B200_OK) return rc;
        gate = &ctl->mode;      // the LSD kernels below return at once when the hybrid path took the call
        if (status_bytes > status_head) {
            gated_zero_kernel<<<devinfo().sm_count * 4, 512, 0, st>>>(reinterpret_cast<uint4 *>(status + status_head),
                                                                        (status_bytes - status_head) / 16, gate);
            B200_CHECK_LAUNCH();
        }
    }
#endif

    // ---- digit histograms of all passes in one read, then their exclusive scans
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
        radix_hist_kernel<K><<<(unsigned int)grid, 512, hist_smem, st>>>(static_cast<const K *>(c.keys_src ? c.keys_src : c.keys), n, c.xf, ghist, gate);
        B200_CHECK_LAUNCH();
        radix_bins_kernel<<<P, RX_RADIX, 0, st>>>(ghist, bins, gate);
        B200_CHECK_LAUNCH();
    }

    const int flag_shift = wide ? 62 : 30;
    for (int pass = 0; pass < P; ++pass) {
        OnesweepParams p{};
        p.n = n;
        p.shift = 8 * pass;
        p.bins = bins + pass * RX_RADIX;
        p.status = status;
        p.ticket = &ctl->lsd_ticket[pass];
        p.tiles = tiles;
        p.gate = gate;
        p.flag_incl = (unsigned long long)(1 + pass % 3) << flag_shift;
        p.flag_part = (unsigned long long)(1 + (pass + 1) % 3) << flag_shift;
        p.xf = c.xf;
        p.encode_in = pass == 0;
        static const int stable_first = tune_int("B200_SORT_STABLE_FIRST", 0);   // A/B switch
        p.unstable_ok = pass == 0 && c.mode == SORT_KEYS && !stable_first;
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
            p.keys_in = from_user ? (pass == 0 && c.keys_src ? c.keys_src : c.keys) : kbuf0;
            p.keys_out = from_user ? kbuf0 : c.keys;
            if (pay) {
                p.vals_in = from_user ? c.payload : vbuf0;
                p.vals_out = from_user ? vbuf0 : c.payload;
            }
        }
        int32_t rc = wide ? dispatch_pass<uint64_t>(p, pay, st) : dispatch_pass<uint32_t>(p, pay, st);
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

// k-th smallest key of `keys` (0-based k) -> out[0]; workspace = SelectState + one 256-bin histogram per pass
// (+ for large n: the sampled fast path's bracket, a second set of histograms and the candidate buffer of n / 8 keys).
constexpr int64_t SELECT_FAST_MIN_N = 1LL << 22;
int32_t B200_CAT(select_kth_k, B200_KEY_BYTES)(void *ws, size_t wsb, const void *keys, int64_t n, int64_t k,
                                               const SortXform &xf, void *out, cudaStream_t st, size_t *query) {
    static const int fast_on = tune_int("B200_SELECT_FAST", 1);
    const bool fastp = fast_on && n >= SELECT_FAST_MIN_N && sizeof(K) >= 2;
    const unsigned long long cap = fastp ? (unsigned long long)(n / 8) : 0;
    Carver w(ws);
    SelectState *state = w.take<SelectState>(1);
    unsigned int *hist = w.take<unsigned int>(P * RX_RADIX);
    SelectState *cstate = fastp ? w.take<SelectState>(1) : nullptr;
    unsigned int *chist = fastp ? w.take<unsigned int>(P * RX_RADIX) : nullptr;
    SelectFast *fast = fastp ? w.take<SelectFast>(1) : nullptr;
    const size_t clear = w.bytes();
    K *cand = fastp ? w.take<K>((size_t)cap) : nullptr;
    if (query) { *query = w.bytes(); return B200_OK; }
    if (w.bytes() > wsb) return B200_WORKSPACE_TOO_SMALL;
    if (ws == nullptr) return B200_INVALID_VALUE;
    B200_CUDA_TRY(cudaMemsetAsync(ws, 0, clear, st));
    const DevInfo &dev = devinfo();
    int64_t grid = (int64_t)dev.sm_count * 4;
    const int64_t need = cdiv(n, 512 * (16 / (int64_t)sizeof(K)));
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if (fastp) {
        select_sample_kernel<K><<<1, 1024, 0, st>>>(static_cast<const K *>(keys), n, (unsigned long long)k, xf, fast);
        B200_CHECK_LAUNCH();
        select_filter_kernel<K><<<(unsigned int)grid, 512, 0, st>>>(static_cast<const K *>(keys), n, xf, fast, cand, cap);
        B200_CHECK_LAUNCH();
        select_decide_kernel<<<1, 1, 0, st>>>(fast, cstate, (unsigned long long)k, cap);
        B200_CHECK_LAUNCH();
        // byte passes over the candidates (already in ordered-key space: identity transform; the count is on the device)
        SortXform ident{};
        ident.inf_bits = ~0ull;
        int64_t cgrid = (int64_t)dev.sm_count * 2;
        const int64_t cneed = cdiv((int64_t)cap, 512 * (16 / (int64_t)sizeof(K)));
        if (cgrid > cneed) cgrid = cneed;
        if (cgrid < 1) cgrid = 1;
        for (int pass = 0; pass < P; ++pass) {
            const int shift = 8 * (P - 1 - pass);
            unsigned int *h = chist + pass * RX_RADIX;
            select_hist_kernel<K><<<(unsigned int)cgrid, 512, 0, st>>>(cand, 0, ident, shift, cstate, h, fast, 1, true);
            B200_CHECK_LAUNCH();
            select_pick_kernel<K><<<1, RX_RADIX, 0, st>>>(h, cstate, shift, 0, 0ull, pass == P - 1, xf, static_cast<K *>(out), fast, 1);
            B200_CHECK_LAUNCH();
        }
    }
    for (int pass = 0; pass < P; ++pass) {
        const int shift = 8 * (P - 1 - pass);
        unsigned int *h = hist + pass * RX_RADIX;
        select_hist_kernel<K><<<(unsigned int)grid, 512, 0, st>>>(static_cast<const K *>(keys), n, xf, shift, state, h, fast, fastp ? 2 : 0, false);
        B200_CHECK_LAUNCH();
        select_pick_kernel<K><<<1, RX_RADIX, 0, st>>>(h, state, shift, pass == 0, (unsigned long long)k, pass == P - 1, xf,
                                                      static_cast<K *>(out), fast, fastp ? 2 : 0);
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
