"""Model check (CPU, no GPU) of the self-chained prefix rule of scan_chain2_kernel
(cuda.jl_b200/csrc/scan.cuh, chain_issue / chain_consume):

    a persistent CTA owns tickets g_0 < g_1 < ...;  for tile g = g_{j+1} of segment s (first tile seg_start)
      t == 0                      -> prefix = seed                                   (pf_base 0, empty range)
      g_j in the same segment     -> prefix = inclusive(g_j) (+) sum partial(g_j+1 .. g-1)      (pf_base 2)
      otherwise                   -> prefix = sum published(seg_start .. g-1)                   (pf_base 1)
    where tile 0 of a segment publishes seed (+) aggregate and every other tile its plain aggregate.

The kernel's arithmetic is restated here over integers for ANY assignment of tickets to CTAs (tickets are handed
out by an atomic counter, so the assignment is timing-dependent) and compared with a plain exclusive scan with
`init` folded in once per element, which is what CUDACore/src/accumulate.jl:90-92,123-125 computes."""
import numpy as np
import pytest


def chained_prefixes(aggs, tiles_per_seg, owners, seed):
    """aggs[g]: aggregate of global tile g; owners[g]: CTA that drew ticket g.  Returns the exclusive tile prefixes."""
    total = len(aggs)
    published = [None] * total
    for g in range(total):
        t = g % tiles_per_seg
        published[g] = seed + aggs[g] if t == 0 else aggs[g]          # finish_pass1
    prefix = [None] * total
    last = {}                                                          # per CTA: (previous ticket, its inclusive prefix)
    for g in range(total):                                             # a CTA sees its own tickets in increasing order
        cta = owners[g]
        seg_start = g - g % tiles_per_seg
        if g == seg_start:
            base, lo, hi = seed, 0, -1
        elif cta in last and last[cta][0] >= seg_start:
            base, lo, hi = last[cta][1], last[cta][0] + 1, g - 1
        else:
            base, lo, hi = 0, seg_start, g - 1
        prefix[g] = base + sum(published[u] for u in range(lo, hi + 1))
        last[cta] = (g, prefix[g] + aggs[g])                           # incl_own
    return prefix


@pytest.mark.parametrize("nseg,tiles_per_seg,nctas", [(1, 1, 1), (1, 40, 4), (1, 1000, 148), (7, 13, 5), (64, 2, 148),
                                                      (3, 500, 296), (200, 1, 148)])
@pytest.mark.parametrize("seed", [0, 7])
def test_chain_rule_matches_exclusive_scan(nseg, tiles_per_seg, nctas, seed):
    rng = np.random.default_rng(nseg * 1000 + tiles_per_seg + nctas)
    total = nseg * tiles_per_seg
    aggs = [int(x) for x in rng.integers(-1000, 1000, size=total)]
    for trial in range(4):
        if trial == 0:
            owners = [g % nctas for g in range(total)]                 # lockstep
        elif trial == 1:
            owners = [(g // 3) % nctas for g in range(total)]          # prologue-like runs of consecutive tickets
        else:
            owners = [int(x) for x in rng.integers(0, nctas, size=total)]   # arbitrary interleaving, uneven gaps
        got = chained_prefixes(aggs, tiles_per_seg, owners, seed)
        for s in range(nseg):
            run = seed
            for t in range(tiles_per_seg):
                g = s * tiles_per_seg + t
                assert got[g] == run, (s, t, trial)
                run += aggs[g]


def test_tiling_choice_prefers_exact_fit():
    """Host rule of launch_scan for several segments: chain2's tile unless the look-back kernel's tile leaves less of a
    segment in a ragged last tile (and more than 12.5 % would be ragged)."""
    def chain2_fits(n, post, tile2, tile1):
        pad2 = -(-n // tile2) * tile2 - n
        pad1 = -(-n // tile1) * tile1 - n
        return post == 1 or pad2 <= pad1 or pad2 * 8 <= n
    t2, t1 = 480 * 4 * 7, 512 * 4 * 8                                   # Float32: 13440 and 16384 elements
    assert chain2_fits(1 << 30, 1, t2, t1)
    assert not chain2_fits(16384, 16384, t2, t1)                        # 16384 x 16384 along dims=1: exact fit of 16384
    assert chain2_fits(13440 * 3, 100, t2, t1)
    assert chain2_fits(1_000_000, 8, t2, t1)                            # < 12.5 % ragged either way
