"""The warp-ballot tie resolution of the percentile rank kernel (rank_sorted_pairs in
voxlogica_b200/csrc/vx_percentiles.cu) restated on Python integers: for 32 consecutive sorted
positions the start / end of every run of equal keys comes from one ballot of run heads / tails;
only runs that cross the warp's first / last position fall back to a binary search."""
import numpy as np
import pytest


def clz32(x):
    return 32 - int(x).bit_length()


def ffs32(x):
    return (int(x) & -int(x)).bit_length()


def ranks_by_warp_rules(keys):
    n = len(keys)
    less_out, up_out, fallbacks = np.zeros(n, np.int64), np.zeros(n, np.int64), 0
    for pb in range(0, n, 32):
        heads = tails = 0
        for lane in range(32):
            p = pb + lane
            if p >= n:
                continue
            if p == 0 or keys[p] != keys[p - 1]:
                heads |= 1 << lane
            if p + 1 >= n or keys[p + 1] != keys[p]:
                tails |= 1 << lane
        for lane in range(32):
            p = pb + lane
            if p >= n:
                continue
            below = heads & (0xffffffff >> (31 - lane))
            if below:
                less = pb + (31 - clz32(below))
            else:
                less = int(np.searchsorted(keys[:pb], keys[p], side="left"))
                fallbacks += 1
            above = tails & ((0xffffffff << lane) & 0xffffffff)
            if above:
                up = pb + (ffs32(above) - 1) + 1
            else:
                up = pb + 32 + int(np.searchsorted(keys[pb + 32:], keys[p], side="right"))
                fallbacks += 1
            less_out[p], up_out[p] = less, up
    return less_out, up_out, fallbacks


@pytest.mark.parametrize("n,distinct", [(1, 1), (31, 5), (32, 1), (33, 2), (1000, 7), (1000, 400), (257, 257)])
def test_run_bounds_from_ballots(n, distinct):
    rng = np.random.default_rng(n * 1000 + distinct)
    keys = np.sort(rng.integers(0, distinct, n)).astype(np.uint32)
    less, up, _ = ranks_by_warp_rules(keys)
    assert np.array_equal(less, np.searchsorted(keys, keys, side="left"))
    assert np.array_equal(up, np.searchsorted(keys, keys, side="right"))


def test_fallback_only_for_runs_crossing_a_warp():
    keys = np.repeat(np.arange(40, dtype=np.uint32), 8)     # runs of 8 never straddle a 32-position warp
    assert ranks_by_warp_rules(keys)[2] == 0
    keys = np.zeros(100, np.uint32)                         # one run over four warps: inner warps search both ways
    assert ranks_by_warp_rules(keys)[2] > 0


def test_rank_mode_comparison_on_integers():
    """code_compare_kernel's rank mode, restated.  The interpreter decides `table[code] OP c` of a deferred
    percentile image on float positions: x = inside the mask ? float(position of the code in the volume's code
    range, in value order) : -0.5, result = x >= T (GT, GE) or x < T (LT, LE) with the break point T = -1 or a
    position.  The fast kernel compares the code itself: for T >= 0  `position >= T`  is  `code >= lo + T`
    (increasing scaling) or `code <= hi - T` (decreasing), voxels outside the mask fail; for T = -1 everything
    passes.  Both forms must agree for every code, mask bit, direction, flip and break point."""
    rng = np.random.default_rng(17)
    for incr in (True, False):
        for flip in (0, 0x8000):
            raw = rng.integers(0, 65536, size=4000).astype(np.int64)         # the 16 bits as stored
            u = raw ^ flip                                                    # position order of the codes
            lo, hi = int(u.min()), int(u.max())
            m = rng.random(raw.size) < 0.6
            for T in (-1, 0, 1, 17, (hi - lo) // 2, hi - lo, hi - lo + 1, 65535):
                pos = (u - lo) if incr else (hi - u)
                x = np.where(m, pos.astype(np.float32), np.float32(-0.5))
                ge_float = x >= np.float32(T)
                if T < 0:
                    ge_int = np.ones(raw.size, bool)
                else:
                    ge_int = ((u >= lo + T) if incr else (u <= hi - T)) & m
                assert np.array_equal(ge_float, ge_int), (incr, flip, T)
                assert np.array_equal(x < np.float32(T), ~ge_int), (incr, flip, T)
