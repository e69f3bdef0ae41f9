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
