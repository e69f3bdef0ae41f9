"""The tile ranking of the radix scatter (rs_scatter_kernel in voxlogica_b200/csrc/vx_percentiles.cu)
restated in numpy: warp-private running counts per digit, exclusive prefix over the warps inside a
digit, exclusive prefix over the digits of the tile, global base per digit and tile.  The result
must be the stable counting sort of the pass, tile after tile."""
import numpy as np
import pytest

TILE, WARPS, LANES, DIGITS = 4096, 8, 32, 512
WARP_KEYS = TILE // WARPS


def scatter_pass(keys, shift):
    n = len(keys)
    digit = (keys >> shift) & (DIGITS - 1)
    ntiles = (n + TILE - 1) // TILE
    # histogram kernel + row scan: table[d][tile] = keys of digit d in earlier tiles, totals[d]
    table = np.zeros((DIGITS, ntiles), np.int64)
    for t in range(ntiles):
        table[:, t] = np.bincount(digit[t * TILE:(t + 1) * TILE], minlength=DIGITS)
    totals = table.sum(axis=1)
    table = np.cumsum(table, axis=1) - table
    dig_base = np.cumsum(totals) - totals
    out = np.zeros_like(keys)
    for t in range(ntiles):
        tile = keys[t * TILE:(t + 1) * TILE]
        d_t = digit[t * TILE:(t + 1) * TILE]
        tile_n = len(tile)
        wcnt = np.zeros((WARPS, DIGITS), np.int64)
        lrank = np.zeros(tile_n, np.int64)
        for w in range(WARPS):                      # rounds of 32 keys: rank among the peers of the round + running count
            for r in range(WARP_KEYS // LANES):
                lo = w * WARP_KEYS + r * LANES
                for lane in range(LANES):
                    p = lo + lane
                    if p >= tile_n:
                        break
                    peers_before = int(np.sum(d_t[lo:p] == d_t[p]))
                    lrank[p] = wcnt[w, d_t[p]] + peers_before
                hi = min(lo + LANES, tile_n)
                if lo < hi:
                    np.add.at(wcnt[w], d_t[lo:hi], 1)
        woff = np.cumsum(wcnt, axis=0) - wcnt       # offsets of the warps inside a digit
        cnt = wcnt.sum(axis=0)
        dstart = np.cumsum(cnt) - cnt               # first tile-local position of a digit
        stage = np.zeros(tile_n, keys.dtype)
        for p in range(tile_n):
            stage[dstart[d_t[p]] + woff[p // WARP_KEYS, d_t[p]] + lrank[p]] = tile[p]
        sd = (stage >> shift) & (DIGITS - 1)
        for p in range(tile_n):                     # runs of one digit leave for consecutive addresses
            out[dig_base[sd[p]] + table[sd[p], t] + (p - dstart[sd[p]])] = stage[p]
    return out


@pytest.mark.parametrize("n", [1, 31, 4096, 4097, 9000])
def test_tile_ranking_is_a_stable_counting_sort(n):
    rng = np.random.default_rng(n)
    # 27-bit keys with a payload in the low bits of a 64-bit word, as the kernel sorts them
    key = rng.integers(0, 1 << 27, n).astype(np.uint64)
    key[rng.random(n) < 0.3] = key[0]               # ties: stability is observable through the payload
    pairs = (key << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    cur = pairs
    for p in range(3):
        cur = scatter_pass(cur, 32 + 9 * p)
        want = pairs[np.argsort((pairs >> np.uint64(32)) & np.uint64((1 << (9 * (p + 1))) - 1), kind="stable")]
        assert np.array_equal(cur, want), p
    assert np.array_equal(cur, pairs[np.argsort(key, kind="stable")])
