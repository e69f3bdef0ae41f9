"""The sliding-ball scheme of the crossCorrelation kernel (xcorr_ring_kernel in
voxlogica_b200/csrc/vx_xcorr.cu) restated in numpy: the ball as columns (dx, dz, half-extent ey)
along y; one histogram per output column of voxels that slides along y and only moves the voxels
entering and leaving each ball column; voxels outside the image or outside [m, M) sit in a dummy
bin k whose count is taken out of n1 and S11; the coefficient from exact integers
    r = (k*P - n1*n2) / sqrt((k*S11 - n1^2) * (k*S22 - n2^2))
with two int -> binary64 conversions, one product, one sqrt, one division, one narrowing.
Checked bit for bit against the oracle, which recounts every neighbourhood from scratch."""
import numpy as np
import pytest


def columns(sp, rho):
    r2 = float(rho) * float(rho)
    inside = lambda dx, dy, dz: (dx * sp[0]) ** 2 + (dy * sp[1]) ** 2 + (dz * sp[2]) ** 2 <= r2
    wx = wz = 0
    while inside(wx + 1, 0, 0): wx += 1
    while inside(0, 0, wz + 1): wz += 1
    cols = []
    for dz in range(-wz, wz + 1):
        for dx in range(-wx, wx + 1):
            if inside(dx, 0, dz):
                ey = 0
                while inside(dx, ey + 1, dz): ey += 1
                cols.append((dx, dz, ey))
    return cols


def bins_of(a, m, M, k):
    dv = a.astype(np.float64)
    t = ((dv - m) * float(k)) / (M - m) if M > m else np.zeros_like(dv)
    with np.errstate(invalid="ignore"):
        i = np.minimum(np.nan_to_num(t, nan=0.0, posinf=0.0, neginf=0.0).astype(np.int64), k - 1)
        return np.where((dv >= m) & (dv < M), i, k)           # k = the dummy bin: not counted


def sliding_xcorr(a, b, region, rho, m, M, k, sp):
    nz, ny, nx = a.shape
    cols = columns(sp, rho)
    ball = sum(2 * ey + 1 for _, _, ey in cols)
    bins = bins_of(a, m, M, k)
    hb = bins_of(b, m, M, k)[region != 0]
    h2 = np.bincount(hb[hb < k], minlength=k).astype(np.int64)
    n2, d2 = int(h2.sum()), int(k * (h2 * h2).sum() - h2.sum() ** 2)

    def bin_at(x, y, z):
        return bins[z, y, x] if (0 <= x < nx and 0 <= y < ny and 0 <= z < nz) else k

    out = np.zeros(a.shape, np.float32)
    for z in range(nz):
        for x in range(nx):
            h = np.zeros(k + 1, np.int64)
            for dx, dz, ey in cols:                              # full ball around (x, 0, z)
                for dy in range(-ey, ey + 1):
                    h[bin_at(x + dx, dy, z + dz)] += 1
            for y in range(ny):
                n1 = ball - int(h[k])
                P, S11 = int((h[:k] * h2).sum()), int((h[:k] * h[:k]).sum())
                num, d1 = k * P - n1 * n2, k * S11 - n1 * n1
                if d1 > 0 and d2 > 0:
                    out[z, y, x] = np.float32(np.float64(num) / np.sqrt(np.float64(d1) * np.float64(d2)))
                for dx, dz, ey in cols:                          # slide: one voxel enters, one leaves per column
                    h[bin_at(x + dx, y + 1 + ey, z + dz)] += 1
                    h[bin_at(x + dx, y - ey, z + dz)] -= 1
    return out


@pytest.mark.parametrize("sp,rho,k", [((1.0, 1.0, 1.0), 2.0, 10), ((1.0, 1.0, 1.0), 1.5, 7),
                                      ((0.5, 1.0, 2.0), 2.0, 16), ((1.0, 2.0, 1.0), 3.0, 5)])
def test_sliding_histograms_equal_recounting(oracle, sp, rho, k):
    rng = np.random.default_rng(int(rho * 10) + k)
    shape = (4, 9, 11)
    a = rng.random(shape).astype(np.float32) * 1.4 - 0.2          # some values outside [0, 1): dummy bin
    b = rng.random(shape).astype(np.float32)
    region = (rng.random(shape) < 0.4).astype(np.uint8)
    got = sliding_xcorr(a, b, region, rho, 0.0, 1.0, k, sp)
    want = oracle.cross_correlation(rho, a, b, region, 0.0, 1.0, k, sp)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_region_statistic_128bit():
    """k*S22 - n2^2 exceeds 63 bits once a bin holds > 4.3e8 voxels (a 1024^3 region, the 2048^3 slab
    path): the library's two-limb arithmetic + hand-written rounding (vx_xcorr_region_stat, the
    function xc_stat_kernel runs on the device) against the oracle's __int128 and Python integers."""
    import ctypes as C

    import oracle as orc
    from voxlogica_b200 import _abi

    lib = _abi.load_library()
    olib = orc.lib()
    olib.or_region_d2.restype = C.c_double
    rng = np.random.default_rng(11)
    cases = []
    for k in (1, 2, 7, 100, 254):
        cases.append((k, np.zeros(k, np.uint64)))
        cases.append((k, np.full(k, 0xffffffff, np.uint64)))                 # the largest legal histogram
        h = np.zeros(k, np.uint64); h[k // 2] = 600_000_000; cases.append((k, h))   # one huge bin
        h = np.zeros(k, np.uint64); h[0] = 0xffffffff; h[-1] = 1; cases.append((k, h))
        for _ in range(40):
            mag = int(rng.integers(1, 33))
            cases.append((k, rng.integers(0, 2 ** mag, size=k, dtype=np.uint64)))
        for _ in range(10):   # ties of the rounding: sparse histograms with huge counters
            h = np.zeros(k, np.uint64)
            h[rng.integers(0, k, size=min(k, 3))] = rng.integers(2 ** 31, 2 ** 32, size=min(k, 3), dtype=np.uint64)
            cases.append((k, h))
    for k, h in cases:
        hv = [int(x) for x in h]
        exact = k * sum(x * x for x in hv) - sum(hv) ** 2
        assert exact >= 0
        d2 = C.c_double(); pos = C.c_int32()
        assert lib.vx_xcorr_region_stat(h.ctypes.data_as(C.POINTER(C.c_uint64)), k, C.byref(d2), C.byref(pos)) == 0
        opos = C.c_int()
        od2 = olib.or_region_d2(h.ctypes.data_as(C.c_void_p), k, C.byref(opos))
        assert d2.value == float(exact) == od2, (k, hv[:4], d2.value, float(exact), od2)   # int -> float is RN-even
        assert bool(pos.value) == (exact > 0) == bool(opos.value)


def test_flat_columns_at_the_top_and_bottom_of_the_image():
    """The two border rules of xcorr_tma_kernel, restated.  Where everything a ball sees INSIDE the image is
    one code `ref` (rows above / below the image hold nothing: the not-counted code):
      * first ball of a column at row y0: a ball column with half-extent e that lies inside the image in x, z
        puts nin(e) = #rows of [y0-e, y0+e] inside [0, ny) into counter[ref] and the other 2e+1-nin(e) into
        the not-counted counter; a column outside the image in x or z puts all 2e+1 there;
      * slide y -> y+1 whose window is flat: d = sum over in-image columns of
        [leaving row y-e < 0 and entering row y+1+e < ny] - [entering row >= ny and leaving row >= 0]
        voxels move from the not-counted counter h0 to counter[ref] (h1), and the sum of squares over ALL
        counters changes by 2 d (h1 - h0) + 2 d^2.
    Checked against moving every voxel (the scheme above), for every (x, z) of small images -- columns at the
    x / z border included -- for images lower than the ball, and for a ref that is itself not counted."""
    rng = np.random.default_rng(3)
    for sp, rho in (((1.0, 1.0, 1.0), 2.0), ((1.0, 1.0, 1.0), 3.0), ((1.0, 1.5, 1.0), 3.0)):
        cols = columns(sp, rho)
        wy = max(e for _, _, e in cols)
        for (nz, ny, nx) in ((5, 9, 7), (3, 2 * wy - 1, 6), (4, 3, 5)):
            for ref in (0, 4):                                   # 0 plays the not-counted code here
                k = 6
                bins = np.full((nz, ny, nx), ref, np.int64)

                def bin_at(x, y, z):
                    return bins[z, y, x] if (0 <= x < nx and 0 <= y < ny and 0 <= z < nz) else 0

                for z in range(nz):
                    for x in range(nx):
                        inside = [(dx, dz, e) for dx, dz, e in cols if 0 <= x + dx < nx and 0 <= z + dz < nz]
                        y0 = int(rng.integers(0, ny))
                        # ---- first ball from row counts
                        h = np.zeros(k, np.int64)
                        for dx, dz, e in cols:
                            if (dx, dz, e) in inside:
                                nin = min(ny - 1, y0 + e) - max(0, y0 - e) + 1
                                h[ref] += nin
                                h[0] += 2 * e + 1 - nin
                            else:
                                h[0] += 2 * e + 1
                        want = np.zeros(k, np.int64)
                        for dx, dz, e in cols:
                            for dy in range(-e, e + 1):
                                want[bin_at(x + dx, y0 + dy, z + dz)] += 1
                        assert np.array_equal(h, want), (sp, rho, (nz, ny, nx), ref, x, z, y0)
                        # ---- slides down the rest of the column in closed form
                        S = int((h * h).sum())
                        for y in range(y0, ny - 1):
                            d = sum((1 if (y - e < 0 and y + 1 + e < ny) else 0) - (1 if (y + 1 + e >= ny and y - e >= 0) else 0)
                                    for _, _, e in inside)
                            if ref != 0 and d != 0:
                                h0, h1 = int(h[0]), int(h[ref])
                                h[0] -= d
                                h[ref] += d
                                S += 2 * d * (h1 - h0) + 2 * d * d
                            for dx, dz, e in cols:
                                want[bin_at(x + dx, y + 1 + e, z + dz)] += 1
                                want[bin_at(x + dx, y - e, z + dz)] -= 1
                            assert np.array_equal(h, want), (sp, rho, (nz, ny, nx), ref, x, z, y)
                            assert S == int((want * want).sum())


def test_block_order_visits_every_place_once():
    """The dispatch order of xcorr_tma_kernel's blocks, restated: units of 2 x chunks by 2 plane blocks (per y
    segment) sorted from the centre outwards, four place slots each, 0xffffffff where the unit sticks out of
    the grid of places; block i -> volume (i / 4) % nb, slot order[(i / (4 nb)) * 4 + i % 4].  Every (volume,
    place) must be visited exactly once, whatever the grid's parity."""
    for gx, nzb, nseg, nb in ((8, 26, 1, 16), (1, 1, 1, 1), (3, 5, 2, 2), (7, 1, 3, 1), (2, 2, 1, 5)):
        ux, uz = (gx + 1) // 2, (nzb + 1) // 2
        keyed = []
        for sg in range(nseg):
            for kz in range(uz):
                for kx in range(ux):
                    cx, cz = 0.5 * (2 * kx + min(2 * kx + 2, gx)), 0.5 * (2 * kz + min(2 * kz + 2, nzb))
                    fx, fz, fy = cx / gx - 0.5, cz / nzb - 0.5, (sg + 0.5) / nseg - 0.5
                    keyed.append((fx * fx + fz * fz + fy * fy, (sg * uz + kz) * ux + kx))
        keyed.sort(key=lambda t: t[0])                      # stable, like std::stable_sort
        order = []
        for _, u in keyed:
            kx, kz, sg = u % ux, u // ux % uz, u // ux // uz
            for j in range(4):
                bx, zb = 2 * kx + (j & 1), 2 * kz + (j >> 1)
                order.append(bx | ((zb * nseg + sg) << 16) if bx < gx and zb < nzb else 0xffffffff)
        assert len(order) == 4 * ux * uz * nseg
        seen = set()
        for i in range(len(order) * nb):
            unit, rem = divmod(i, 4 * nb)
            vol, place = rem >> 2, order[unit * 4 + (rem & 3)]
            if place == 0xffffffff:
                continue
            key = (vol, place & 0xffff, place >> 16)
            assert key not in seen
            seen.add(key)
        assert seen == {(v, bx, by) for v in range(nb) for bx in range(gx) for by in range(nzb * nseg)}
        # the first unit is the most central one, the last one touches the edge of the grid
        assert keyed[0][0] == min(k for k, _ in keyed) and keyed[-1][0] == max(k for k, _ in keyed)
