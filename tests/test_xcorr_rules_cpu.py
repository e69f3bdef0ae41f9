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
