"""The CPU oracle against independent implementations (scipy.ndimage, OpenCV) and
against the algebraic laws stated in the slides.  No GPU.  These cross-checks guard the
oracle against self-consistent mistakes; they are NOT SimpleITK (parity unpinned)."""
import numpy as np
import pytest
from scipy import ndimage as ndi

S6 = ndi.generate_binary_structure(3, 1)
S26 = ndi.generate_binary_structure(3, 3)


def rnd_mask(shape, p, seed):
    return (np.random.default_rng(seed).random(shape) < p).astype(np.uint8)


def blobs(shape, seed, thr=0.55):
    r = np.random.default_rng(seed).random(shape).astype(np.float32)
    return (ndi.gaussian_filter(r, 2.0) > np.quantile(ndi.gaussian_filter(r, 2.0), thr)).astype(np.uint8)


@pytest.mark.parametrize("conn,st", [(6, S6), (26, S26)])
def test_near_interior_vs_scipy(oracle, conn, st):
    a = rnd_mask((9, 11, 13), 0.2, 1)
    assert np.array_equal(oracle.near(a, conn), ndi.binary_dilation(a, st).astype(np.uint8))
    # erosion that does not erode from outside the image: border_value=1
    assert np.array_equal(oracle.interior(a | 1 - rnd_mask(a.shape, 0.9, 2), conn),
                          ndi.binary_erosion(a | 1 - rnd_mask(a.shape, 0.9, 2), st, border_value=1).astype(np.uint8))


def test_near_2d_vs_opencv(oracle):
    import cv2
    a = rnd_mask((40, 37), 0.15, 3)
    assert np.array_equal(oracle.near(a, 26), cv2.dilate(a, np.ones((3, 3), np.uint8)))
    cross = cv2.getStructuringElement(cv2.MORPH_CROSS, (3, 3))
    assert np.array_equal(oracle.near(a, 6), cv2.dilate(a, cross))


def test_closure_laws(oracle):
    """Greta.pdf p.24-25: closure is extensive and monotone, preserves union, and is NOT
    idempotent on grids; p.35: I = !N!."""
    a, b = rnd_mask((6, 8, 10), 0.1, 4), rnd_mask((6, 8, 10), 0.1, 5)
    for conn in (6, 26):
        na = oracle.near(a, conn)
        assert np.all(na >= a)
        assert np.array_equal(oracle.near(a | b, conn), na | oracle.near(b, conn))
        assert not np.array_equal(oracle.near(na, conn), na)
        assert np.array_equal(oracle.interior(a, conn), 1 - oracle.near(1 - a, conn))


def canon(lab):
    """relabel any labelling to 1 + first raster index of each component"""
    out = np.zeros(lab.shape, np.uint32)
    flat = lab.ravel()
    first = {}
    for i, l in enumerate(flat):
        if l and l not in first:
            first[l] = i + 1
    o = out.ravel()
    for i, l in enumerate(flat):
        if l:
            o[i] = first[l]
    return out


@pytest.mark.parametrize("conn,st", [(6, S6), (26, S26)])
@pytest.mark.parametrize("p", [0.1, 0.35, 0.7])
def test_components_vs_scipy(oracle, conn, st, p):
    a = rnd_mask((7, 12, 15), p, 6)
    lab, _ = ndi.label(a, st)
    assert np.array_equal(oracle.components(a, conn), canon(lab))


def test_components_2d_vs_opencv(oracle):
    import cv2
    a = blobs((64, 70), 7)
    for conn, c in ((6, 4), (26, 8)):
        _, lab = cv2.connectedComponents(a, connectivity=c)
        assert np.array_equal(oracle.components(a, conn), canon(lab))


@pytest.mark.parametrize("conn,st", [(6, S6), (26, S26)])
def test_through_and_maxvol_vs_scipy(oracle, conn, st):
    b = blobs((10, 24, 24), 8)
    a = rnd_mask(b.shape, 0.002, 9)
    lab, n = ndi.label(b, st)
    hit = np.unique(lab[(a > 0) & (b > 0)])
    want = np.isin(lab, hit[hit > 0]).astype(np.uint8)
    assert np.array_equal(oracle.through(a, b, conn), want)
    sizes = ndi.sum(b, lab, index=np.arange(1, n + 1))
    wantmv = np.isin(lab, 1 + np.flatnonzero(sizes == sizes.max())).astype(np.uint8)
    assert np.array_equal(oracle.maxvol(b, conn), wantmv)


def test_reach_duality(oracle):
    """Greta.pdf p.36: phi1 R phi2 = !((!phi2) S (!phi1)) is a law of the logic; here we
    check the weaker, implementation-level facts through() must satisfy."""
    b = blobs((6, 20, 20), 10)
    a = rnd_mask(b.shape, 0.01, 11)
    t = oracle.through(a, b)
    assert np.all(t <= b) and np.all((a & b) <= t)
    assert np.array_equal(oracle.through(t, b), t)            # idempotent
    assert np.array_equal(oracle.through(a, b), oracle.through(a & b, b))


@pytest.mark.parametrize("spacing", [(1, 1, 1), (0.5, 1.25, 3.0)])
@pytest.mark.parametrize("p", [0.001, 0.05, 0.5])
def test_edt_vs_scipy(oracle, spacing, p):
    a = rnd_mask((9, 17, 21), p, 12)
    a[4, 8, 10] = 1
    got = oracle.edt(a, spacing)
    want = ndi.distance_transform_edt(1 - a, sampling=spacing[::-1])
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=0)
    if spacing == (1, 1, 1):  # squared distances are exact integers
        assert np.array_equal(np.rint(got.astype(np.float64) ** 2), np.rint(want ** 2))


def test_edt_empty_and_full(oracle):
    z = np.zeros((3, 4, 5), np.uint8)
    assert np.all(np.isinf(oracle.edt(z)))
    assert np.all(oracle.edt(1 - z) == 0)
    assert np.all(oracle.distlt(3.0, z) == 0) and np.all(oracle.distgeq(3.0, z) == 1)


def test_dist_ops_and_filter(oracle):
    a = blobs((8, 30, 30), 13)
    d = ndi.distance_transform_edt(1 - a)
    assert np.array_equal(oracle.distlt(2.0, a), (d < 2.0).astype(np.uint8))
    assert np.array_equal(oracle.distgeq(5.0, a), (d >= 5.0).astype(np.uint8))
    assert np.array_equal(oracle.distleq(np.sqrt(5.0), a), (np.rint(d ** 2) <= 5).astype(np.uint8))
    f = oracle.flt(2.0, a)
    assert np.all(f <= a)                               # opening is anti-extensive
    assert np.array_equal(oracle.flt(2.0, f), f)        # and idempotent


def test_edt_signed(oracle):
    a = blobs((6, 20, 20), 14)
    s = oracle.edt_signed(a)
    assert np.all(s[a == 0] > 0) and np.all(s[a == 1] <= 0)
    bnd = a & ndi.binary_dilation(1 - a, S26)
    assert np.all(s[bnd == 1] == 0)
    inside = ndi.distance_transform_edt(1 - bnd)
    np.testing.assert_allclose(-s[a == 1], inside[a == 1], rtol=1e-5)


def test_percentiles(oracle):
    rng = np.random.default_rng(15)
    v = np.round(rng.random((5, 9, 11)).astype(np.float32), 1)  # many ties
    m = rnd_mask(v.shape, 0.6, 16)
    vals = v[m == 1]
    for corr in (0.0, 0.5, 1.0):
        got = oracle.percentiles(v, m, corr)
        want = np.zeros_like(v)
        for idx in np.argwhere(m == 1):
            x = v[tuple(idx)]
            want[tuple(idx)] = np.float32((np.sum(vals < x) + corr * np.sum(vals == x)) / vals.size)
        assert np.array_equal(got, want)
    assert np.all(oracle.percentiles(v, np.zeros_like(m))== 0)


def test_cross_correlation_forms_agree(oracle):
    rng = np.random.default_rng(17)
    a = rng.random((7, 12, 13)).astype(np.float32)
    b = rng.random(a.shape).astype(np.float32)
    reg = rnd_mask(a.shape, 0.3, 18)
    exact = oracle.cross_correlation(2.0, a, b, reg, 0.0, 1.0, 10)
    text = oracle.cross_correlation(2.0, a, b, reg, 0.0, 1.0, 10, textbook=True)
    np.testing.assert_allclose(exact, text, rtol=1e-5, atol=1e-6)
    assert np.all(np.abs(exact) <= 1.0 + 1e-6)
    # brute-force numpy statement of the definition at a few voxels
    k, m, M, rho = 10, 0.0, 1.0, 2.0
    h2 = np.histogram(b[reg == 1], bins=k, range=(m, M))[0].astype(float)
    zz, yy, xx = np.indices(a.shape)
    for (z, y, x) in [(0, 0, 0), (3, 6, 6), (6, 11, 12), (2, 0, 7)]:
        ball = (zz - z) ** 2 + (yy - y) ** 2 + (xx - x) ** 2 <= rho * rho
        vals = a[ball]
        vals = vals[(vals >= m) & (vals < M)]
        h1 = np.histogram(vals, bins=k, range=(m, M))[0].astype(float)
        want = np.corrcoef(h1, h2)[0, 1]
        assert abs(exact[z, y, x] - want) < 1e-5
    # a region's own texture correlates perfectly with itself when the ball covers the image
    small = rng.random((3, 3, 3)).astype(np.float32)
    full = oracle.cross_correlation(10.0, small, small, np.ones_like(small, np.uint8), 0.0, 1.0, 5)
    np.testing.assert_allclose(full, 1.0, rtol=1e-6)


def test_reductions_and_pointwise(oracle):
    rng = np.random.default_rng(19)
    a = rng.standard_normal((4, 5, 6)).astype(np.float32)
    m = rnd_mask(a.shape, 0.5, 20)
    assert oracle.volume(m) == m.sum()
    assert oracle.min_(a) == a.min() and oracle.max_(a) == a.max()
    assert oracle.min_(a, m) == a[m == 1].min() and oracle.max_(a, m) == a[m == 1].max()
    assert np.array_equal(oracle.cmp_scalar(a, ">", 0.1), (a > np.float32(0.1)).astype(np.uint8))
    assert np.array_equal(oracle.arith_scalar(a, "r/", 2.0), (np.float32(2.0) / a))
    assert np.array_equal(oracle.border(np.zeros((3, 4, 5)))[1, 1:-1, 1:-1], np.zeros((2, 3), np.uint8))
    assert oracle.border(np.zeros((1, 4, 5)))[0, 1:-1, 1:-1].sum() == 0      # 2D: z is no face
    assert oracle.border(np.zeros((1, 4, 5))).sum() == 4 * 5 - 2 * 3


def test_derived_touch_grow(oracle):
    b = blobs((6, 24, 24), 21)
    bord = oracle.border(b)
    t = oracle.touch(b, bord)
    lab, n = ndi.label(b, S26)
    hit = np.unique(lab[(ndi.binary_dilation(bord, S26)) & (b > 0)])
    assert np.array_equal(t, np.isin(lab, hit[hit > 0]).astype(np.uint8))
    g = oracle.grow(oracle.flt(2.0, b), b)
    assert np.all(g <= b) and np.all(oracle.flt(2.0, b) <= g)
