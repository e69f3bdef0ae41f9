"""Parity of the CUDA operators (through the C ABI) against the CPU oracle.

Bar (task section 3): bit-exact for boolean, label and index outputs; distances and
correlations are ALSO compared bit-for-bit here because oracle and kernels share one
arithmetic contract (see vx_edt.cu / vx_xcorr.cu headers); the 1e-5 relative tolerance
of north_star is checked against scipy as an independent witness.

Every test runs through libvoxcuda.so; the `ctx` fixture fails (never skips) when the
extension or the device is missing.
"""
import numpy as np
import pytest
from scipy import ndimage as ndi

pytestmark = pytest.mark.gpu

# (nb, nz, ny, nx): aligned rows, ragged rows, 2D, single row/column, BraTS-like x extent
SHAPES = [(2, 9, 20, 48), (1, 7, 13, 37), (3, 1, 33, 70), (1, 5, 1, 19), (1, 1, 1, 1), (2, 6, 10, 240),
          (1, 20, 18, 32)]


def rnd_mask(shape, p, seed):
    return (np.random.default_rng(seed).random(shape) < p).astype(np.uint8)


def blobs(shape, seed, sigma=2.0, q=0.55):
    r = np.random.default_rng(seed).random(shape).astype(np.float32)
    g = np.stack([ndi.gaussian_filter(v, sigma) for v in r])
    return (g > np.quantile(g, q)).astype(np.uint8)


def per_volume(fn, *arrs, **kw):
    return np.stack([fn(*[a[i] for a in arrs], **kw) for i in range(arrs[0].shape[0])])


# ------------------------------------------------------------------ pointwise / fused
@pytest.mark.parametrize("shape", SHAPES)
def test_boolean_ops(ctx, oracle, shape):
    a, b = rnd_mask(shape, 0.4, 1), rnd_mask(shape, 0.6, 2)
    A, B = ctx.upload(a), ctx.upload(b)
    assert np.array_equal(ctx.not_(A).numpy(), oracle.not_(a))
    assert np.array_equal(ctx.and_(A, B).numpy(), oracle.and_(a, b))
    assert np.array_equal(ctx.or_(A, B).numpy(), oracle.or_(a, b))
    assert np.array_equal(ctx.xor(A, B).numpy(), oracle.xor(a, b))
    assert np.all(ctx.tt(A).numpy() == 1) and np.all(ctx.ff(A).numpy() == 0)
    assert ctx.tt(A).numpy().shape == shape


@pytest.mark.parametrize("shape", SHAPES)
def test_numeric_ops(ctx, oracle, shape):
    rng = np.random.default_rng(3)
    a = rng.standard_normal(shape).astype(np.float32)
    b = rng.standard_normal(shape).astype(np.float32)
    a.ravel()[0] = 0.5  # an exact tie with the threshold
    A, B = ctx.upload(a), ctx.upload(b)
    for op in ("<", "<=", ">", ">=", "==", "!="):
        assert np.array_equal(ctx.cmp_scalar(A, op, 0.5).numpy(), oracle.cmp_scalar(a, op, 0.5)), op
        assert np.array_equal(ctx.cmp(A, op, B).numpy(), oracle.cmp(a, op, b)), op
    for op in ("+", "-", "*", "/", "r-", "r/", "min", "max"):
        got, want = ctx.arith_scalar(A, op, 1.7).numpy(), oracle.arith_scalar(a, op, 1.7)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), op
        got, want = ctx.arith(A, op, B).numpy(), oracle.arith(a, op, b)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), op
    m = rnd_mask(shape, 0.5, 4)
    assert np.array_equal(ctx.mask(A, ctx.upload(m), -1.0).numpy(), oracle.mask(a, m, -1.0))
    assert np.array_equal(ctx.to_f32(ctx.upload(m)).numpy(), oracle.to_f32(m))
    assert np.array_equal(ctx.to_bool(A).numpy(), oracle.to_bool(a))
    assert np.all(ctx.constant(A, 2.5).numpy() == np.float32(2.5))


def test_per_volume_scalars(ctx, oracle):
    shape = (3, 5, 7, 11)  # 385 voxels per volume: 16-voxel groups straddle volumes
    a = np.random.default_rng(5).random(shape).astype(np.float32)
    s = np.array([0.2, 0.5, 0.8])
    A = ctx.upload(a)
    want = np.stack([oracle.cmp_scalar(a[i], ">", s[i]) for i in range(3)])
    assert np.array_equal(ctx.cmp_scalar(A, ">", s).numpy(), want)
    want = np.stack([oracle.arith_scalar(a[i], "/", s[i]) for i in range(3)])
    assert np.array_equal(ctx.arith_scalar(A, "/", s).numpy(), want)


def test_fused_program(ctx, oracle):
    from voxlogica_b200 import _abi as A
    shape = (2, 6, 9, 40)
    rng = np.random.default_rng(6)
    f = rng.random(shape).astype(np.float32)
    g = rng.random(shape).astype(np.float32)
    m = rnd_mask(shape, 0.5, 7)
    # out = ((f > 0.3) & !(g <= f)) | (m ^ (f*2 - g >= 0.4))
    prog = [
        (A.VXP_LOAD, 0, 0, 0, 0, 0.0), (A.VXP_LOAD, 1, 1, 0, 0, 0.0), (A.VXP_LOAD, 2, 2, 0, 0, 0.0),
        (A.VXP_CMPS, 3, 0, 0, A.VX_GT, 0.3),
        (A.VXP_CMP, 4, 1, 0, A.VX_LE, 0.0),
        (A.VXP_NOT, 4, 4, 0, 0, 0.0),
        (A.VXP_AND, 3, 3, 4, 0, 0.0),
        (A.VXP_ARITHS, 5, 0, 0, A.VX_MUL, 2.0),
        (A.VXP_ARITH, 5, 5, 1, A.VX_SUB, 0.0),
        (A.VXP_CMPS, 6, 5, 0, A.VX_GE, 0.4),
        (A.VXP_XOR, 6, 2, 6, 0, 0.0),
        (A.VXP_OR, 7, 3, 6, 0, 0.0),
    ]
    got = ctx.fused(prog, [ctx.upload(f), ctx.upload(g), ctx.upload(m)]).numpy()
    t = oracle.arith(oracle.arith_scalar(f, "*", 2.0), "-", g)
    want = oracle.or_(oracle.and_(oracle.cmp_scalar(f, ">", 0.3), oracle.not_(oracle.cmp(g, "<=", f))),
                      oracle.xor(m, oracle.cmp_scalar(t, ">=", 0.4)))
    assert np.array_equal(got, want)
    # numeric root + select
    prog2 = [(A.VXP_LOAD, 0, 0, 0, 0, 0.0), (A.VXP_LOAD, 1, 2, 0, 0, 0.0),
             (A.VXP_ARITHS, 2, 0, 0, A.VX_ADD, 1.0), (A.VXP_SELECT, 3, 1, 2, 0, -7.0)]
    got2 = ctx.fused(prog2, [ctx.upload(f), ctx.upload(g), ctx.upload(m)]).numpy()
    assert np.array_equal(got2, oracle.mask(oracle.arith_scalar(f, "+", 1.0), m, -7.0))


def test_error_convention(ctx):
    from voxlogica_b200.ops import VoxError
    a = ctx.upload(np.zeros((2, 3, 4), np.uint8))
    b = ctx.upload(np.zeros((2, 3, 5), np.uint8))
    f = ctx.upload(np.zeros((2, 3, 4), np.float32))
    with pytest.raises(VoxError, match="SHAPE"):
        ctx.and_(a, b)
    with pytest.raises(VoxError, match="DTYPE"):
        ctx.near(f)
    with pytest.raises(VoxError, match="DTYPE"):
        ctx.and_(a, f)
    with pytest.raises(VoxError, match="INVALID_ARG"):
        ctx.near(a, conn=5)
    a.release()
    with pytest.raises(VoxError, match="INVALID_ARG"):
        ctx.not_(a)


# ------------------------------------------------------------------ border / near / interior
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("conn", [6, 26])
def test_near_interior(ctx, oracle, shape, conn):
    for p, seed in ((0.03, 8), (0.5, 9), (0.97, 10)):
        a = rnd_mask(shape, p, seed)
        A = ctx.upload(a)
        assert np.array_equal(ctx.near(A, conn).numpy(), per_volume(oracle.near, a, conn=conn)), (p, "near")
        assert np.array_equal(ctx.interior(A, conn).numpy(), per_volume(oracle.interior, a, conn=conn)), (p, "int")


def test_near_wide_rows(ctx, oracle):
    """rows wider than one 16-chunk strip (256 voxels) exercise the cross-strip edge bytes"""
    a = rnd_mask((1, 3, 5, 1040), 0.01, 11)
    a[0, 1, 2, 255] = 1; a[0, 1, 2, 256] = 0; a[0, 2, 3, 512] = 1; a[0, 0, 0, 1039] = 1
    A = ctx.upload(a)
    for conn in (6, 26):
        assert np.array_equal(ctx.near(A, conn).numpy(), per_volume(oracle.near, a, conn=conn))
        assert np.array_equal(ctx.interior(ctx.not_(A), conn).numpy(), per_volume(oracle.interior, 1 - a, conn=conn))


@pytest.mark.parametrize("shape", SHAPES)
def test_border(ctx, oracle, shape):
    A = ctx.upload(np.zeros(shape, np.uint8))
    assert np.array_equal(ctx.border(A).numpy(), per_volume(oracle.border, np.zeros(shape, np.uint8)))


# ------------------------------------------------------------------ connected components
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("conn", [6, 26])
def test_components_and_through(ctx, oracle, shape, conn):
    for p, seed in ((0.05, 12), (0.25, 13), (0.45, 14), (0.8, 15)):
        b = rnd_mask(shape, p, seed)
        a = rnd_mask(shape, 0.02, seed + 100)
        B, Aimg = ctx.upload(b), ctx.upload(a)
        assert np.array_equal(ctx.components(B, conn).numpy(), per_volume(oracle.components, b, conn=conn)), p
        assert np.array_equal(ctx.through(Aimg, B, conn).numpy(), per_volume(oracle.through, a, b, conn=conn)), p
        assert np.array_equal(ctx.maxvol(B, conn).numpy(), per_volume(oracle.maxvol, b, conn=conn)), p


@pytest.mark.parametrize("conn", [6, 26])
def test_components_structured(ctx, oracle, conn):
    """blobs, long snakes and diagonal staircases: many cross-row merges per component"""
    shape = (2, 12, 40, 96)
    b = blobs(shape, 16)
    z, y, x = np.indices(shape[1:])
    b[1] = (((x + y + z) % 7 == 0) | ((x - y) % 11 == 0)).astype(np.uint8)
    a = rnd_mask(shape, 0.001, 17)
    B, Aimg = ctx.upload(b), ctx.upload(a)
    assert np.array_equal(ctx.components(B, conn).numpy(), per_volume(oracle.components, b, conn=conn))
    assert np.array_equal(ctx.through(Aimg, B, conn).numpy(), per_volume(oracle.through, a, b, conn=conn))
    assert np.array_equal(ctx.maxvol(B, conn).numpy(), per_volume(oracle.maxvol, b, conn=conn))


def test_components_spiral_2d(ctx, oracle):
    """one long spiral corridor: worst case for label propagation (Greta.pdf p.78 needed
    ~29 sweeps on a small scene); union-find must resolve it in one pass"""
    n = 129
    b = np.zeros((n, n), np.uint8)
    x0, y0, x1, y1 = 0, 0, n - 1, n - 1
    while x0 <= x1 and y0 <= y1:
        b[y0, x0:x1 + 1] = 1; b[y0:y1 + 1, x1] = 1; b[y1, x0 + 2:x1 + 1] = 1; b[y0 + 2:y1 + 1, x0 + 2] = 1
        x0 += 4; y0 += 4; x1 -= 4; y1 -= 4
    for conn in (6, 26):
        got = ctx.components(ctx.upload(b), conn).numpy()[0, 0]
        assert np.array_equal(got, oracle.components(b, conn))


# ------------------------------------------------------------------ distances
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("spacing", [(1.0, 1.0, 1.0), (0.5, 1.25, 3.0)])
def test_edt(ctx, oracle, shape, spacing):
    for p, seed in ((0.002, 18), (0.05, 19), (0.6, 20)):
        a = rnd_mask(shape, p, seed)
        a[0].ravel()[a[0].size // 2] = 1
        got = ctx.edt(ctx.upload(a, spacing=spacing)).numpy()
        want = per_volume(oracle.edt, a, spacing=spacing)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), p
        sci = per_volume(lambda v: ndi.distance_transform_edt(1 - v, sampling=spacing[::-1]) if v.any()
                         else np.full(v.shape, np.inf), a)
        np.testing.assert_allclose(got, sci, rtol=1e-5)


def test_edt_empty_volume_in_batch(ctx, oracle):
    a = rnd_mask((3, 4, 9, 21), 0.05, 21)
    a[1] = 0
    got = ctx.edt(ctx.upload(a)).numpy()
    assert np.all(np.isinf(got[1]))
    assert np.array_equal(got, per_volume(oracle.edt, a))
    assert np.all(ctx.distlt(2.0, ctx.upload(a)).numpy()[1] == 0)
    assert np.all(ctx.distgeq(2.0, ctx.upload(a)).numpy()[1] == 1)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("spacing", [(1.0, 1.0, 1.0), (0.7, 1.0, 2.5)])
def test_dist_thresholds(ctx, oracle, shape, spacing):
    a = rnd_mask(shape, 0.01, 22)
    A = ctx.upload(a, spacing=spacing)
    # 5.0 and 2.0 are the GTV radii; sqrt(5), 3.0 hit exact ties; 40 forces the wide-radius path
    for r in (2.0, 5.0, float(np.sqrt(np.float32(5.0))), 3.0, 0.0, -1.0, 40.0):
        for name, op in (("distlt", "<"), ("distleq", "<="), ("distgeq", ">="), ("distgt", ">")):
            got = getattr(ctx, name)(r, A).numpy()
            want = per_volume(oracle.dist_cmp, a, op=op, r=r, spacing=spacing)
            assert np.array_equal(got, want), (name, r)


def test_filter_is_opening(ctx, oracle):
    a = blobs((2, 10, 30, 48), 23)
    A = ctx.upload(a)
    f = ctx.distlt(2.0, ctx.distgeq(2.0, ctx.not_(A)))
    want = np.stack([oracle.flt(2.0, v) for v in a])
    assert np.array_equal(f.numpy(), want)
    again = ctx.distlt(2.0, ctx.distgeq(2.0, ctx.not_(f)))
    assert np.array_equal(again.numpy(), f.numpy())


def test_edt_signed(ctx, oracle):
    a = blobs((2, 8, 20, 32), 24)
    got = ctx.edt_signed(ctx.upload(a)).numpy()
    want = per_volume(oracle.edt_signed, a)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ------------------------------------------------------------------ reductions / percentiles
@pytest.mark.parametrize("shape", SHAPES)
def test_reductions(ctx, oracle, shape):
    rng = np.random.default_rng(25)
    a = rng.standard_normal(shape).astype(np.float32)
    m = rnd_mask(shape, 0.3, 26)
    A, M = ctx.upload(a), ctx.upload(m)
    nb = shape[0]
    assert np.array_equal(ctx.volume(M), [oracle.volume(m[i]) for i in range(nb)])
    assert np.array_equal(ctx.min(A), [oracle.min_(a[i]) for i in range(nb)])
    assert np.array_equal(ctx.max(A), [oracle.max_(a[i]) for i in range(nb)])
    assert np.array_equal(ctx.min(A, M), [oracle.min_(a[i], m[i]) for i in range(nb)])
    assert np.array_equal(ctx.max(A, M), [oracle.max_(a[i], m[i]) for i in range(nb)])
    np.testing.assert_allclose(ctx.sum(A), [oracle.sum_(a[i]) for i in range(nb)], rtol=1e-9, atol=1e-9)
    # one-pass min+max: the same values as the separate reductions, masked or not, NaNs skipped
    lo, hi = ctx.minmax(A)
    assert np.array_equal(lo, ctx.min(A)) and np.array_equal(hi, ctx.max(A))
    lo, hi = ctx.minmax(A, M)
    assert np.array_equal(lo, ctx.min(A, M)) and np.array_equal(hi, ctx.max(A, M))
    b = a.copy()
    b.reshape(nb, -1)[:, ::7] = np.nan
    lo, hi = ctx.minmax(ctx.upload(b))
    for i in range(nb):      # a volume of NaNs only has no minimum: +inf / -inf, like an empty mask
        some = not np.all(np.isnan(b[i]))
        assert lo[i] == (np.nanmin(b[i]) if some else np.inf) and hi[i] == (np.nanmax(b[i]) if some else -np.inf)


@pytest.mark.parametrize("shape", SHAPES + [(2, 12, 40, 50)])
def test_percentiles(ctx, oracle, shape):
    rng = np.random.default_rng(27)
    a = rng.random(shape).astype(np.float32)
    a[rng.random(shape) < 0.3] = 0.25        # long runs of equal keys
    a[rng.random(shape) < 0.05] *= -1.0      # negative keys
    a.ravel()[:2] = [0.0, -0.0] if a.size > 1 else [0.0]
    m = rnd_mask(shape, 0.7, 28)
    A, M = ctx.upload(a), ctx.upload(m)
    for corr in (0.0, 0.5):
        got = ctx.percentiles(A, M, corr).numpy()
        want = per_volume(oracle.percentiles, a, m, correction=corr)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), corr
    empty = ctx.percentiles(A, ctx.ff(M)).numpy()
    assert np.all(empty == 0)


@pytest.mark.parametrize("lo,hi,note", [(0.5, 0.5, "all keys equal: 0 sort passes"),
                                        (1.0, 1.0000305, "8 significant bits: 1 pass"),
                                        (1.0, 1.01, "17 bits: 2 passes"),
                                        (0.1, 1.0, "26 bits: 3 passes (the GTV case)"),
                                        (-3.0, 1.0e6, "mixed signs, many octaves: 4 passes")])
def test_percentiles_key_ranges(ctx, oracle, lo, hi, note):
    """the radix sort only sorts the bits in which the keys differ (device-side pass count)"""
    rng = np.random.default_rng(33)
    shape = (2, 10, 30, 48)
    a = (lo + (hi - lo) * rng.random(shape)).astype(np.float32)
    a[0, 0, 0, :2] = [lo, hi]
    m = rnd_mask(shape, 0.8, 34)
    m[0, 0, 0, :2] = 1
    got = ctx.percentiles(ctx.upload(a), ctx.upload(m), 0.0).numpy()
    want = per_volume(oracle.percentiles, a, m, correction=0.0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), note


@pytest.mark.parametrize("dtype,slope,inter,lo,hi,note", [
    (np.int16, 1.0 / 4095.0, 0.0, 0, 4096, "12-bit scanner data: the shared-memory window"),
    (np.int16, 0.25, -3.0, -2000, 2000, "negative codes"),
    (np.uint16, -2.5, 1e3, 0, 900, "decreasing scaling: value order = reversed code order"),
    (np.uint16, 1.0, 0.0, 0, 65536, "the whole code range: global histogram and table"),
    (np.int16, 1.0, 0.0, -32768, 32768, "whole signed range"),
    (np.int16, 3.0, 0.5, 7, 8, "a single code"),
])
def test_percentiles_of_quantised_images(ctx, oracle, dtype, slope, inter, lo, hi, note):
    """images uploaded as 16-bit codes (vx_upload_scaled) are ranked by counting instead of sorting:
    bit-identical to the oracle and to the sorting path (VX_OPT_QUANTISED_PATHS = 0), ties at every level"""
    from voxlogica_b200 import _abi as A
    rng = np.random.default_rng(71)
    for shape in [(2, 9, 30, 48), (3, 5, 7, 13), (1, 1, 1, 1)]:
        raw = rng.integers(lo, hi, size=shape).astype(dtype)
        raw[rng.random(shape) < 0.2] = dtype(lo)                 # one heavy tie
        m = rnd_mask(shape, 0.7, 72)
        img, M = ctx.upload_scaled(raw, slope, inter), ctx.upload(m)
        f = raw.astype(np.float32) * np.float32(slope) + np.float32(inter)
        assert np.array_equal(img.numpy().view(np.uint32), f.view(np.uint32))
        for corr in (0.0, 0.5, 1.0):
            want = per_volume(oracle.percentiles, f, m, correction=corr)
            got = ctx.percentiles(img, M, corr).numpy()
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (note, shape, corr)
            ctx.set_option(A.VX_OPT_QUANTISED_PATHS, 0)
            try:
                sort = ctx.percentiles(img, M, corr).numpy()
            finally:
                ctx.set_option(A.VX_OPT_QUANTISED_PATHS, 1)
            assert np.array_equal(sort.view(np.uint32), want.view(np.uint32)), (note, shape, corr)
        assert np.all(ctx.percentiles(img, ctx.ff(M)).numpy() == 0)          # empty mask
        assert np.array_equal(ctx.percentiles(img, ctx.tt(M)).numpy().view(np.uint32),
                              per_volume(oracle.percentiles, f, np.ones(shape, np.uint8), correction=0.0).view(np.uint32))


@pytest.mark.parametrize("dtype", [np.int16, np.uint16])
def test_pointwise_programs_read_quantised_leaves_as_codes(ctx, dtype):
    """a pointwise program takes a 16-bit upload by its codes (2 bytes per voxel) and forms the same f32
    values as the upload's conversion: comparisons, arithmetic and mixed programs, switch on and off"""
    from voxlogica_b200 import _abi as A
    rng = np.random.default_rng(95)
    info = np.iinfo(dtype)
    for shape in [(2, 5, 7, 13), (1, 3, 16, 24), (3, 1, 1, 3), (1, 2, 9, 8)]:
        raw = rng.integers(info.min, info.max + 1, size=shape).astype(dtype)
        other = rng.standard_normal(shape).astype(np.float32)
        for slope, inter in [(1.0 / 4095.0, 0.0), (0.0007324442, -3.25), (-2.5, 1e-3)]:
            f = raw.astype(np.float32) * np.float32(slope) + np.float32(inter)
            img, Oth = ctx.upload_scaled(raw, slope, inter), ctx.upload(other)
            for opt in (1, 0):
                ctx.set_option(A.VX_OPT_QUANTISED_PATHS, opt)
                try:
                    thr = np.float32(np.median(f))
                    assert np.array_equal(ctx.cmp_scalar(img, "<", float(thr)).numpy(), (f < thr).astype(np.uint8))
                    got = ctx.arith_scalar(img, "*", 1.7).numpy()
                    assert np.array_equal(got.view(np.uint32), (f * np.float32(1.7)).view(np.uint32))
                    got = ctx.arith(img, "-", Oth).numpy()
                    assert np.array_equal(got.view(np.uint32), (f - other).view(np.uint32))
                    assert np.array_equal(ctx.cmp(Oth, ">=", img).numpy(), (other >= f).astype(np.uint8))
                finally:
                    ctx.set_option(A.VX_OPT_QUANTISED_PATHS, 1)


def test_deferred_percentile_image(ctx, oracle):
    """percentiles of a 16-bit upload returns an image whose f32 array is made on demand: pointwise
    programs read codes + mask + table (before and after the array exists), every other consumer gets
    the array; the operands may be released first; consumers on several threads at once"""
    import threading
    from voxlogica_b200 import _abi as A
    rng = np.random.default_rng(91)
    for shape, dtype in [((3, 5, 7, 13), np.int16), ((2, 8, 16, 32), np.uint16), ((1, 1, 1, 1), np.int16)]:
        lo, hi = (-700, 900) if dtype == np.int16 else (0, 3000)
        raw = rng.integers(lo, hi, size=shape).astype(dtype)
        m = rnd_mask(shape, 0.6, 92)
        f = raw.astype(np.float32) * np.float32(0.37) + np.float32(-1.5)
        want = per_volume(oracle.percentiles, f, m, correction=0.5)
        other = rng.random(shape).astype(np.float32)
        img, M, Oth = ctx.upload_scaled(raw, 0.37, -1.5), ctx.upload(m), ctx.upload(other)
        p = ctx.percentiles(img, M, 0.5)
        del img, M                                            # the deferred form keeps what it reads alive
        thr = ctx.cmp_scalar(p, ">", 0.4).numpy()             # reads the deferred form
        assert np.array_equal(thr, (want > np.float32(0.4)).astype(np.uint8)), shape
        s1 = ctx.arith(p, "+", Oth).numpy()                   # two f32 leaves, one deferred
        assert np.array_equal(s1.view(np.uint32), (want + other).view(np.uint32)), shape
        vol = ctx.volume(ctx.cmp_scalar(p, "<=", 0.25))
        assert np.array_equal(vol, (want <= np.float32(0.25)).reshape(shape[0], -1).sum(1).astype(np.float64))
        results, errors = {}, []

        def work(i):
            try:
                if i % 2:
                    results[i] = ctx.cmp_scalar(p, ">=", 0.1 * i).numpy()
                else:
                    results[i] = ctx.max(p)                   # a reduction: makes the array
            except Exception as e:  # pragma: no cover
                errors.append(e)

        ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errors
        for i in range(8):
            if i % 2:
                assert np.array_equal(results[i], (want >= np.float32(0.1 * i)).astype(np.uint8)), (shape, i)
            else:
                assert np.array_equal(results[i], want.reshape(shape[0], -1).max(1).astype(np.float64)), (shape, i)
        assert np.array_equal(p.numpy().view(np.uint32), want.view(np.uint32)), shape
        thr2 = ctx.cmp_scalar(p, ">", 0.4).numpy()            # the array exists now
        assert np.array_equal(thr2, thr)
    assert ctx.mem_info()["live_handles"] >= 0


def test_deferred_percentile_thresholds_on_codes(ctx, oracle):
    """a deferred percentile image that a program only compares with constants is never looked up: the
    comparison is decided on the code's position against a per-volume break point.  Every comparison, constants
    below / at / above the table's values (0, 1, exact table entries, negative, NaN), a decreasing scaling,
    ties, an empty mask in one volume, several comparisons of one leaf fused in one program -- against the
    oracle's ranks; and corrections outside [0, 1] (no monotone table) take the look-up"""
    from voxlogica_b200 import _abi as A
    rng = np.random.default_rng(97)
    shape = (3, 6, 10, 32)
    ops = {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal}
    for dtype, slope, inter, corr in [(np.int16, 0.25, -3.0, 0.5), (np.uint16, -0.5, 100.0, 0.0), (np.int16, 1.0, 0.0, 1.0),
                                       (np.int16, 0.25, 0.0, 1.5)]:
        lo, hi = (-40, 60) if dtype == np.int16 else (0, 90)           # few distinct values: many ties
        raw = rng.integers(lo, hi, size=shape).astype(dtype)
        m = rnd_mask(shape, 0.5, 98)
        m[1] = 0                                                       # a volume with an empty mask
        f = raw.astype(np.float32) * np.float32(slope) + np.float32(inter)
        want = per_volume(oracle.percentiles, f, m, correction=corr)
        p = ctx.percentiles(ctx.upload_scaled(raw, slope, inter), ctx.upload(m), corr)
        vals = np.unique(want)
        consts = [0.0, 1.0, -0.25, 1.5, 0.5, float("nan"), float(vals[len(vals) // 2]), float(vals[-1]), float(vals[min(1, len(vals) - 1)]),
                  float(np.nextafter(vals[len(vals) // 3], np.float32(2)))]
        for c in consts:
            for name, fn in ops.items():
                got = ctx.cmp_scalar(p, name, c).numpy()
                assert np.array_equal(got, fn(want, np.float32(c)).astype(np.uint8)), (dtype, slope, corr, name, c)
        # two comparisons of the same leaf and a boolean leaf in one fused program: (p > a) & (p <= b) & m
        a, b = float(vals[len(vals) // 4]), float(vals[(3 * len(vals)) // 4])
        prog = [(A.VXP_LOAD, 0, 0, 0, 0, 0.0), (A.VXP_LOAD, 1, 1, 0, 0, 0.0), (A.VXP_CMPS, 2, 0, 0, A.VX_GT, a),
                (A.VXP_CMPS, 3, 0, 0, A.VX_LE, b), (A.VXP_AND, 4, 2, 3, 0, 0.0), (A.VXP_AND, 5, 4, 1, 0, 0.0)]
        got = ctx.fused(prog, [p, ctx.upload(m)]).numpy()
        ref = ((want > np.float32(a)) & (want <= np.float32(b)) & (m != 0)).astype(np.uint8)
        assert np.array_equal(got, ref), (dtype, slope, corr)
        assert np.array_equal(p.numpy().view(np.uint32), want.view(np.uint32))


def test_single_comparison_of_a_coded_image(ctx, oracle):
    """ONE comparison of an image that is read as 16-bit codes -- a 16-bit upload, or a deferred percentile
    image -- with a constant or a per-volume scalar, possibly negated (`!(a >. c)`, what `filter` asks for),
    takes the 32-voxels-per-thread kernel: all six comparisons, NaN and out-of-range constants, voxel counts that
    are not multiples of 32 (volumes share a word of the result), against numpy on the same f32 values and
    against the interpreter (quantised paths off)."""
    from voxlogica_b200 import _abi as A
    rng = np.random.default_rng(131)
    ops = {"<": (np.less, A.VX_LT), "<=": (np.less_equal, A.VX_LE), ">": (np.greater, A.VX_GT),
           ">=": (np.greater_equal, A.VX_GE), "==": (np.equal, A.VX_EQ), "!=": (np.not_equal, A.VX_NE)}
    for shape in [(3, 5, 7, 9), (2, 4, 6, 32), (5, 1, 3, 11)]:
        for dtype, slope, inter in [(np.int16, 0.25, -3.0), (np.uint16, -0.5, 40.0)]:
            lo, hi = (-40, 60) if dtype == np.int16 else (0, 90)
            raw = rng.integers(lo, hi, size=shape).astype(dtype)
            f = raw.astype(np.float32) * np.float32(slope) + np.float32(inter)
            img = ctx.upload_scaled(raw, slope, inter)
            consts = [float(f.flat[3]), float(np.median(f)), -1e9, 1e9, float("nan")]
            for c in consts:
                for name, (fn, code) in ops.items():
                    want = fn(f, np.float32(c)).astype(np.uint8)
                    assert np.array_equal(ctx.cmp_scalar(img, name, c).numpy(), want), (shape, dtype, name, c)
                    prog = [(A.VXP_LOAD, 0, 0, 0, 0, 0.0), (A.VXP_CMPS, 1, 0, 0, code, c), (A.VXP_NOT, 2, 1, 0, 0, 0.0)]
                    assert np.array_equal(ctx.fused(prog, [img]).numpy(), 1 - want), (shape, dtype, "not", name, c)
            # per-volume scalars (VXP_CMPSB)
            bs = np.array([[float(f[v].flat[1]) for v in range(shape[0])]], np.float32)
            prog = [(A.VXP_LOAD, 0, 0, 0, 0, 0.0), (A.VXP_CMPSB, 1, 0, 0, A.VX_GE, 0.0)]
            want = (f >= bs[0][:, None, None, None]).astype(np.uint8)
            assert np.array_equal(ctx.fused(prog, [img], bs).numpy(), want), (shape, dtype, "cmpsb")
            # a deferred percentile image, negated comparison; the interpreter gives the same
            m = rnd_mask(shape, 0.6, 132)
            m[0] = 0
            pw = per_volume(oracle.percentiles, f, m, correction=0.5)
            for opt in (1, 0):
                ctx.set_option(A.VX_OPT_QUANTISED_PATHS, opt)
                try:
                    p = ctx.percentiles(img, ctx.upload(m), 0.5)
                    for c in (0.0, 0.5, float(np.unique(pw)[-1]), 1.0):
                        for name, (fn, code) in list(ops.items())[:4]:
                            prog = [(A.VXP_LOAD, 0, 0, 0, 0, 0.0), (A.VXP_CMPS, 1, 0, 0, code, c), (A.VXP_NOT, 2, 1, 0, 0, 0.0)]
                            want = 1 - fn(pw, np.float32(c)).astype(np.uint8)
                            assert np.array_equal(ctx.fused(prog, [p]).numpy(), want), (shape, dtype, opt, name, c)
                finally:
                    ctx.set_option(A.VX_OPT_QUANTISED_PATHS, 1)


def test_deferred_image_survives_raw_pointers_to_its_operands(ctx, oracle):
    """vx_device_ptr freezes an image (the caller may write through the pointer): a deferred percentile
    image made before keeps the codes and mask bits it reads; one made after takes the general path"""
    rng = np.random.default_rng(93)
    shape = (2, 6, 10, 32)
    raw = rng.integers(0, 500, size=shape).astype(np.int16)
    m = rnd_mask(shape, 0.5, 94)
    f = raw.astype(np.float32)
    want = per_volume(oracle.percentiles, f, m, correction=0.0)
    img, M = ctx.upload_scaled(raw, 1.0, 0.0), ctx.upload(m)
    p = ctx.percentiles(img, M)
    assert ctx.device_ptr(M) != 0 and ctx.device_ptr(img) != 0      # both frozen now
    assert np.array_equal(ctx.cmp_scalar(p, ">", 0.5).numpy(), (want > np.float32(0.5)).astype(np.uint8))
    assert np.array_equal(p.numpy().view(np.uint32), want.view(np.uint32))
    p2 = ctx.percentiles(img, M)                                    # frozen operands: sort path, array at once
    assert np.array_equal(p2.numpy().view(np.uint32), want.view(np.uint32))
    assert np.array_equal(ctx.max(img), f.reshape(2, -1).max(1).astype(np.float64))


def test_extrema_of_quantised_images(ctx):
    """min / max / minmax of an image uploaded as 16-bit codes come from the recorded code range of every
    volume (no pass over the image) and equal the reductions of the f32 image"""
    from voxlogica_b200 import _abi as A
    rng = np.random.default_rng(75)
    for dtype, slope, inter in [(np.int16, 1.0 / 4095.0, 0.0), (np.uint16, -2.5, 1e3), (np.int16, 0.25, -3.0)]:
        for shape in [(3, 6, 20, 32), (2, 3, 5, 7), (1, 1, 1, 1)]:
            info = np.iinfo(dtype)
            raw = np.stack([rng.integers(max(info.min, -500 * (v + 1)), min(info.max, 700 * (v + 2)), size=shape[1:])
                            for v in range(shape[0])]).astype(dtype)
            img = ctx.upload_scaled(raw, slope, inter)
            f = (raw.astype(np.float32) * np.float32(slope) + np.float32(inter)).reshape(shape[0], -1)
            for opt in (1, 0):
                ctx.set_option(A.VX_OPT_QUANTISED_PATHS, opt)
                try:
                    lo, hi = ctx.minmax(img)
                    assert np.array_equal(lo, f.min(1).astype(np.float64)) and np.array_equal(hi, f.max(1).astype(np.float64))
                    assert np.array_equal(ctx.min(img), f.min(1).astype(np.float64))
                    assert np.array_equal(ctx.max(img), f.max(1).astype(np.float64))
                finally:
                    ctx.set_option(A.VX_OPT_QUANTISED_PATHS, 1)


def test_quantisation_tag_needs_a_strictly_monotone_scaling(ctx, oracle):
    """slope 0 or a scaling that collapses neighbouring codes in binary32 gives no tag: the sort decides"""
    rng = np.random.default_rng(73)
    shape = (2, 6, 20, 32)
    raw = rng.integers(-300, 300, size=shape).astype(np.int16)
    m = rnd_mask(shape, 0.8, 74)
    for slope, inter in [(0.0, 1.5), (1e-3, 1e6), (1.0, 3e8)]:
        f = raw.astype(np.float32) * np.float32(slope) + np.float32(inter)
        got = ctx.percentiles(ctx.upload_scaled(raw, slope, inter), ctx.upload(m), 0.5).numpy()
        want = per_volume(oracle.percentiles, f, m, correction=0.5)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (slope, inter)


# ------------------------------------------------------------------ cross correlation
@pytest.mark.parametrize("shape,rho,k", [((2, 9, 20, 48), 2.0, 10), ((1, 7, 13, 37), 3.0, 100),
                                         ((1, 1, 33, 70), 5.0, 16), ((1, 12, 70, 40), 5.0, 100)])
def test_cross_correlation(ctx, oracle, shape, rho, k):
    rng = np.random.default_rng(29)
    a = ndi.gaussian_filter(rng.random(shape).astype(np.float32), (0, 1, 1, 1))
    b = rng.random(shape).astype(np.float32)
    reg = rnd_mask(shape, 0.3, 30)
    A, B, R = ctx.upload(a), ctx.upload(b), ctx.upload(reg)
    m, M = ctx.min(A), ctx.max(A)
    got = ctx.cross_correlation(rho, A, B, R, m, M, k).numpy()
    want = np.stack([oracle.cross_correlation(rho, a[i], b[i], reg[i], m[i], M[i], k) for i in range(shape[0])])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    text = np.stack([oracle.cross_correlation(rho, a[i], b[i], reg[i], m[i], M[i], k, textbook=True)
                     for i in range(shape[0])])
    np.testing.assert_allclose(got, text, rtol=1e-5, atol=1e-6)


def test_cross_correlation_spacing_and_empty_region(ctx, oracle):
    shape = (1, 8, 16, 32)
    rng = np.random.default_rng(31)
    a = rng.random(shape).astype(np.float32)
    sp = (1.0, 2.0, 1.5)
    A = ctx.upload(a, spacing=sp)
    reg = rnd_mask(shape, 0.5, 32)
    got = ctx.cross_correlation(3.0, A, A, ctx.upload(reg, spacing=sp), 0.0, 1.0, 20).numpy()
    want = oracle.cross_correlation(3.0, a[0], a[0], reg[0], 0.0, 1.0, 20, spacing=sp)
    assert np.array_equal(got[0].view(np.uint32), want.view(np.uint32))
    none = ctx.cross_correlation(3.0, A, A, ctx.ff(A), 0.0, 1.0, 20).numpy()
    assert np.all(none == 0)


def test_cross_correlation_constant_regions(ctx, oracle):
    """volumes that are mostly one value (background), where whole warps skip their slides.  Volume 0: zero
    background; volume 1: a counted constant; volume 2: its very first voxel differs from the background"""
    shape = (3, 14, 121, 100)
    rng = np.random.default_rng(41)
    tex = ndi.gaussian_filter(rng.random(shape).astype(np.float32), (0, 1, 1, 1))
    a = np.zeros(shape, np.float32)
    a[1] = 0.37
    a[2] = 0.6
    a[:, 3:11, 50:75, 30:58] = tex[:, 3:11, 50:75, 30:58]
    a[0, 7, 100, 90] = 0.9                      # a lone voxel far inside the background
    a[2, 0, 0, 0] = 0.1
    reg = np.zeros(shape, np.uint8)
    reg[:, 4:9, 55:70, 35:50] = 1
    Aimg, R = ctx.upload(a), ctx.upload(reg)
    for rho, k in ((5.0, 100), (2.0, 16)):
        want = np.stack([oracle.cross_correlation(rho, a[i], a[i], reg[i], 0.0, 1.0, k) for i in range(shape[0])])
        got = ctx.cross_correlation(rho, Aimg, Aimg, R, 0.0, 1.0, k).numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (rho, k)


def test_cross_correlation_flat_rows_at_the_top_and_bottom_of_the_image(ctx, oracle):
    """balls that reach over the first / last rows of the image while everything they see inside it is one
    value: the kernel builds them from row counts and moves voxels between the not-counted code and the
    reference code in closed form (warps whose columns all lie inside the image in x and z; nx = 100 has one
    such chunk, x = 32..63).  Volume 0: counted constant background, texture in the middle rows; volume 1:
    the same with a background that is NOT counted (below m); volume 2: constant everywhere; volume 3:
    texture touching the first rows only.  Also an image with fewer rows than the ball is high (top and
    bottom crossed in the same step)."""
    rng = np.random.default_rng(43)
    shape = (4, 13, 40, 100)
    tex = ndi.gaussian_filter(rng.random(shape).astype(np.float32), (0, 1, 1, 1))
    a = np.full(shape, 0.37, np.float32)
    a[1] = -1.0
    a[0:2, :, 14:26, :] = tex[0:2, :, 14:26, :]
    a[3, :, 0:9, 20:80] = tex[3, :, 0:9, 20:80]
    reg = np.zeros(shape, np.uint8)
    reg[:, 3:10, 16:24, 35:60] = 1
    reg[3, 3:10, 0:6, 35:60] = 1
    Aimg, R = ctx.upload(a), ctx.upload(reg)
    for rho, k in ((5.0, 100), (2.0, 16), (3.0, 33), (4.0, 50)):
        want = np.stack([oracle.cross_correlation(rho, a[i], a[i], reg[i], 0.0, 1.0, k) for i in range(shape[0])])
        got = ctx.cross_correlation(rho, Aimg, Aimg, R, 0.0, 1.0, k).numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (rho, k)
    # anisotropic spacing: the table-driven kernel (column list not compiled in)
    sp = (1.0, 1.25, 1.0)
    As, Rs = ctx.upload(a[:2], spacing=sp), ctx.upload(reg[:2], spacing=sp)
    want = np.stack([oracle.cross_correlation(5.0, a[i], a[i], reg[i], 0.0, 1.0, 40, spacing=sp) for i in range(2)])
    got = ctx.cross_correlation(5.0, As, As, Rs, 0.0, 1.0, 40).numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # 7 rows under a ball of 11
    low = np.full((2, 12, 7, 100), 0.5, np.float32)
    low[1, :, :, 40:60] = tex[1, :12, :7, 40:60]
    lreg = np.zeros(low.shape, np.uint8)
    lreg[:, 2:9, 1:6, 30:70] = 1
    L, LR = ctx.upload(low), ctx.upload(lreg)
    for rho, k in ((5.0, 100), (2.0, 16)):
        want = np.stack([oracle.cross_correlation(rho, low[i], low[i], lreg[i], 0.0, 1.0, k) for i in range(2)])
        got = ctx.cross_correlation(rho, L, L, LR, 0.0, 1.0, k).numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), ("low", rho, k)


def test_cross_correlation_wide_ball_and_out_of_range(ctx, oracle):
    """radius beyond the ring kernel's x halo (the direct sliding kernel), values outside [m, M),
    NaN voxels (never counted), an interval with m == M (nothing counted: all zeros)"""
    shape = (1, 5, 24, 40)
    rng = np.random.default_rng(35)
    a = rng.random(shape).astype(np.float32) * 2 - 0.5        # a third of the values fall outside [0, 1)
    a[0, 2, 3, 4:9] = np.nan
    reg = rnd_mask(shape, 0.4, 36)
    A, R = ctx.upload(a), ctx.upload(reg)
    for rho in (9.0, 0.0):
        got = ctx.cross_correlation(rho, A, A, R, 0.0, 1.0, 12).numpy()[0]
        want = oracle.cross_correlation(rho, a[0], a[0], reg[0], 0.0, 1.0, 12)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), rho
    assert np.all(ctx.cross_correlation(2.0, A, A, R, 0.5, 0.5, 12).numpy() == 0)
    with pytest.raises(Exception):
        ctx.cross_correlation(2.0, A, A, R, 0.0, 1.0, 255)     # more bins than the u8 bin image can hold


def test_large_radius_distance_fallback(ctx, oracle):
    """radii whose ball does not fit the bit-dilation tables take the exact-EDT threshold path"""
    shape = (1, 6, 40, 150)
    a = rnd_mask(shape, 0.0008, 37)
    a[0, 0, 0, 0] = 1
    for sp in ((1.0, 1.0, 1.0), (0.5, 1.0, 2.0)):
        A = ctx.upload(a, spacing=sp)
        for r in (33.0, 40.5):
            assert np.array_equal(ctx.distlt(r, A).numpy()[0], oracle.distlt(r, a[0], sp)), (sp, r)
            assert np.array_equal(ctx.distgt(r, A).numpy()[0], oracle.distgt(r, a[0], sp)), (sp, r)


def test_limits_are_reported(ctx):
    """a volume of 2^31 voxels or more is refused with VX_E_UNSUPPORTED (slab-decompose it),
    shape and dtype mismatches with their own codes -- never a crash"""
    from voxlogica_b200 import _abi as A
    import ctypes as C
    out = A.u64()
    rc = ctx.lib.vx_upload_batch(ctx.handle, C.c_void_p(1), A.VX_U8, 2048, 2048, 512, 1, None, C.byref(out))
    assert rc == A.VX_E_UNSUPPORTED and b"slab" in ctx.lib.vx_last_error()
    a = ctx.upload(np.zeros((2, 3, 4), np.uint8))
    b = ctx.upload(np.zeros((2, 3, 5), np.uint8))
    f = ctx.upload(np.zeros((2, 3, 4), np.float32))
    for fn, code in ((lambda: ctx.and_(a, b), A.VX_E_SHAPE_MISMATCH), (lambda: ctx.near(f), A.VX_E_DTYPE),
                     (lambda: ctx.near(a, 18), A.VX_E_INVALID_ARG), (lambda: ctx.crop_z(a, 1, 5), A.VX_E_INVALID_ARG)):
        with pytest.raises(A.VoxError) as ei:
            fn()
        assert ei.value.code == code


# ------------------------------------------------------------------ runtime
def test_handles_and_memory(ctx):
    import gc
    gc.collect()                 # handles of earlier tests still waiting for the collector would skew the count
    before = ctx.mem_info()
    a = ctx.upload(np.ones((4, 8, 16), np.uint8))
    imgs = [ctx.near(a) for _ in range(50)]
    mid = ctx.mem_info()
    assert mid["live_handles"] == before["live_handles"] + 51
    del imgs
    a.release()
    assert ctx.mem_info()["live_handles"] == before["live_handles"]
    assert ctx.launch_count() > 0


def test_concurrent_callers(ctx, oracle):
    """The F# scheduler calls from many threads: each thread gets its own stream and
    results produced on one stream are consumed on another."""
    import threading
    shape = (1, 8, 24, 48)
    base = rnd_mask(shape, 0.2, 33)
    A = ctx.upload(base)
    want = per_volume(oracle.near, per_volume(oracle.near, base))
    results, errors = {}, []

    def work(i):
        try:
            x = ctx.near(ctx.near(A))        # consumes an image produced on the main thread's stream
            results[i] = x.numpy()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors
    assert all(np.array_equal(r, want) for r in results.values())


@pytest.mark.parametrize("ordered", [1, 0])
def test_concurrent_large_uploads(ctx, ordered):
    """Uploads above 1 MiB from several caller threads at once: chained on the device in arrival order
    (VX_OPT_ORDERED_UPLOADS, the default) or independent -- either way every thread gets its own data."""
    import threading
    from voxlogica_b200 import _abi as A
    shape = (2, 16, 160, 256)          # 1.3 M voxels: 2.6 MB as int16, 5.2 MB as f32
    rng = np.random.default_rng(77)
    raws = [rng.integers(-2000, 2000, size=shape).astype(np.int16) for _ in range(6)]
    got, errors = {}, []

    def work(i):
        try:
            for rep in range(3):
                a = ctx.upload_scaled(raws[i], 0.5, 1.0)
                b = ctx.upload(raws[i].astype(np.float32))
                got[i] = (a.numpy(), b.numpy())
        except Exception as e:  # pragma: no cover
            errors.append(e)

    ctx.set_option(A.VX_OPT_ORDERED_UPLOADS, ordered)
    try:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    finally:
        ctx.set_option(A.VX_OPT_ORDERED_UPLOADS, 1)
    assert not errors
    for i in range(6):
        f = raws[i].astype(np.float32)
        assert np.array_equal(got[i][0], f * np.float32(0.5) + np.float32(1.0))
        assert np.array_equal(got[i][1], f)


# ------------------------------------------------------------------ transfers in the file's own width
@pytest.mark.parametrize("dtype", [np.int16, np.uint16])
def test_upload_scaled_matches_host_conversion(ctx, dtype):
    """16-bit voxels + scl_slope/scl_inter converted on the device (vx_upload_scaled) against the host
    loader's arithmetic raw.astype(f32) * slope + inter: the same bits (two roundings, no FMA)."""
    rng = np.random.default_rng(41)
    info = np.iinfo(dtype)
    for shape in [(2, 5, 7, 13), (1, 3, 16, 24), (1, 1, 1, 1), (3, 2, 3, 5)]:
        raw = rng.integers(info.min, info.max + 1, size=shape).astype(dtype)
        raw.flat[0] = info.min
        raw.flat[-1] = info.max
        for slope, inter in [(1.0, 0.0), (1.0 / 4095.0, 0.0), (0.0007324442, -3.25), (-2.5, 1e-3), (3.0, 1e6)]:
            got = ctx.upload_scaled(raw, slope, inter).numpy()
            want = raw.astype(np.float32) * np.float32(slope) + np.float32(inter)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (shape, slope, inter)
    with pytest.raises(Exception):
        ctx.upload_scaled(np.zeros((2, 2, 2), np.int32))


def test_download_bits(ctx):
    rng = np.random.default_rng(43)
    for shape in [(2, 5, 7, 13), (1, 3, 16, 32), (1, 1, 1, 1), (1, 1, 1, 33), (3, 4, 9, 31)]:
        a = (rng.random(shape) < 0.4).astype(np.uint8)
        img = ctx.upload(a)
        assert np.array_equal(ctx.download_bits(img), a), shape
    f = ctx.upload(np.zeros((1, 2, 2, 2), np.float32))
    with pytest.raises(Exception):
        ctx.download_bits(f)            # boolean images only
