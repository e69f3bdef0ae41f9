"""Parity at BASELINE.json's full sizes: the whole GTV spec at the BraTS shape 240x240x155 against the
oracle composition (bit-exact, the f32 maps included; ~10 s of CPU), size-independent properties
and independent scipy checks, a 256^3 cube and the 512^3 cube of BASELINE config 3 (26-connectivity
components + exact EDT of the three generators) against scipy.ndimage."""
import numpy as np
import pytest
from scipy import ndimage as ndi

pytestmark = pytest.mark.gpu
BRATS = (155, 240, 240)
S26 = ndi.generate_binary_structure(3, 3)
S6 = ndi.generate_binary_structure(3, 1)


@pytest.fixture(scope="module")
def flair():
    from voxlogica_b200.synth import brats_like
    return brats_like(seed=1234, shape=BRATS)


def test_brats_gtv_properties(ctx, flair):
    from voxlogica_b200.imgql import run_gtv
    F = ctx.upload(flair[None])
    r = run_gtv(ctx, F)
    g = {k: r[k].numpy()[0] for k in ("background", "brain", "normFlair", "hyperIntense", "veryIntense",
                                      "growTum", "tumSim", "tumStatCC", "gtv")}
    # independent scipy reconstruction of the first stages
    dark = flair < np.float32(0.1)
    lab, _ = ndi.label(dark, S26)
    border = np.zeros(BRATS, bool)
    border[[0, -1]] = border[:, [0, -1]] = border[:, :, [0, -1]] = True
    touching = np.unique(lab[ndi.binary_dilation(border, S26) & dark])
    bg = np.isin(lab, touching[touching > 0])
    assert np.array_equal(g["background"], bg.astype(np.uint8))
    assert np.array_equal(g["brain"], 1 - g["background"])
    # percentiles: exact ranks against a numpy sort
    vals = np.sort(flair[g["brain"] != 0])
    want = (np.searchsorted(vals, flair, side="left").astype(np.float64) / vals.size).astype(np.float32)
    assert np.array_equal(g["normFlair"], np.where(g["brain"] != 0, want, np.float32(0)))
    # filter = Euclidean opening: scipy's exact EDT gives the same masks (unit spacing: exact integers)
    def flt(r, a):
        d1 = ndi.distance_transform_edt(a != 0)                 # distance to !a
        core = d1 >= r
        return (ndi.distance_transform_edt(~core) < r) if core.any() else np.zeros_like(core)
    assert np.array_equal(g["hyperIntense"], flt(5.0, g["normFlair"] > np.float32(0.95)).astype(np.uint8))
    assert np.array_equal(g["veryIntense"], flt(2.0, g["normFlair"] > np.float32(0.86)).astype(np.uint8))
    # grow(a, b) = a | (components of b touching a)
    lb, _ = ndi.label(g["veryIntense"], S26)
    keep = np.unique(lb[ndi.binary_dilation(g["hyperIntense"], S26) & (g["veryIntense"] != 0)])
    assert np.array_equal(g["growTum"], (g["hyperIntense"].astype(bool) | np.isin(lb, keep[keep > 0])).astype(np.uint8))
    # correlation coefficients are correlation coefficients; the lesion correlates with itself
    assert np.all(np.abs(g["tumSim"]) <= 1.0 + 1e-6) and np.isfinite(g["tumSim"]).all()
    assert g["gtv"].sum() > 0 and np.all(g["gtv"] >= g["growTum"]) and np.all(g["gtv"] <= g["brain"])
    assert float(r["_printed"]["gtvVolume"][0]) == float(g["gtv"].sum())


def test_brats_gtv_matches_oracle_at_full_size(ctx, oracle):
    """BASELINE config 2 at its own size (Greta.pdf p.48): every named intermediate of the spec on
    the benchmark's own input -- int16 voxels with scl_slope, converted on the device -- against the
    oracle on the host-converted image.  normFlair and tumSim compared as bit patterns."""
    from gtv_ref import gtv_oracle
    from voxlogica_b200.imgql import run_gtv
    from voxlogica_b200.synth import brats_like_int16, int16_to_f32
    raw, slope = brats_like_int16(seed=1234, shape=BRATS)
    got = run_gtv(ctx, ctx.upload_scaled(raw[None], slope))
    want = gtv_oracle(oracle, int16_to_f32(raw, slope))
    for name in ("background", "brain", "hyperIntense", "veryIntense", "growTum", "tumStatCC", "gtv"):
        assert np.array_equal(got[name].numpy()[0], want[name]), name
    for name in ("normFlair", "tumSim"):
        assert np.array_equal(got[name].numpy()[0].view(np.uint32), want[name].view(np.uint32)), name
    assert got["_printed"]["gtvVolume"][0] == want["gtv"].sum() > 0
    assert np.array_equal(ctx.download_bits(got["gtv"])[0], want["gtv"])      # the e2e path's result format


def test_brats_batch_equals_single(ctx, flair):
    """one launch over a batch computes exactly what separate launches compute"""
    from voxlogica_b200.imgql import run_gtv
    from voxlogica_b200.synth import brats_like
    other = brats_like(seed=77, shape=BRATS)
    both = run_gtv(ctx, ctx.upload(np.stack([flair, other])))
    one = run_gtv(ctx, ctx.upload(other[None]))
    for k in ("gtv", "tumSim", "normFlair"):
        a, b = both[k].numpy()[1], one[k].numpy()[0]
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), k


@pytest.mark.parametrize("gen", ["blobs", "bernoulli30", "sparse"])
def test_cube256_ccl_and_edt_vs_scipy(ctx, gen):
    from voxlogica_b200 import synth
    shape = (256, 256, 256)
    a = {"blobs": lambda: synth.blob_volume(shape), "bernoulli30": lambda: synth.bernoulli_volume(shape, 0.3),
         "sparse": lambda: synth.bernoulli_volume(shape, 1e-4)}[gen]()
    A = ctx.upload(a)
    for conn, st in ((26, S26), (6, S6)):
        lab = ctx.components(A, conn).numpy()[0]
        ref, n = ndi.label(a, st)
        assert np.array_equal(lab != 0, a != 0)
        # same partition: the map between the two labelings is a bijection
        fg = a != 0
        pairs = np.unique(np.stack([lab[fg].astype(np.int64), ref[fg].astype(np.int64)]), axis=1)
        assert pairs.shape[1] == n == np.unique(pairs[0]).size == np.unique(pairs[1]).size
        # canonical: a label is 1 + the raster index of its component's first voxel
        first = ndi.minimum(np.arange(a.size, dtype=np.int64).reshape(shape), ref, index=np.arange(1, n + 1))
        lut = np.zeros(n + 1, np.int64); lut[1:] = first + 1
        assert np.array_equal(lab.astype(np.int64), lut[ref])
    d = ctx.edt(A).numpy()[0]
    want = ndi.distance_transform_edt(a == 0)
    np.testing.assert_allclose(d, want, rtol=1e-6, atol=0)      # unit spacing: sqrt of an exact integer
    for r in (3.0, 7.5):
        assert np.array_equal(ctx.distlt(r, A).numpy()[0], (want < r).astype(np.uint8))
        assert np.array_equal(ctx.distgeq(r, A).numpy()[0], (want >= r).astype(np.uint8))


@pytest.mark.parametrize("gen", ["blobs", "bernoulli05", "bernoulli30", "bernoulli60", "sparse"])
def test_cube512_ccl_and_edt(ctx, oracle, gen):
    """BASELINE config 3 at its stated size: one 512^3 volume of every generator of the config.
    26-connectivity components (partition + canonical labels) against scipy.ndimage.label; the exact
    EDT of the config's two EDT inputs (blobs, sparse seeds) bit for bit against the CPU oracle (scipy's single-threaded EDT
    needs 70 s per volume at this size; it checks the same kernels at 256^3 above)."""
    from voxlogica_b200 import synth
    shape = (512, 512, 512)
    a = {"blobs": lambda: synth.blob_volume(shape),
         "bernoulli05": lambda: synth.bernoulli_volume(shape, 0.05),
         "bernoulli30": lambda: synth.bernoulli_volume(shape, 0.3),
         "bernoulli60": lambda: synth.bernoulli_volume(shape, 0.6),
         "sparse": lambda: synth.bernoulli_volume(shape, 1e-4)}[gen]()
    A = ctx.upload(a)
    lab = ctx.components(A, 26).numpy()[0]
    ref, n = ndi.label(a, S26)
    assert np.array_equal(lab != 0, a != 0)
    # canonical labels: 1 + the raster index of the component's first voxel; equality with the
    # relabelled scipy result proves that the partition is the same
    first = ndi.minimum(np.arange(a.size, dtype=np.int64).reshape(shape), ref, index=np.arange(1, n + 1))
    lut = np.zeros(n + 1, np.uint32)
    lut[1:] = np.asarray(first, dtype=np.int64) + 1
    assert np.array_equal(lab, lut[ref])
    del lab, ref, lut, first
    if gen in ("blobs", "sparse"):       # the EDT inputs of the config: (ii) blobs, (iii) sparse seeds
        d = ctx.edt(A).numpy()[0]
        want = oracle.edt(a)
        assert np.array_equal(d.view(np.uint32), want.view(np.uint32))
        assert np.array_equal(ctx.distlt(5.0, A).numpy()[0], (want < np.float32(5.0)).astype(np.uint8))
