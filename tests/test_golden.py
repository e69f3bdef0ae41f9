"""Committed fixtures (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py).

PARITY UNPINNED: the reference ships no vectors and cannot be run here, so
the fixtures hold this repo's oracle outputs plus independent scipy/OpenCV outputs of the same
seeded inputs.  CPU tests pin the oracle to the fixtures and the fixtures to scipy; the GPU tests
pin the CUDA path (through the C ABI) to the fixtures.
"""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BIN_SHAPE = (20, 24, 28)
GTV_SHAPE = (32, 56, 64)
SPACING = (0.5, 1.25, 3.0)


@pytest.fixture(scope="module")
def G():
    return dict(np.load(os.path.join(HERE, "golden", "golden_v1.npz")))


def unpack(G, key, shape):
    n = int(np.prod(shape))
    return np.unpackbits(G[key])[:n].reshape(shape)


# ------------------------------------------------------------------ fixtures vs independent code
@pytest.mark.parametrize("conn", [6, 26])
def test_fixture_oracle_equals_scipy(G, conn):
    for op in ("near", "interior", "through", "maxvol"):
        assert np.array_equal(G[f"bin.{op}{conn}.oracle"], G[f"bin.{op}{conn}.scipy"]), op
    assert np.array_equal(G[f"bin.components{conn}.oracle"], G[f"bin.components{conn}.scipy"])


def test_fixture_edt_equals_scipy(G):
    for name in ("a", "sparse"):
        for tag in ("unit", "aniso"):
            o, s = G[f"bin.edt.{name}.{tag}.oracle"], G[f"bin.edt.{name}.{tag}.scipy"]
            np.testing.assert_allclose(o, s, rtol=1e-5, atol=0)   # north_star tolerance for distances
    assert np.array_equal(G["maze.fromStart6.oracle"], G["maze.fromStart6.scipy"])
    np.testing.assert_allclose(G["num.percentiles0.0.oracle"], G["num.percentiles0.0.numpy"], rtol=1e-6)


# ------------------------------------------------------------------ oracle vs fixtures (regression pin)
def test_oracle_reproduces_fixtures(G, oracle):
    a, seeds, sparse = (unpack(G, k, BIN_SHAPE) for k in ("bin.a", "bin.seeds", "bin.sparse"))
    for conn in (6, 26):
        assert np.array_equal(oracle.near(a, conn), unpack(G, f"bin.near{conn}.oracle", BIN_SHAPE))
        assert np.array_equal(oracle.through(seeds, a, conn), unpack(G, f"bin.through{conn}.oracle", BIN_SHAPE))
        assert np.array_equal(oracle.components(a, conn), G[f"bin.components{conn}.oracle"])
    assert np.array_equal(oracle.edt(sparse, SPACING).view(np.uint32), G["bin.edt.sparse.aniso.oracle"].view(np.uint32))
    f, m = G["num.f"], unpack(G, "num.mask", BIN_SHAPE)
    assert np.array_equal(oracle.percentiles(f, m, 0.5).view(np.uint32), G["num.percentiles0.5.oracle"].view(np.uint32))
    got = oracle.cross_correlation(3.0, f, f, m, 0.0, 1.0, 100)
    assert np.array_equal(got.view(np.uint32), G["num.xcorr.rho3.0.k100.oracle"].view(np.uint32))


def test_oracle_gtv_reproduces_fixture(G, oracle):
    from gtv_ref import gtv_oracle
    from voxlogica_b200.synth import brats_like
    vol = brats_like(seed=int(G["gtv.seed"][0]), shape=GTV_SHAPE)
    sha = np.frombuffer(hashlib.sha256(vol.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, G["gtv.flair.sha"]), "synthetic generator changed: regenerate the fixtures"
    r = gtv_oracle(oracle, vol)
    for name in ("background", "brain", "hyperIntense", "veryIntense", "growTum", "tumStatCC", "gtv"):
        assert np.array_equal(r[name], unpack(G, f"gtv.{name}.oracle", GTV_SHAPE)), name
    assert np.array_equal(r["tumSim"].view(np.uint32), G["gtv.tumSim.oracle"].view(np.uint32))
    assert unpack(G, "gtv.gtv.oracle", GTV_SHAPE).sum() > 0


# ------------------------------------------------------------------ CUDA path vs fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("conn", [6, 26])
def test_cuda_binary_ops_match_fixtures(G, ctx, conn):
    a, seeds = (ctx.upload(unpack(G, k, BIN_SHAPE)) for k in ("bin.a", "bin.seeds"))
    for name, got in (("near", ctx.near(a, conn)), ("interior", ctx.interior(a, conn)),
                      ("through", ctx.through(seeds, a, conn)), ("maxvol", ctx.maxvol(a, conn))):
        want = unpack(G, f"bin.{name}{conn}.scipy", BIN_SHAPE)     # the INDEPENDENT result
        assert np.array_equal(got.numpy()[0], want), name
    assert np.array_equal(ctx.components(a, conn).numpy()[0], G[f"bin.components{conn}.scipy"])


@pytest.mark.gpu
def test_cuda_distance_ops_match_fixtures(G, ctx):
    for name in ("a", "sparse"):
        src = unpack(G, f"bin.{name}", BIN_SHAPE)
        for tag, sp in (("unit", (1.0, 1.0, 1.0)), ("aniso", SPACING)):
            got = ctx.edt(ctx.upload(src, spacing=sp)).numpy()[0]
            assert np.array_equal(got.view(np.uint32), G[f"bin.edt.{name}.{tag}.oracle"].view(np.uint32))
            np.testing.assert_allclose(got, G[f"bin.edt.{name}.{tag}.scipy"], rtol=1e-5, atol=0)
    sparse = ctx.upload(unpack(G, "bin.sparse", BIN_SHAPE))
    for r in (1.0, 2.0, 2.5, 5.0):
        assert np.array_equal(ctx.distlt(r, sparse).numpy()[0], unpack(G, f"bin.distlt{r}.oracle", BIN_SHAPE))
        assert np.array_equal(ctx.distgeq(r, sparse).numpy()[0], unpack(G, f"bin.distgeq{r}.oracle", BIN_SHAPE))
        x = ctx.near(ctx.near(sparse))
        flt = ctx.distlt(r, ctx.distgeq(r, ctx.not_(x)))
        assert np.array_equal(flt.numpy()[0], unpack(G, f"bin.flt{r}.oracle", BIN_SHAPE))


@pytest.mark.gpu
def test_cuda_numeric_ops_match_fixtures(G, ctx):
    f, m = ctx.upload(G["num.f"]), ctx.upload(unpack(G, "num.mask", BIN_SHAPE))
    for corr in (0.0, 0.5):
        got = ctx.percentiles(f, m, corr).numpy()[0]
        assert np.array_equal(got.view(np.uint32), G[f"num.percentiles{corr}.oracle"].view(np.uint32))
    for rho, k in ((2.0, 16), (3.0, 100)):
        got = ctx.cross_correlation(rho, f, f, m, 0.0, 1.0, k).numpy()[0]
        assert np.array_equal(got.view(np.uint32), G[f"num.xcorr.rho{rho}.k{k}.oracle"].view(np.uint32))
    fa = ctx.upload(G["num.f"], spacing=SPACING)
    ma = ctx.upload(unpack(G, "num.mask", BIN_SHAPE), spacing=SPACING)
    got = ctx.cross_correlation(3.0, fa, fa, ma, 0.1, 0.9, 32).numpy()[0]
    assert np.array_equal(got.view(np.uint32), G["num.xcorr.aniso.oracle"].view(np.uint32))
    sc = G["num.scalars.oracle"]
    got = [ctx.volume(m)[0], ctx.min(f)[0], ctx.max(f)[0], ctx.sum(f)[0], ctx.min(f, m)[0], ctx.max(f, m)[0]]
    assert np.array_equal(np.array(got)[[0, 1, 2, 4, 5]], sc[[0, 1, 2, 4, 5]])
    assert abs(got[3] - sc[3]) <= 1e-9 * abs(sc[3])      # sum: binary64, different association order
    assert np.array_equal(ctx.cmp_scalar(f, "<", 0.4).numpy()[0], unpack(G, "num.lt0.4.oracle", BIN_SHAPE))
    assert np.array_equal(ctx.arith_scalar(f, "*", 1.7).numpy()[0].view(np.uint32), G["num.mul.oracle"].view(np.uint32))


@pytest.mark.gpu
def test_cuda_gtv_and_maze_match_fixtures(G, ctx):
    from voxlogica_b200.imgql import run_gtv
    from voxlogica_b200.synth import brats_like
    vol = brats_like(seed=int(G["gtv.seed"][0]), shape=GTV_SHAPE)
    got = run_gtv(ctx, ctx.upload(vol[None]))
    for name in ("background", "brain", "hyperIntense", "veryIntense", "growTum", "tumStatCC", "gtv"):
        assert np.array_equal(got[name].numpy()[0], unpack(G, f"gtv.{name}.oracle", GTV_SHAPE)), name
    assert np.array_equal(got["normFlair"].numpy()[0].view(np.uint32), G["gtv.normFlair.oracle"].view(np.uint32))
    assert np.array_equal(got["tumSim"].numpy()[0].view(np.uint32), G["gtv.tumSim.oracle"].view(np.uint32))
    img = G["maze.rgb"]
    path = (img.sum(axis=2) > 0).astype(np.uint8)
    start = np.zeros_like(path); start[1, 1] = 1
    for conn in (6, 26):
        res = ctx.through(ctx.upload(start), ctx.upload(path), conn).numpy()[0, 0]
        assert np.array_equal(res, unpack(G, f"maze.fromStart{conn}.oracle", path.shape))
