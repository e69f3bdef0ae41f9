"""Algebraic laws of the spatial logic on the CUDA path.

The reference's implementation is not available (parity unpinned), but the slides state laws that
ANY correct implementation of the operators must satisfy.  They are checked here on random
inputs drawn by hypothesis, through the C ABI and the ImgQL driver -- no oracle involved."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

pytestmark = pytest.mark.gpu
# derandomize: the same examples every run (a law test must not fail for one run in fifty)
SET = settings(max_examples=12, deadline=None, derandomize=True, database=None,
               suppress_health_check=[HealthCheck.function_scoped_fixture])


@st.composite
def volumes(draw, n=2):
    nz, ny = draw(st.integers(1, 9)), draw(st.integers(1, 20))
    nx = draw(st.sampled_from([1, 5, 16, 31, 32, 33, 48, 70]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    p = draw(st.sampled_from([0.02, 0.15, 0.5, 0.9]))
    rng = np.random.default_rng(seed)
    return [(rng.random((nz, ny, nx)) < p).astype(np.uint8) for _ in range(n)]


@SET
@given(volumes(2), st.sampled_from([6, 26]))
def test_closure_and_interior_laws(ctx, vols, conn):
    """Greta.pdf p.24-25 (closure is extensive, monotone, distributes over union), p.35 (I = !N!)"""
    a, b = vols
    A, B = ctx.upload(a), ctx.upload(b)
    na, nb = ctx.near(A, conn).numpy(), ctx.near(B, conn).numpy()
    assert np.all(na[0] >= a)
    assert np.array_equal(ctx.near(ctx.or_(A, B), conn).numpy(), na | nb)
    assert np.array_equal(ctx.near(ctx.and_(A, B), conn).numpy() <= (na & nb), np.ones_like(na, bool))
    assert np.array_equal(ctx.interior(A, conn).numpy(), 1 - ctx.near(ctx.not_(A), conn).numpy())
    assert np.all(ctx.interior(A, conn).numpy()[0] <= a)
    assert np.array_equal(ctx.near(ctx.ff(A), conn).numpy(), np.zeros_like(na))
    assert np.array_equal(ctx.interior(ctx.tt(A), conn).numpy(), np.ones_like(na))


@SET
@given(volumes(2), st.sampled_from([6, 26]))
def test_reachability_laws(ctx, vols, conn):
    """through(a, b): inside b, contains a & b, idempotent, depends on a only through a & b,
    monotone in a, a union of whole components (closed under one more step inside b);
    p.36 duality  a R b = !((!b) S (!a))  through the driver's stdlib definitions"""
    from voxlogica_b200.imgql import run_spec
    a, b = vols
    A, B = ctx.upload(a), ctx.upload(b)
    t = ctx.through(A, B, conn)
    tn = t.numpy()[0]
    assert np.all(tn <= b) and np.all((a & b) <= tn)
    assert np.array_equal(ctx.through(t, B, conn).numpy()[0], tn)
    assert np.array_equal(ctx.through(ctx.and_(A, B), B, conn).numpy()[0], tn)
    assert np.array_equal(ctx.and_(ctx.near(t, conn), B).numpy()[0], tn)      # nothing of b adjacent is left out
    a2 = a | np.roll(a, 1, axis=1)
    assert np.all(ctx.through(ctx.upload(a2), B, conn).numpy()[0] >= tn)
    assert np.array_equal(ctx.through(B, B, conn).numpy()[0], b)
    # every component is reached from its own first voxel: labels partition b
    lab = ctx.components(B, conn).numpy()[0]
    assert np.array_equal(lab != 0, b != 0)
    if b.any():
        first = np.zeros_like(b).ravel()
        first[np.unique(lab[lab != 0]) - 1] = 1
        assert np.array_equal(ctx.through(ctx.upload(first.reshape(b.shape)), B, conn).numpy()[0], b)
    it = run_spec(ctx, '''
        import "stdlib.imgql"
        load a = "a"
        load b = "b"
        let r = reach(a, b)
        let s = surrounded(a, b)
        let g = grow(a, b)
        let t = touch(b, a)
    ''', {"a": A, "b": B}, conn=conn)
    g = it.globals
    assert np.all(g["g"].numpy()[0] >= a) and np.all(g["t"].numpy()[0] <= b)
    assert np.array_equal(g["g"].numpy()[0], a | g["t"].numpy()[0])
    assert np.all(g["s"].numpy()[0] <= a)                                  # surrounded(a,b) is a part of a


@SET
@given(volumes(1), st.sampled_from([1.0, 1.5, 2.0, 3.2]), st.sampled_from([(1.0, 1.0, 1.0), (0.7, 1.0, 2.5)]))
def test_distance_laws(ctx, vols, r, sp):
    """p.37 distance modality: dist ops are thresholds of one map; distlt/distgeq are complements;
    the Euclidean opening filter(r, a) is anti-extensive and idempotent; distances are 1-Lipschitz"""
    (a,) = vols
    A = ctx.upload(a, spacing=sp)
    d = ctx.edt(A).numpy()[0]
    assert np.all(d[a != 0] == 0) and (np.all(d[a == 0] > 0) or not a.any())
    for op, f in ((ctx.distlt, np.less), (ctx.distleq, np.less_equal), (ctx.distgeq, np.greater_equal), (ctx.distgt, np.greater)):
        assert np.array_equal(op(r, A).numpy()[0], f(d, np.float32(r)).astype(np.uint8))
    assert np.array_equal(ctx.distlt(r, A).numpy(), 1 - ctx.distgeq(r, A).numpy())
    flt = lambda X: ctx.distlt(r, ctx.distgeq(r, ctx.not_(X)))
    o = flt(A)
    assert np.all(o.numpy()[0] <= a)
    assert np.array_equal(flt(o).numpy(), o.numpy())
    if a.any():
        for axis, s in ((2, sp[0]), (1, sp[1]), (0, sp[2])):
            if d.shape[axis] > 1:
                # both distances carry a binary32 rounding of up to half an ulp of the largest one
                slack = 2.0 * float(np.spacing(np.float32(d.max())))
                assert np.all(np.abs(np.diff(d.astype(np.float64), axis=axis)) <= s * (1 + 1e-6) + slack)


@SET
@given(volumes(1), st.integers(0, 2 ** 31 - 1))
def test_reduction_and_rank_laws(ctx, vols, seed):
    """volume is additive; percentiles are a monotone map of the values into [0, 1) that is
    invariant under increasing transformations of the intensities; maxvol is a component"""
    (m,) = vols
    rng = np.random.default_rng(seed)
    f = rng.standard_normal(m.shape).astype(np.float32)
    M, F = ctx.upload(m), ctx.upload(f)
    assert ctx.volume(M)[0] + ctx.volume(ctx.not_(M))[0] == m.size
    p = ctx.percentiles(F, M).numpy()[0]
    assert np.all(p[m == 0] == 0) and np.all((p >= 0) & (p < 1))
    if m.sum() > 1:
        vals, ranks = f[m != 0], p[m != 0]
        order = np.argsort(vals, kind="stable")
        assert np.all(np.diff(ranks[order]) >= 0)
        p2 = ctx.percentiles(ctx.upload(np.exp(f.astype(np.float64) / 4).astype(np.float32) * 3 + 1), M).numpy()[0]
        g = np.exp(f.astype(np.float64) / 4).astype(np.float32) * 3 + 1
        if np.unique(g[m != 0]).size == np.unique(vals).size:      # the map stayed injective in binary32
            assert np.array_equal(p2, p)
    mv = ctx.maxvol(M, 26).numpy()[0]
    assert np.all(mv <= m)
    if m.any():
        lab = ctx.components(M, 26).numpy()[0]
        sizes = np.bincount(lab[lab != 0])
        assert mv.sum() % sizes.max() == 0 and mv.sum() >= sizes.max()
