"""world_size-2 (and 3) gloo runs of the slab decomposition (voxlogica_b200/slab.py) on CPU:
the host logic -- partition, halo exchange, boundary-label merge, scalar all-reduce -- with an
oracle-backed local engine, checked against the oracle on the WHOLE volume."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    for p in (HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["OMP_NUM_THREADS"] = "2"
    import oracle as orc
    from slab_cpu_engine import OracleEngine
    from voxlogica_b200.slab import SlabComm, SlabOps, slab_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        sp = (1.0, 1.0, 1.5)
        ops = SlabOps(OracleEngine(sp), SlabComm())
        rng = np.random.default_rng(42)
        shape = (13, 18, 22)
        up = lambda a, s: np.ascontiguousarray(a)[None]
        fails = []

        def check(name, slab, want):
            z0, z1 = slab_bounds(shape[0], world)[rank]
            got = slab.data[0]
            if not np.array_equal(got.view(np.uint32) if got.dtype == np.float32 else got,
                                  want[z0:z1].view(np.uint32) if want.dtype == np.float32 else want[z0:z1]):
                fails.append(name)

        # a snake that crosses the slab boundaries several times + random clutter
        b = (rng.random(shape) < 0.22).astype(np.uint8)
        for z in range(shape[0]):
            b[z, 3 + (z % 3), 2:20] = 1 if z % 2 == 0 else b[z, 3 + (z % 3), 2:20]
            b[z, 4, 2 if (z // 2) % 2 == 0 else 19] = 1
        a = np.zeros(shape, np.uint8); a[0, 3, 2] = 1; a[shape[0] - 1, 10, 10] = b[shape[0] - 1, 10, 10]
        A, B = ops.scatter_from_global(up, a, sp), ops.scatter_from_global(up, b, sp)
        for conn in (6, 26):
            check(f"near{conn}", ops.near(B, conn), orc.near(b, conn))
            check(f"interior{conn}", ops.interior(B, conn), orc.interior(b, conn))
            check(f"through{conn}", ops.through(A, B, conn), orc.through(a, b, conn))
            check(f"touch{conn}", ops.touch(B, A, conn), orc.touch(b, a, conn))
            check(f"grow{conn}", ops.grow(A, B, conn), orc.grow(a, b, conn))
            ops.use_ccl_state = False                      # the stateless variant of the same merge
            check(f"through{conn}/stateless", ops.through(A, B, conn), orc.through(a, b, conn))
            ops.use_ccl_state = True
            check(f"maxvol{conn}", ops.maxvol(B, conn), orc.maxvol(b, conn))
        # largest component: two equal winners in different slabs + a smaller one that spans all of them
        t = np.zeros(shape, np.uint8)
        t[0:2, 1, 1:9] = 1                                  # 16 voxels, first slab
        t[shape[0] - 2:, 16, 1:9] = 1                       # 16 voxels, last slab
        t[:, 9, 11] = 1                                     # 13 voxels through every slab
        T = ops.scatter_from_global(up, t, sp)
        check("maxvol/ties", ops.maxvol(T, 6), orc.maxvol(t, 6))
        t[:, 9, 12:15] = 1                                  # now the spanning one wins with 52
        check("maxvol/spanning", ops.maxvol(ops.scatter_from_global(up, t, sp), 26), orc.maxvol(t, 26))
        check("maxvol/empty", ops.maxvol(ops.scatter_from_global(up, np.zeros(shape, np.uint8), sp), 26),
              np.zeros(shape, np.uint8))
        sparse = (rng.random(shape) < 0.01).astype(np.uint8)
        S = ops.scatter_from_global(up, sparse, sp)
        for r in (1.0, 2.5, 4.0):
            check(f"distlt{r}", ops.distlt(r, S), orc.distlt(r, sparse, sp))
            check(f"distgeq{r}", ops.distgeq(r, S), orc.distgeq(r, sparse, sp))
            check(f"flt{r}", ops.flt(r, ops.near(S)), orc.flt(r, orc.near(sparse), sp))
        want = orc.edt(sparse, sp)
        check("edt", ops.edt(S), want)
        check("edt_dense", ops.edt(B), orc.edt(b, sp))
        f = rng.random(shape, dtype=np.float32)
        F = ops.scatter_from_global(up, f, sp)
        want = orc.cross_correlation(3.0, f, f, b, 0.0, 1.0, 16, sp)
        check("xcorr", ops.cross_correlation(3.0, F, F, B, 0.0, 1.0, 16), want)
        # rank normalisation: ties (a quarter of the voxels share one value), negative values, both zeros
        g = f.copy()
        g[rng.random(shape) < 0.25] = 0.5
        g[rng.random(shape) < 0.05] *= -1.0
        g[0, 0, :2] = [0.0, -0.0]
        G = ops.scatter_from_global(up, g, sp)
        for corr in (0.0, 0.5):
            check(f"percentiles/{corr}", ops.percentiles(G, B, corr), orc.percentiles(g, b, correction=corr))
        const = np.full(shape, 0.75, np.float32)           # every key equal: one rank receives everything
        check("percentiles/constant", ops.percentiles(ops.scatter_from_global(up, const, sp), B, 0.5),
              orc.percentiles(const, b, correction=0.5))
        few = rng.integers(0, 3, shape).astype(np.float32)  # three distinct values: splitters coincide
        check("percentiles/few", ops.percentiles(ops.scatter_from_global(up, few, sp), B, 0.0), orc.percentiles(few, b))
        empty = np.zeros(shape, np.uint8)
        check("percentiles/empty", ops.percentiles(G, ops.scatter_from_global(up, empty, sp)), np.zeros(shape, np.float32))
        one = empty.copy(); one[shape[0] - 1, 2, 3] = 1     # every masked voxel on the last rank
        check("percentiles/one", ops.percentiles(G, ops.scatter_from_global(up, one, sp)), orc.percentiles(g, one))
        check("border", ops.border(B), orc.border(b))
        # the whole GTV spec on slabs == the oracle composition on the whole volume
        from gtv_ref import gtv_oracle
        from voxlogica_b200.synth import brats_like
        gshape = (40, 64, 80)
        flair = brats_like(seed=1234, shape=gshape)
        gops = SlabOps(OracleEngine((1.0, 1.0, 1.0)), ops.c)
        want_g = gtv_oracle(orc, flair)
        got_g = gops.gtv(gops.scatter_from_global(up, flair, (1.0, 1.0, 1.0)))
        gz0, gz1 = slab_bounds(gshape[0], world)[rank]
        for name in ("background", "brain", "normFlair", "hyperIntense", "veryIntense", "growTum", "tumSim", "tumStatCC", "gtv"):
            gg, ww = got_g[name].data[0], want_g[name][gz0:gz1]
            if gg.dtype == np.float32:
                gg, ww = gg.view(np.uint32), ww.view(np.uint32)
            if not np.array_equal(gg, ww):
                fails.append("gtv/" + name)
        if got_g["gtvVolume"] != float(want_g["gtv"].sum()) or want_g["gtv"].sum() == 0:
            fails.append("gtv/volume")
        # ... and the spec TEXT through the ImgQL driver on slabs (existing .imgql specs run unchanged)
        from voxlogica_b200.imgql import GTV_SPEC
        from voxlogica_b200.slab import run_spec_on_slabs
        it = run_spec_on_slabs(gops, GTV_SPEC, {"flair": gops.scatter_from_global(up, flair, (1.0, 1.0, 1.0))})
        for name in ("brain", "normFlair", "growTum", "tumSim", "gtv"):
            gg, ww = it.globals[name].data[0], want_g[name][gz0:gz1]
            if gg.dtype == np.float32:
                gg, ww = gg.view(np.uint32), ww.view(np.uint32)
            if not np.array_equal(gg, ww):
                fails.append("imgql/" + name)
        if it.printed["gtvVolume"][0] != float(want_g["gtv"].sum()):
            fails.append("imgql/volume")
        extra = run_spec_on_slabs(gops, """
            load f = "f"
            let c = constant(0.25)
            let hot = (f >. c) & !(f >=. 0.9) | ff
            let n = volume(hot & tt)
            let scaled = (f *. 2 +. 1) /. max(f)
            """, {"f": gops.scatter_from_global(up, flair, (1.0, 1.0, 1.0))})
        hot = ((flair > np.float32(0.25)) & ~(flair >= np.float32(0.9)))
        if extra.globals["n"].v[0] != float(hot.sum()):
            fails.append("imgql/pointwise")
        sc = ((flair * np.float32(2) + np.float32(1)) / np.float32(flair.max())).astype(np.float32)
        if not np.array_equal(extra.globals["scaled"].data[0], sc[gz0:gz1]):
            fails.append("imgql/arith")
        if ops.volume(B) != float(b.sum()): fails.append("volume")
        if ops.min(F) != float(f.min()) or ops.max(F) != float(f.max()): fails.append("minmax")
        check("lt", ops.cmp_scalar(F, "<", 0.3), orc.cmp_scalar(f, "<", 0.3))
        q.put((rank, fails))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_ops_match_whole_volume_oracle(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=240) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, fails in res:
        assert not fails, f"rank {rank}: {fails}"


def test_boundary_class_sizes():
    from voxlogica_b200.slab import boundary_class_sizes
    # classes {1,2,3} (sizes 5+7+1), {4} alone (9), {7,8} (2 + 2^33: beyond 32 bits)
    pairs = np.array([[1, 2], [3, 2], [7, 8]], dtype=np.int64)
    sizes = np.array([[1, 5], [2, 7], [3, 1], [4, 9], [7, 2], [8, 2 ** 33]], dtype=np.int64)
    ids, tot = boundary_class_sizes(pairs, sizes)
    assert dict(zip(ids.tolist(), tot.tolist())) == {1: 13, 2: 13, 3: 13, 4: 9, 7: 2 + 2 ** 33, 8: 2 + 2 ** 33}
    ids, tot = boundary_class_sizes(np.zeros((0, 2), np.int64), sizes[:2])
    assert tot.tolist() == [5, 7]
    assert boundary_class_sizes(np.zeros((0, 2), np.int64), np.zeros((0, 2), np.int64))[0].size == 0


def test_boundary_graph_and_bounds():
    from voxlogica_b200.slab import resolve_boundary_graph, slab_bounds
    assert slab_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    with pytest.raises(ValueError):
        slab_bounds(2, 3)
    # chain 1-2-3-4 with 4 reached, separate pair 7-8 not reached
    pairs = np.array([[1, 2], [3, 2], [3, 4], [7, 8]], dtype=np.int64)
    assert sorted(resolve_boundary_graph(pairs, np.array([4]))) == [1, 2, 3]
    assert resolve_boundary_graph(pairs, np.array([], dtype=np.int64)).size == 0
    assert resolve_boundary_graph(np.zeros((0, 2), np.int64), np.array([4])).size == 0
