"""Slab decomposition on the GPU.

(1) single-GPU: two "ranks" as two host threads of one process, each with its own libvoxcuda
    stream, joined by an in-process communicator -- the whole CudaEngine path (zero-copy tensor
    views, crop/concat, halo planes, boundary-label merge) against the single-image result;
(2) multi-GPU: scripts/slab_gpu.py under torchrun with NCCL when the box has >= 2 GPUs.
"""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class ThreadComm:
    """SlabComm look-alike for `world` threads of one process (tensors stay on the device)."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = {}
            self.lock = threading.Lock()

    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world

    def _swap(self, key, value):
        with self.s.lock:
            self.s.slots[(key, self.rank)] = value
        self.s.barrier.wait()
        vals = [self.s.slots[(key, r)] for r in range(self.world)]
        self.s.barrier.wait()
        return vals

    def exchange_planes(self, to_lower, to_upper):
        import torch
        torch.cuda.synchronize()
        vals = self._swap("x", (to_lower.clone(), to_upper.clone()))
        torch.cuda.synchronize()
        lo = vals[self.rank - 1][1] if self.rank > 0 else None
        hi = vals[self.rank + 1][0] if self.rank + 1 < self.world else None
        return lo, hi

    def barrier(self):
        self.s.barrier.wait()

    def share_mailbox(self, ctx, ptr, handle):
        ptrs = self._swap("m", ptr)            # same process: the raw device pointers are valid
        return (ptrs[self.rank - 1] if self.rank > 0 else None,
                ptrs[self.rank + 1] if self.rank + 1 < self.world else None)

    def exchange(self, send, recv_shapes):
        import torch
        torch.cuda.synchronize()
        allsend = self._swap("e", [t.clone() for t in send])
        torch.cuda.synchronize()
        return [allsend[r][self.rank] for r in range(self.world)]

    def all_gather_rows(self, t):
        import torch
        torch.cuda.synchronize()
        return torch.cat(self._swap("g", t.clone()), dim=0)

    def all_reduce(self, values, op, device):
        arr = np.array(self._swap("r", list(values)), dtype=np.float64)
        return {"sum": arr.sum(0), "min": arr.min(0), "max": arr.max(0)}[op]

    def all_reduce_i64(self, arr, device):
        return np.sum(np.array(self._swap("i", np.asarray(arr, dtype=np.int64))), axis=0)


def test_slab_helpers_single_gpu(ctx):
    rng = np.random.default_rng(1)
    v = rng.random((2, 9, 6, 20), dtype=np.float32)
    img = ctx.upload(v)
    assert np.array_equal(ctx.crop_z(img, 2, 4).numpy(), v[:, 2:6])
    parts = [ctx.crop_z(img, 0, 3), ctx.crop_z(img, 3, 1), ctx.crop_z(img, 4, 5)]
    assert np.array_equal(ctx.concat_z(parts).numpy(), v)
    t = ctx.as_tensor(img)
    ctx.sync()
    assert tuple(t.shape) == v.shape and np.array_equal(t.cpu().numpy(), v)
    back = ctx.from_tensor(t * 2.0)
    assert np.array_equal(back.numpy(), v * 2.0)
    m = (rng.random(v.shape) < 0.5).astype(np.uint8)
    mi = ctx.upload(m)
    h2 = ctx.region_histogram(img, mi, 0.0, 1.0, 16)
    want = np.stack([np.bincount(np.minimum((v[b][m[b] != 0].astype(np.float64) * 16).astype(int), 15),
                                 minlength=16) for b in range(2)])
    assert np.array_equal(h2, want.astype(np.uint64))
    a = ctx.cross_correlation(2.0, img, img, mi, 0.0, 1.0, 16).numpy()
    b = ctx.cross_correlation_h2(2.0, img, h2, 0.0, 1.0, 16).numpy()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("world,peer", [(2, False), (3, False), (2, True), (3, True)])
def test_slab_ops_two_threads_one_gpu(ctx, world, peer):
    from voxlogica_b200.slab import CudaEngine, SlabOps, slab_bounds
    rng = np.random.default_rng(42)
    shape = (21, 40, 64)
    sp = (1.0, 1.0, 1.5)
    b = (rng.random(shape) < 0.2).astype(np.uint8)
    for z in range(shape[0]):
        b[z, 3 + (z % 3), 2:60] |= np.uint8(z % 2 == 0)
        b[z, 4, 2 if (z // 2) % 2 == 0 else 59] = 1
    a = np.zeros(shape, np.uint8); a[0, 3, 2] = 1
    sparse = (rng.random(shape) < 0.004).astype(np.uint8)
    f = rng.random(shape, dtype=np.float32)
    Bi, Ai, Si, Fi = (ctx.upload(x, spacing=sp) for x in (b, a, sparse, f))
    want = {}
    for conn in (6, 26):
        want[f"near{conn}"] = ctx.near(Bi, conn).numpy()[0]
        want[f"interior{conn}"] = ctx.interior(Bi, conn).numpy()[0]
        want[f"through{conn}"] = ctx.through(Ai, Bi, conn).numpy()[0]
        want[f"maxvol{conn}"] = ctx.maxvol(Bi, conn).numpy()[0]
    # two equally large components in the first and the last slab: both must survive
    t = np.zeros(shape, np.uint8)
    t[0:2, 1, 1:13] = 1; t[shape[0] - 2:, 16, 1:13] = 1; t[:, 9, 11] = 1    # 24, 24 and 21 voxels
    Ti = ctx.upload(t, spacing=sp)
    want["maxvol_ties"] = ctx.maxvol(Ti, 6).numpy()[0]
    assert want["maxvol_ties"].sum() == 48 and want["maxvol26"].sum() > 0
    for r in (2.5, 4.0):
        want[f"distlt{r}"] = ctx.distlt(r, Si).numpy()[0]
        want[f"flt{r}"] = ctx.distlt(r, ctx.distgeq(r, ctx.not_(ctx.near(Si)))).numpy()[0]
    want["xcorr"] = ctx.cross_correlation(3.0, Fi, Fi, Bi, 0.0, 1.0, 16).numpy()[0]
    # rank normalisation with ties, negative values and both zeros; mask = b
    gvals = f.copy()
    gvals[rng.random(shape) < 0.25] = 0.5
    gvals[rng.random(shape) < 0.05] *= -1.0
    gvals[0, 0, :2] = [0.0, -0.0]
    Gi = ctx.upload(gvals, spacing=sp)
    for corr in (0.0, 0.5):
        want[f"percentiles{corr}"] = ctx.percentiles(Gi, Bi, corr).numpy()[0]
    one = np.zeros(shape, np.uint8); one[shape[0] - 1, 2, 3] = 1
    want["percentiles_one"] = ctx.percentiles(Gi, ctx.upload(one, spacing=sp)).numpy()[0]
    want["edt"] = ctx.edt(Si).numpy()[0]
    want["percentiles_empty"] = np.zeros(shape, np.float32)
    assert want["through6"].sum() > a.sum() and want["through6"][shape[0] - 1].any()   # crosses every slab
    shared = ThreadComm.Shared(world)
    errors, results = [], {}

    def run(rank):
        try:
            ops = SlabOps(CudaEngine(ctx), ThreadComm(shared, rank))
            if peer:      # fused halo: boundary planes read from the neighbour's mailbox by the kernel
                ops.enable_peer_halo(shape[1], shape[2])
            else:         # and the three-pass `through` without the kept union-find state
                ops.use_ccl_state = False
            up = lambda arr, s: ctx.upload(arr, spacing=s)
            A, B, S, F = (ops.scatter_from_global(up, x, sp) for x in (a, b, sparse, f))
            got = {}
            for conn in (6, 26):
                got[f"near{conn}"] = ops.near(B, conn)
                got[f"interior{conn}"] = ops.interior(B, conn)
                got[f"through{conn}"] = ops.through(A, B, conn)
                got[f"maxvol{conn}"] = ops.maxvol(B, conn)
            got["maxvol_ties"] = ops.maxvol(ops.scatter_from_global(up, t, sp), 6)
            for r in (2.5, 4.0):
                got[f"distlt{r}"] = ops.distlt(r, S)
                got[f"flt{r}"] = ops.flt(r, ops.near(S))
            got["xcorr"] = ops.cross_correlation(3.0, F, F, B, 0.0, 1.0, 16)
            G = ops.scatter_from_global(up, gvals, sp)
            for corr in (0.0, 0.5):
                got[f"percentiles{corr}"] = ops.percentiles(G, B, corr)
            got["percentiles_one"] = ops.percentiles(G, ops.scatter_from_global(up, one, sp))
            got["percentiles_empty"] = ops.percentiles(G, ops.scatter_from_global(up, np.zeros(shape, np.uint8), sp))
            got["edt"] = ops.edt(S)
            vol = ops.volume(B)
            z0, z1 = slab_bounds(shape[0], world)[rank]
            for k, sl in got.items():
                g = sl.data.numpy()[0]
                w = want[k][z0:z1]
                if not np.array_equal(g.view(np.uint32) if g.dtype == np.float32 else g,
                                      w.view(np.uint32) if w.dtype == np.float32 else w):
                    errors.append((rank, k))
            if vol != float(b.sum()):
                errors.append((rank, "volume"))
            results[rank] = True
        except Exception as ex:  # pragma: no cover
            errors.append((rank, repr(ex)))
            shared.barrier.abort()

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ths]
    [t.join(timeout=300) for t in ths]
    assert not errors, errors
    assert len(results) == world


@pytest.mark.parametrize("world,peer", [(2, True), (3, False)])
def test_gtv_spec_on_slabs_equals_whole_volume(ctx, world, peer):
    """the whole GTV spec with every operator slab-decomposed (threads as ranks on one GPU) against the
    ImgQL driver on the whole volume: every named intermediate, the f32 maps bit for bit"""
    from voxlogica_b200.imgql import run_gtv
    from voxlogica_b200.slab import CudaEngine, SlabOps, slab_bounds
    from voxlogica_b200.synth import brats_like
    shape, sp = (48, 80, 96), (1.0, 1.0, 1.0)
    flair = brats_like(seed=1234, shape=shape)
    whole = run_gtv(ctx, ctx.upload(flair[None], spacing=sp))
    names = ("background", "brain", "normFlair", "hyperIntense", "veryIntense", "growTum", "tumSim", "tumStatCC", "gtv")
    want = {k: whole[k].numpy()[0] for k in names}
    assert want["gtv"].sum() > 0
    shared = ThreadComm.Shared(world)
    errors, done = [], {}

    def run(rank):
        try:
            ops = SlabOps(CudaEngine(ctx), ThreadComm(shared, rank))
            if peer:
                ops.enable_peer_halo(shape[1], shape[2])
            F = ops.scatter_from_global(lambda arr, s: ctx.upload(arr, spacing=s), flair, sp)
            got = ops.gtv(F)
            z0, z1 = slab_bounds(shape[0], world)[rank]
            for k in names:
                g_, w_ = got[k].data.numpy()[0], want[k][z0:z1]
                if g_.dtype == np.float32:
                    g_, w_ = g_.view(np.uint32), w_.view(np.uint32)
                if not np.array_equal(g_, w_):
                    errors.append((rank, k))
            if got["gtvVolume"] != float(want["gtv"].sum()):
                errors.append((rank, "gtvVolume"))
            # the spec TEXT through the ImgQL driver on slabs
            from voxlogica_b200.imgql import GTV_SPEC
            from voxlogica_b200.slab import run_spec_on_slabs
            it = run_spec_on_slabs(ops, GTV_SPEC, {"flair": F})
            for k in ("normFlair", "tumSim", "gtv"):
                g_, w_ = it.globals[k].data.numpy()[0], want[k][z0:z1]
                if g_.dtype == np.float32:
                    g_, w_ = g_.view(np.uint32), w_.view(np.uint32)
                if not np.array_equal(g_, w_):
                    errors.append((rank, "imgql/" + k))
            if it.printed["gtvVolume"][0] != float(want["gtv"].sum()):
                errors.append((rank, "imgql/gtvVolume"))
            done[rank] = True
        except Exception as ex:  # pragma: no cover
            errors.append((rank, repr(ex)))
            shared.barrier.abort()

    ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ths]
    [t.join(timeout=300) for t in ths]
    assert not errors, errors
    assert len(done) == world


@pytest.mark.parametrize("peer", [False, True])
def test_slab_multi_gpu_nccl(peer):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 4)}",
           "--master-addr", "127.0.0.1", "--master-port", "29517" if not peer else "29518",
           os.path.join(ROOT, "scripts", "slab_gpu.py"), "--size", "96", "--check"] + (["--peer"] if peer else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SLAB CHECK OK" in out.stdout
