"""Slab-decomposed volume across N GPUs (torchrun, NCCL): SURVEY.md section 8e row 2 / BASELINE
config 5b.  `--check` compares every slab operator with the single-GPU result of the whole
volume (small sizes); without it the script times the operators on a size^3 volume.

    torchrun --nproc-per-node N scripts/slab_gpu.py --size 96 --check
    torchrun --nproc-per-node 8 scripts/slab_gpu.py --size 2048
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def blob_slab(shape, z0, z1, device, seed=7, coarse=24):
    """planes [z0, z1) of a thresholded band-limited noise volume, generated on the GPU (torch is
    plumbing here: trilinear upsampling of a coarse seeded grid that is identical on all ranks)"""
    import torch
    import torch.nn.functional as F
    nz, ny, nx = shape
    g = torch.Generator().manual_seed(seed)
    cz, cy, cx = nz // coarse + 2, ny // coarse + 2, nx // coarse + 2
    c = torch.rand((1, 1, cz, cy, cx), generator=g).to(device)
    out = torch.empty((z1 - z0, ny, nx), dtype=torch.uint8, device=device)
    ys = (torch.arange(ny, device=device, dtype=torch.float32) + 0.5) / (coarse * (cy - 1)) * 2 - 1
    xs = (torch.arange(nx, device=device, dtype=torch.float32) + 0.5) / (coarse * (cx - 1)) * 2 - 1
    step = 8
    for s in range(z0, z1, step):
        e = min(s + step, z1)
        zs = (torch.arange(s, e, device=device, dtype=torch.float32) + 0.5) / (coarse * (cz - 1)) * 2 - 1
        gz, gy, gx = torch.meshgrid(zs, ys, xs, indexing="ij")
        grid = torch.stack([gx, gy, gz], dim=-1)[None]
        v = F.grid_sample(c, grid, mode="bilinear", padding_mode="border", align_corners=True)[0, 0]
        out[s - z0:e - z0] = (v > 0.58).to(torch.uint8)
    return out


def _last_plane_mask(ops, B, rank, world, local):
    import torch
    nzl, ny, nx = B.data.shape[1:]
    m = torch.zeros((nzl, ny, nx), dtype=torch.uint8, device=f"cuda:{local}")
    if rank == world - 1:
        m[-1] = 1
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--peer", action="store_true", help="fused halo: kernels read neighbour planes from IPC peer memory")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from voxlogica_b200 import Context
    from voxlogica_b200.slab import CudaEngine, Slab, SlabComm, SlabOps, slab_bounds
    ctx = Context(local)
    ops = SlabOps(CudaEngine(ctx), SlabComm())
    n = args.size
    if args.peer:
        ops.enable_peer_halo(n, n)
    shape = (n, n, n)
    z0, z1 = slab_bounds(n, world)[rank]
    sp = (1.0, 1.0, 1.0)
    b_t = blob_slab(shape, z0, z1, f"cuda:{local}")
    B = Slab(ctx.from_tensor(b_t[None], spacing=sp), z0, z1, n, sp)
    seed_t = torch.zeros_like(b_t)
    if rank == 0:
        seed_t[0] = b_t[0]                        # everything connected to the first plane
    A = Slab(ctx.from_tensor(seed_t[None], spacing=sp), z0, z1, n, sp)
    del seed_t

    if args.check:
        # gather the whole volume on every rank, run single-GPU, compare slabs
        parts = [torch.empty((slab_bounds(n, world)[r][1] - slab_bounds(n, world)[r][0], n, n), dtype=torch.uint8,
                             device=f"cuda:{local}") for r in range(world)]
        dist.all_gather(parts, b_t)
        whole = torch.cat(parts).cpu().numpy()
        W = ctx.upload(whole, spacing=sp)
        wseed = np.zeros_like(whole); wseed[0] = whole[0]
        WS = ctx.upload(wseed, spacing=sp)
        f_whole = np.random.default_rng(3).random(shape, dtype=np.float32)
        FW = ctx.upload(f_whole, spacing=sp)
        F = Slab(ctx.upload(f_whole[z0:z1], spacing=sp), z0, z1, n, sp)
        bad = []
        tests = {
            "near26": (lambda: ops.near(B, 26), lambda: ctx.near(W, 26)),
            "interior6": (lambda: ops.interior(B, 6), lambda: ctx.interior(W, 6)),
            "through26": (lambda: ops.through(A, B, 26), lambda: ctx.through(WS, W, 26)),
            "through6": (lambda: ops.through(A, B, 6), lambda: ctx.through(WS, W, 6)),
            "maxvol26": (lambda: ops.maxvol(B, 26), lambda: ctx.maxvol(W, 26)),
            "maxvol6": (lambda: ops.maxvol(B, 6), lambda: ctx.maxvol(W, 6)),
            "percentiles": (lambda: ops.percentiles(F, B, 0.5), lambda: ctx.percentiles(FW, W, 0.5)),
            "distlt5": (lambda: ops.distlt(5.0, B), lambda: ctx.distlt(5.0, W)),
            "flt2": (lambda: ops.flt(2.0, B), lambda: ctx.distlt(2.0, ctx.distgeq(2.0, ctx.not_(W)))),
            "edt": (lambda: ops.edt(B), lambda: ctx.edt(W)),
            "xcorr": (lambda: ops.cross_correlation(3.0, F, F, B, 0.0, 1.0, 32),
                      lambda: ctx.cross_correlation(3.0, FW, FW, W, 0.0, 1.0, 32)),
        }
        for name, (fs, fw) in tests.items():
            got = fs().data.numpy()[0]
            want = fw().numpy()[0][z0:z1]
            if got.dtype == np.float32:
                got, want = got.view(np.uint32), want.view(np.uint32)
            if not np.array_equal(got, want):
                bad.append(name)
        vol = ops.volume(B)
        if vol != float(whole.sum()):
            bad.append("volume")
        # the whole GTV spec, every operator slab-decomposed, against the ImgQL driver on the whole volume
        from voxlogica_b200.imgql import run_gtv
        from voxlogica_b200.synth import brats_like
        flair = brats_like(seed=1234, shape=shape)
        ref = run_gtv(ctx, ctx.upload(flair[None], spacing=sp))
        got_g = ops.gtv(Slab(ctx.upload(flair[z0:z1], spacing=sp), z0, z1, n, sp))
        for name in ("background", "normFlair", "growTum", "tumSim", "gtv"):
            g_, w_ = got_g[name].data.numpy()[0], ref[name].numpy()[0][z0:z1]
            if g_.dtype == np.float32:
                g_, w_ = g_.view(np.uint32), w_.view(np.uint32)
            if not np.array_equal(g_, w_):
                bad.append("gtv/" + name)
        if got_g["gtvVolume"] != float(ref["gtv"].numpy().sum()):
            bad.append("gtv/volume")
        flags = [None] * world
        dist.all_gather_object(flags, bad)
        if rank == 0:
            print("SLAB CHECK OK" if not any(flags) else f"SLAB CHECK FAILED {flags}", flush=True)
        dist.destroy_process_group()
        sys.exit(0 if not any(flags) else 1)

    del b_t
    torch.cuda.empty_cache()

    def timed(fn):
        for _ in range(1):
            fn()
        ts = []
        for _ in range(args.reps):
            ctx.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            ctx.sync(); torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t[0]))
        return float(np.median(ts))

    res = {}
    fg = ops.volume(B)
    reach = {}
    for name, fn in (("near26", lambda: ops.near(B, 26)), ("through26", lambda: ops.through(A, B, 26)),
                     ("maxvol26", lambda: ops.maxvol(B, 26)),
                     ("distlt5", lambda: ops.distlt(5.0, B)), ("flt2", lambda: ops.flt(2.0, B)),
                     ("edt", lambda: ops.edt(B)), ("volume", lambda: ops.volume(B))):
        s = timed(fn)
        res[name] = {"ms": round(1e3 * s, 3), "gvox_per_s": round(n ** 3 / s / 1e9, 2)}
    reach["foreground_voxels"] = fg
    reach["reached_from_first_plane"] = ops.volume(ops.through(A, B, 26))
    reach["reached_last_plane"] = ops.volume(ops.and_(ops.through(A, B, 26), Slab(
        ctx.from_tensor(_last_plane_mask(ops, B, rank, world, local)[None], spacing=sp), z0, z1, n, sp)))
    if rank == 0:
        print(json.dumps({"slab_bench": {"size": n, "n_gpus": world, "voxels": n ** 3, "peer_halo": bool(args.peer), "ops": res, "check": reach,
                                         "timing": "host clock around device-synchronised regions, max over ranks"}}),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
