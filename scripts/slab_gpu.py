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


def blob_slab(shape, z0, z1, seed=7, coarse=24):
    """thresholded band-limited noise, generated per slab (deterministic in the global z)"""
    nz, ny, nx = shape
    rng = np.random.default_rng(seed)
    cz, cy, cx = nz // coarse + 2, ny // coarse + 2, nx // coarse + 2
    c = rng.random((cz, cy, cx)).astype(np.float32)
    from scipy import ndimage as ndi
    zz = (np.arange(z0, z1, dtype=np.float32) + 0.5) / coarse
    yy = (np.arange(ny, dtype=np.float32) + 0.5) / coarse
    xx = (np.arange(nx, dtype=np.float32) + 0.5) / coarse
    out = np.empty((z1 - z0, ny, nx), np.uint8)
    step = 16
    for s in range(0, z1 - z0, step):
        g = np.meshgrid(zz[s:s + step], yy, xx, indexing="ij")
        v = ndi.map_coordinates(c, [a.ravel() for a in g], order=1, mode="nearest").reshape(g[0].shape)
        out[s:s + step] = v > 0.58
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
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
    shape = (n, n, n)
    z0, z1 = slab_bounds(n, world)[rank]
    sp = (1.0, 1.0, 1.0)
    b_local = blob_slab(shape, z0, z1)
    B = Slab(ctx.upload(b_local, spacing=sp), z0, z1, n, sp)
    seed = np.zeros_like(b_local)
    if rank == 0:
        seed[0] = b_local[0]                      # everything connected to the first plane
    A = Slab(ctx.upload(seed, spacing=sp), z0, z1, n, sp)

    if args.check:
        # gather the whole volume on every rank, run single-GPU, compare slabs
        parts = [torch.empty((slab_bounds(n, world)[r][1] - slab_bounds(n, world)[r][0], n, n), dtype=torch.uint8,
                             device=f"cuda:{local}") for r in range(world)]
        dist.all_gather(parts, torch.as_tensor(b_local, device=f"cuda:{local}"))
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
            "distlt5": (lambda: ops.distlt(5.0, B), lambda: ctx.distlt(5.0, W)),
            "flt2": (lambda: ops.flt(2.0, B), lambda: ctx.distlt(2.0, ctx.distgeq(2.0, ctx.not_(W)))),
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
        flags = [None] * world
        dist.all_gather_object(flags, bad)
        if rank == 0:
            print("SLAB CHECK OK" if not any(flags) else f"SLAB CHECK FAILED {flags}", flush=True)
        dist.destroy_process_group()
        sys.exit(0 if not any(flags) else 1)

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
    for name, fn in (("near26", lambda: ops.near(B, 26)), ("through26", lambda: ops.through(A, B, 26)),
                     ("distlt5", lambda: ops.distlt(5.0, B)), ("flt2", lambda: ops.flt(2.0, B)),
                     ("volume", lambda: ops.volume(B))):
        s = timed(fn)
        res[name] = {"ms": round(1e3 * s, 3), "gvox_per_s": round(n ** 3 / s / 1e9, 2)}
    if rank == 0:
        print(json.dumps({"slab_bench": {"size": n, "n_gpus": world, "voxels": n ** 3, "ops": res,
                                         "timing": "host clock around device-synchronised regions, max over ranks"}}),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
