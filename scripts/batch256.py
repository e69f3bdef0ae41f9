"""BASELINE config 5a: a fixed batch of 256 BraTS-shaped volumes sharded across N GPUs (STRONG
scaling: total work fixed).  Volume i goes to rank i mod N; every rank streams its share through
pinned host memory in batches (H2D, GTV spec, D2H of the u8 masks), two host threads per rank
so copies overlap kernels.  No data-path collective; one all-reduce(MAX) of the elapsed time.

    python scripts/batch256.py                       # 1 GPU
    torchrun --nproc-per-node 8 scripts/batch256.py  # 8 GPUs
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--volumes", type=int, default=256)
    ap.add_argument("--batch", type=int, default=8, help="volumes per upload / per host thread")
    args = ap.parse_args()
    import torch
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from voxlogica_b200 import Context
    from voxlogica_b200.imgql import run_gtv
    ctx = Context(local)
    mine = list(range(rank, args.volumes, world))            # volume ids of this rank
    B = args.batch
    nbatches = (len(mine) + B - 1) // B
    host = bench.make_batch(B, rank)                        # B distinct volumes, reused per batch (documented)
    nthr = 2
    pins_in = [ctx.host_alloc(B * bench.NVOX * 4) for _ in range(nthr)]
    pins_out = [ctx.host_alloc(B * bench.NVOX) for _ in range(nthr)]
    for t in range(nthr):
        C.memmove(pins_in[t], host.ctypes.data, host.nbytes)
    vox = [0.0] * nthr

    def worker(t, batches, errors):
        try:
            for b in batches:
                n = min(B, len(mine) - b * B)
                f = ctx.upload_ptr(pins_in[t], np.float32, (n,) + bench.BRATS)
                out = run_gtv(ctx, f)
                ctx.download_ptr(out["gtv"], pins_out[t], n * bench.NVOX)
                vox[t] += float(np.sum(out["_printed"]["gtvVolume"]))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    def run(batches_per_thread):
        errors = []
        ths = [threading.Thread(target=worker, args=(t, batches_per_thread[t], errors)) for t in range(nthr)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        if errors:
            raise errors[0]
        ctx.sync()

    run([[0], [0]])                                          # warm-up: allocator, clocks
    vox[:] = [0.0] * nthr
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    run([list(range(t, nbatches, nthr)) for t in range(nthr)])
    sec = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([sec], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt[0])
        tv = torch.tensor([sum(vox)], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tv)
        total_vox = float(tv[0])
    else:
        total_vox = sum(vox)
    if rank == 0:
        print(json.dumps({"config": "batch of 256 BraTS-shaped volumes, GTV spec, end to end from pinned host memory",
                          "volumes": args.volumes, "n_gpus": world, "seconds": round(sec, 4),
                          "volumes_per_s": round(args.volumes / sec, 1), "scaling": "strong",
                          "gtv_voxels_total": total_vox, "batch_per_upload": B, "host_threads_per_rank": nthr,
                          "data_note": "each rank cycles min(batch,4) generated volumes + axis flips"}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
