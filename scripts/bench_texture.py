#!/usr/bin/env python
"""BASELINE config 4: the crossCorrelation texture-similarity spec (imgql.TEXTURE_SPEC) on
BraTS-shaped multi-modal synthetic subjects (flair / t1 / t1c / t2, shared anatomy).

    python scripts/bench_texture.py [--batch B] [--steps K] [--warmup W] [--shape brats|small]

A step = the whole spec over a resident batch of B subjects (4*B volumes of 240x240x155 f32:
2.3 GB at B = 16 -- larger than L2).  Timed with CUDA events on the library's stream, one host
thread.  Prints one JSON line: subjects/s, the per-kernel table of the timed steps and the
crossCorrelation kernel's share."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from voxlogica_b200 import Context, synth  # noqa: E402
from voxlogica_b200.imgql import run_texture  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--shape", default="brats")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    shape = {"brats": synth.BRATS_SHAPE, "small": (48, 80, 96)}[args.shape]
    nvox = shape[0] * shape[1] * shape[2]
    ctx = Context(0)
    uniq = [synth.brats_multimodal(1234 + 4 * i, shape) for i in range(min(args.batch, 2))]
    imgs = {}
    for m in synth.MODALITIES:
        vols = [np.ascontiguousarray(uniq[i % len(uniq)][m][:, :, ::-1] if (i // len(uniq)) & 1 else uniq[i % len(uniq)][m])
                for i in range(args.batch)]
        imgs[m] = ctx.upload(np.stack(vols))
    ctx.sync()
    vol = None
    for _ in range(max(args.warmup, 3)):
        vol = run_texture(ctx, imgs)["_printed"]["tumourVolume"]
    ctx.sync()
    ctx.prof_reset(); ctx.prof_enable(True)
    l0 = ctx.launch_count()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(args.steps):
        vol = run_texture(ctx, imgs)["_printed"]["tumourVolume"]
    ctx.record(e1)
    ctx.sync()
    ms = ctx.elapsed_ms(e0, e1)
    ctx.prof_enable(False)
    prof = ctx.prof_report()
    tot = sum(v[1] for v in prof.values()) or 1.0
    kernels = {k: {"launches_per_step": c / args.steps, "avg_launch_ms": round(m / c, 5), "share": round(m / tot, 4)}
               for k, (c, m) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    line = {"metric": "brats_texture_spec_throughput", "value": args.batch * args.steps / (ms / 1e3), "unit": "subjects/s",
            "volumes_per_s": 4 * args.batch * args.steps / (ms / 1e3), "ms_per_step": ms / args.steps,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "config": {"workload": "BASELINE config 4: crossCorrelation (rho=5, k=100 over [min,max]) of flair/t1/t1c/t2 against "
                                   "the histogram of growTum, mean coefficient > 0.6, filter, grow",
                       "subjects_per_step": args.batch, "voxels_per_volume": nvox, "modalities": list(synth.MODALITIES),
                       "l2": f"resident inputs {4 * args.batch * nvox * 4 / 1e6:.0f} MB f32"},
            "gpu_launches": ctx.launch_count() - l0,
            "xcorr_share": kernels.get("xcorr_kernel", {}).get("share"),
            "kernels": kernels, "tumour_voxels": [float(v) for v in vol][:4]}
    s = json.dumps(line)
    print(s, flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
