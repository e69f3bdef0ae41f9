"""Per-operator microbenchmark: Gvoxel/s and achieved GB/s (algorithmic bytes) against the
measured HBM peak, per SURVEY.md section 8(d).  Algorithmic bytes count a boolean voxel as ONE BIT
(round 2: boolean images live as bits on the device), a number as 4 bytes, a 16-bit code as 2.  Every operator is timed with CUDA events on
the library's stream, after warm-up, with an L2 flush before each timed call.

    python scripts/bench_ops.py [--shape brats|512|256] [--batch B] [--reps R] [--only name,name]

Writes one JSON object per operator to stdout (and all of them to --out if given)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from voxlogica_b200 import Context  # noqa: E402
from voxlogica_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="brats")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    shape = {"brats": (155, 240, 240), "512": (512, 512, 512), "256": (256, 256, 256),
             "small": (48, 80, 96)}[args.shape]
    B = args.batch
    nvox = shape[0] * shape[1] * shape[2]
    total = nvox * B
    peak = 6551.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    ctx = Context(0)
    rng = np.random.default_rng(7)

    def rep(a):
        return np.stack([np.roll(a, i * 3, axis=2) for i in range(B)])

    # inputs ---------------------------------------------------------------------------
    flair_h = rep(synth.brats_like(1234, shape))
    flair = ctx.upload(flair_h)
    # the same volumes as a 16-bit file holds them (the bench's input path): the image keeps its codes
    flair16 = ctx.upload_scaled(np.rint(flair_h * np.float32(synth.INT16_SCALE)).astype(np.int16), synth.INT16_SLOPE)
    noise = ctx.upload(rng.random((B,) + shape, dtype=np.float32))
    zeros = ctx.upload(np.zeros((B,) + shape, np.float32))
    bg = ctx.cmp_scalar(flair, "<", 0.1)                      # dense mask (~75% set, one giant component)
    bern = {p: ctx.upload(rep(synth.bernoulli_volume(shape, p))) for p in (0.05, 0.3, 0.6)}
    blobs = ctx.upload(rep(synth.blob_volume(shape)))
    sparse = ctx.upload(rep(synth.bernoulli_volume(shape, 1e-4)))
    seeds = ctx.upload(rep(synth.bernoulli_volume(shape, 1e-3, seed=11)))
    brain = ctx.not_(ctx.and_(bg, ctx.through(ctx.near(ctx.border(bg)), bg)))
    lesion = ctx.cmp_scalar(flair, ">", 0.75)
    mn, mx = ctx.min(flair), ctx.max(flair)
    mn16, mx16 = ctx.min(flair16), ctx.max(flair16)
    ctx.sync()

    # (name, callable, algorithmic bytes per voxel)
    ops = [
        ("threshold f32->bool", lambda: ctx.cmp_scalar(flair, "<", 0.1), 4.125),
        ("not", lambda: ctx.not_(bg), 0.25),
        ("and", lambda: ctx.and_(bg, blobs), 0.375),
        ("arith f32*s", lambda: ctx.arith_scalar(flair, "*", 2.0), 8),
        ("volume", lambda: ctx.volume(bg), 0.125),
        ("max f32", lambda: ctx.max(flair), 4),
        ("border", lambda: ctx.border(bg), 0.125),
        ("near26", lambda: ctx.near(blobs, 26), 0.25),
        ("near6", lambda: ctx.near(blobs, 6), 0.25),
        ("interior26", lambda: ctx.interior(blobs, 26), 0.25),
        ("through26 bg-mask", lambda: ctx.through(seeds, bg, 26), 0.375),
        ("through26 blobs", lambda: ctx.through(seeds, blobs, 26), 0.375),
        ("through6 blobs", lambda: ctx.through(seeds, blobs, 6), 0.375),
        ("ccl26 bernoulli .05", lambda: ctx.components(bern[0.05], 26), 4.125),
        ("ccl26 bernoulli .3", lambda: ctx.components(bern[0.3], 26), 4.125),
        ("ccl26 bernoulli .6", lambda: ctx.components(bern[0.6], 26), 4.125),
        ("ccl26 blobs", lambda: ctx.components(blobs, 26), 4.125),
        ("maxvol26 blobs", lambda: ctx.maxvol(blobs, 26), 0.25),
        ("edt blobs", lambda: ctx.edt(blobs), 4.125),
        ("edt bernoulli .05", lambda: ctx.edt(bern[0.05]), 4.125),
        ("edt sparse 1e-4", lambda: ctx.edt(sparse), 4.125),
        ("edt lesion", lambda: ctx.edt(lesion), 4.125),
        ("distlt r=5 blobs", lambda: ctx.distlt(5.0, blobs), 0.25),
        ("distgeq r=2 blobs", lambda: ctx.distgeq(2.0, blobs), 0.25),
        ("distlt r=12 blobs", lambda: ctx.distlt(12.0, blobs), 0.25),
        ("percentiles brain f32 (sort)", lambda: ctx.percentiles(flair, brain), 8.125),
        # the counting path returns a DEFERRED image (tables only); device_ptr asks for the f32 array
        ("percentiles brain 16-bit upload (counting)", lambda: ctx.device_ptr(ctx.percentiles(flair16, brain)), 6.125),
        ("percentiles brain 16-bit upload (tables only)", lambda: ctx.percentiles(flair16, brain), 2.125),
        ("percentiles then cmp 16-bit upload (deferred)", lambda: ctx.cmp_scalar(ctx.percentiles(flair16, brain), ">", 0.95), 2.125 + 2.25),
        ("percentiles then cmp f32 (sort)", lambda: ctx.cmp_scalar(ctx.percentiles(flair, brain), ">", 0.95), 8.125 + 4.125),
        ("percentiles all f32 noise (sort)", lambda: ctx.percentiles(noise, ctx.tt(bg)), 8.125),
        ("minmax f32", lambda: ctx.minmax(flair), 4),
        ("minmax 16-bit upload (code range)", lambda: ctx.minmax(flair16), 0),
        ("xcorr rho5 k100 flair", lambda: ctx.cross_correlation(5.0, flair, flair, lesion, mn, mx, 100), 13),
        ("xcorr rho5 k100 16-bit upload", lambda: ctx.cross_correlation(5.0, flair16, flair16, lesion, mn16, mx16, 100), 13),
        ("xcorr rho5 k100 noise", lambda: ctx.cross_correlation(5.0, noise, noise, lesion, 0.0, 1.0, 100), 13),
        ("xcorr rho5 k100 zeros", lambda: ctx.cross_correlation(5.0, zeros, zeros, lesion, 0.0, 1.0, 100), 13),
        ("xcorr rho2 k100 flair", lambda: ctx.cross_correlation(2.0, flair, flair, lesion, mn, mx, 100), 13),
    ]
    only = [s.replace("_", " ") for s in args.only.split(",") if s]      # "_" stands for a blank: shell-friendly
    results = []
    for name, fn, bpv in ops:
        if only and not any(o in name for o in only):
            continue
        for _ in range(2):
            fn()
        ctx.sync()
        ms_list, kern = [], {}
        for _ in range(args.reps):
            ctx.flush_l2()
            ctx.prof_reset(); ctx.prof_enable(True)
            e0, e1 = ctx.event(), ctx.event()
            ctx.record(e0)
            r = fn()
            ctx.record(e1)
            ms_list.append(ctx.elapsed_ms(e0, e1))
            ctx.prof_enable(False)
            for k, (c, m) in ctx.prof_report().items():
                kern.setdefault(k, []).append(m)
            del r
        ms = float(np.median(ms_list))
        gbs = bpv * total / (ms * 1e-3) / 1e9
        rec = {"op": name, "shape": list(shape), "batch": B, "ms": round(ms, 4),
               "gvox_per_s": round(total / (ms * 1e-3) / 1e9, 2), "alg_bytes_per_voxel": bpv,
               "achieved_gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4),
               "kernels_ms": {k: round(float(np.median(v)), 4) for k, v in sorted(kern.items(), key=lambda kv: -np.median(kv[1]))}}
        results.append(rec)
        print(json.dumps(rec), flush=True)
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
