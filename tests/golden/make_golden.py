"""Generate tests/golden/golden_v1.npz: seeded inputs with (i) the CPU oracle's outputs and
(ii) INDEPENDENT outputs of scipy.ndimage / OpenCV for the same inputs, where such a
counterpart exists.

    python tests/golden/make_golden.py        # rewrites golden_v1.npz

Why this and not reference vectors: /root/reference holds no sources, tests or fixtures and
SimpleITK is not installed, so the reference cannot be run to emit
vectors -- PARITY UNPINNED.  The fixtures pin (a) the oracle against regressions, (b) the oracle
against an independent implementation, (c) the CUDA path against both (tests/test_golden.py).
Boolean volumes are stored bit-packed; everything is deterministic in the seeds below.
"""
import os
import sys

import numpy as np
from scipy import ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle as orc  # noqa: E402
from gtv_ref import gtv_oracle  # noqa: E402
from voxlogica_b200 import synth  # noqa: E402

S6 = ndi.generate_binary_structure(3, 1)
S26 = ndi.generate_binary_structure(3, 3)
BIN_SHAPE = (20, 24, 28)
GTV_SHAPE = (32, 56, 64)
SPACING = (0.5, 1.25, 3.0)   # (sx, sy, sz)


def pack(a):
    return np.packbits(np.asarray(a, dtype=np.uint8).ravel())


def canon_labels(lab):
    """relabel so that each component carries 1 + the raster index of its first voxel"""
    flat = lab.ravel()
    out = np.zeros_like(flat, dtype=np.uint32)
    idx = np.flatnonzero(flat)
    first = {}
    for i in idx:
        first.setdefault(flat[i], i)
    lut = {k: v + 1 for k, v in first.items()}
    out[idx] = [lut[v] for v in flat[idx]]
    return out.reshape(lab.shape)


def main():
    g = {}
    rng = np.random.default_rng(20261020)
    # ---------------------------------------------------------------- binary volumes
    a = (rng.random(BIN_SHAPE) < 0.3).astype(np.uint8)
    seeds = (rng.random(BIN_SHAPE) < 0.01).astype(np.uint8)
    sparse = (rng.random(BIN_SHAPE) < 0.004).astype(np.uint8)
    g["bin.a"], g["bin.seeds"], g["bin.sparse"] = pack(a), pack(seeds), pack(sparse)
    for conn, st in ((6, S6), (26, S26)):
        g[f"bin.near{conn}.oracle"] = pack(orc.near(a, conn))
        g[f"bin.near{conn}.scipy"] = pack(ndi.binary_dilation(a, st))
        g[f"bin.interior{conn}.oracle"] = pack(orc.interior(a, conn))
        g[f"bin.interior{conn}.scipy"] = pack(ndi.binary_erosion(a, st, border_value=1))
        lab, n = ndi.label(a, st)
        g[f"bin.components{conn}.oracle"] = orc.components(a, conn).astype(np.uint32)
        g[f"bin.components{conn}.scipy"] = canon_labels(lab)
        keep = np.unique(lab[(seeds != 0) & (a != 0)])
        g[f"bin.through{conn}.oracle"] = pack(orc.through(seeds, a, conn))
        g[f"bin.through{conn}.scipy"] = pack(np.isin(lab, keep[keep > 0]))
        sizes = ndi.sum(a, lab, index=np.arange(1, n + 1))
        g[f"bin.maxvol{conn}.oracle"] = pack(orc.maxvol(a, conn))
        g[f"bin.maxvol{conn}.scipy"] = pack(np.isin(lab, 1 + np.flatnonzero(sizes == sizes.max())))
    for name, src in (("a", a), ("sparse", sparse)):
        for tag, sp in (("unit", (1.0, 1.0, 1.0)), ("aniso", SPACING)):
            g[f"bin.edt.{name}.{tag}.oracle"] = orc.edt(src, sp)
            g[f"bin.edt.{name}.{tag}.scipy"] = ndi.distance_transform_edt(src == 0, sampling=sp[::-1]).astype(np.float64)
    for r in (1.0, 2.0, 2.5, 5.0):
        g[f"bin.distlt{r}.oracle"] = pack(orc.distlt(r, sparse))
        g[f"bin.distgeq{r}.oracle"] = pack(orc.distgeq(r, sparse))
        g[f"bin.flt{r}.oracle"] = pack(orc.flt(r, orc.near(orc.near(sparse))))
    # ---------------------------------------------------------------- numeric volume
    f = rng.random(BIN_SHAPE, dtype=np.float32)
    f[rng.random(BIN_SHAPE) < 0.2] = 0.5        # ties
    m = (rng.random(BIN_SHAPE) < 0.6).astype(np.uint8)
    g["num.f"], g["num.mask"] = f, pack(m)
    for corr in (0.0, 0.5):
        g[f"num.percentiles{corr}.oracle"] = orc.percentiles(f, m, corr)
    vals = np.sort(f[m != 0])
    rank = np.searchsorted(vals, f, side="left").astype(np.float64) / vals.size
    g["num.percentiles0.0.numpy"] = np.where(m != 0, rank, 0.0).astype(np.float32)
    for rho, k in ((2.0, 16), (3.0, 100)):
        g[f"num.xcorr.rho{rho}.k{k}.oracle"] = orc.cross_correlation(rho, f, f, m, 0.0, 1.0, k)
    g["num.xcorr.aniso.oracle"] = orc.cross_correlation(3.0, f, f, m, 0.1, 0.9, 32, SPACING)
    g["num.scalars.oracle"] = np.array([orc.volume(m), orc.min_(f), orc.max_(f), orc.sum_(f),
                                        orc.min_(f, m), orc.max_(f, m)], dtype=np.float64)
    g["num.lt0.4.oracle"] = pack(orc.cmp_scalar(f, "<", 0.4))
    g["num.mul.oracle"] = orc.arith_scalar(f, "*", 1.7)
    # ---------------------------------------------------------------- GTV spec (config 2, small)
    vol = synth.brats_like(seed=5, shape=GTV_SHAPE)
    r = gtv_oracle(orc, vol)
    g["gtv.seed"] = np.array([5], dtype=np.int64)
    for name in ("background", "brain", "hyperIntense", "veryIntense", "growTum", "tumStatCC", "gtv"):
        g[f"gtv.{name}.oracle"] = pack(r[name])
    g["gtv.normFlair.oracle"] = r["normFlair"]
    g["gtv.tumSim.oracle"] = r["tumSim"]
    g["gtv.flair.sha"] = np.frombuffer(__import__("hashlib").sha256(vol.tobytes()).digest(), dtype=np.uint8)
    # ---------------------------------------------------------------- maze (config 1)
    img = synth.maze(24, seed=0)
    path = (img.sum(axis=2) > 0).astype(np.uint8)
    start = np.zeros_like(path); start[1, 1] = 1
    g["maze.rgb"] = img
    for conn in (6, 26):
        g[f"maze.fromStart{conn}.oracle"] = pack(orc.through(start, path, conn))
    lab4, _ = ndi.label(path, ndi.generate_binary_structure(2, 1))
    g["maze.fromStart6.scipy"] = pack(lab4 == lab4[1, 1])
    out = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
