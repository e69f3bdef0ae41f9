"""Synthetic inputs of SURVEY.md section 8(d): BraTS-shaped FLAIR-like volumes, a DFS maze
PNG-like image and the 512^3-style binary microbenchmark volumes.  Host-side numpy only;
used by bench.py, the tests and smoke().  (BraTS data itself is not available here.)"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

BRATS_SHAPE = (155, 240, 240)  # [nz][ny][nx]  = 8,928,000 voxels (Greta.pdf p.48 "9 M voxels")


def _smooth_noise(rng, shape, coarse=12):
    """band-limited noise in [0,1]: coarse grid, trilinear zoom (fast at full BraTS size)"""
    cs = tuple(max(2, int(np.ceil(s / coarse)) + 1) for s in shape)
    c = rng.random(cs).astype(np.float32)
    z = ndi.zoom(c, [s / k for s, k in zip(shape, cs)], order=1, mode="nearest", grid_mode=True)
    z = z[: shape[0], : shape[1], : shape[2]]
    pad = [(0, s - t) for s, t in zip(shape, z.shape)]
    if any(p[1] for p in pad):
        z = np.pad(z, pad, mode="edge")
    return z


def brats_like(seed: int = 1234, shape=BRATS_SHAPE) -> np.ndarray:
    """f32 [nz][ny][nx] in [0,1]: ellipsoidal head of smooth low-frequency tissue noise in
    [0.15,0.6], background exactly 0 with a thin dim rim (< 0.1) around the head, 1-3
    ellipsoidal hyper-intense lesions (0.8-1.0) each with a 0.65-0.8 halo, Gaussian noise
    sigma 0.02 inside the head, clipped to [0,1]  (SURVEY.md section 8d, config 2)."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.arange(nz, dtype=np.float32), np.arange(ny, dtype=np.float32),
                          np.arange(nx, dtype=np.float32), indexing="ij", sparse=True)
    cz, cy, cx = (nz - 1) / 2, (ny - 1) / 2, (nx - 1) / 2
    az, ay, ax = 65 / 155 * nz, 100 / 240 * ny, 85 / 240 * nx        # semi-axes
    rr = ((z - cz) / az) ** 2 + ((y - cy) / ay) ** 2 + ((x - cx) / ax) ** 2
    head = rr <= 1.0
    rim = (rr > 1.0) & (rr <= 1.06)
    vol = np.zeros(shape, np.float32)
    tissue = 0.15 + 0.45 * _smooth_noise(rng, shape)
    vol[head] = tissue[head]
    scale = min(nz / 155, ny / 240, nx / 240)
    for _ in range(int(rng.integers(1, 4))):
        r = max(rng.uniform(8, 25) * scale, 6.5)   # filter(5.0, .) keeps only blobs of radius > 5
        # keep lesions inside the head
        lz = cz + rng.uniform(-0.45, 0.45) * az
        ly = cy + rng.uniform(-0.45, 0.45) * ay
        lx = cx + rng.uniform(-0.45, 0.45) * ax
        ez, ey, ex = r * rng.uniform(0.7, 1.0), r * rng.uniform(0.7, 1.0), r * rng.uniform(0.7, 1.0)
        d = ((z - lz) / ez) ** 2 + ((y - ly) / ey) ** 2 + ((x - lx) / ex) ** 2
        core = d <= 1.0
        halo = (d > 1.0) & (d <= 1.7) & head
        vol[halo] = np.maximum(vol[halo], rng.uniform(0.65, 0.8))
        vol[core & head] = rng.uniform(0.8, 1.0)
    noise = rng.normal(0.0, 0.02, size=shape).astype(np.float32)
    vol = np.where(head, vol + noise, vol)
    vol[rim] = rng.uniform(0.0, 0.09, size=int(rim.sum())).astype(np.float32)
    return np.clip(vol, 0.0, 1.0).astype(np.float32)


INT16_SCALE = 4095          # 12-bit acquisition range, as stored in a 16-bit NIfTI file
INT16_SLOPE = np.float32(1.0 / INT16_SCALE)


def brats_like_int16(seed: int = 1234, shape=BRATS_SHAPE):
    """The same synthetic volume as a scanner / NIfTI file holds it: int16 voxels in [0, 4095] plus
    the header's scl_slope (1/4095, scl_inter 0) that maps them back to [0, 1].  Returns
    (raw int16 [nz][ny][nx], slope); the numeric image every arm works on is
    ``raw.astype(f32) * slope`` (int16_to_f32), formed on the host or -- bit-identically -- by
    vx_upload_scaled on the device."""
    v = brats_like(seed, shape)
    return np.rint(v * np.float32(INT16_SCALE)).astype(np.int16), INT16_SLOPE


def int16_to_f32(raw: np.ndarray, slope=INT16_SLOPE, inter=0.0) -> np.ndarray:
    """host-side twin of vx_upload_scaled: two binary32 roundings, multiply then add"""
    return raw.astype(np.float32) * np.float32(slope) + np.float32(inter)


MODALITIES = ("flair", "t1", "t1c", "t2")


def brats_multimodal(seed: int = 1234, shape=BRATS_SHAPE) -> dict:
    """Four co-registered modalities of one synthetic subject (SURVEY.md section 8d, config 4:
    seeds seed..seed+3, shared anatomy).  `flair` is brats_like(seed); the others keep its head,
    rim and lesion geometry (they are derived from the noise-free part of its intensities) but
    have their own tissue contrast, lesion contrast and noise -- like T1 / T1c / T2 of a
    BraTS case, where the lesion is dark, ring-bright and bright respectively."""
    flair = brats_like(seed, shape)
    head = flair > 0.0
    lesion = ndi.uniform_filter(flair, 3) > 0.72          # core + halo of the lesions, denoised
    out = {"flair": flair}
    contrast = {"t1": (0.55, -0.35, 0.25), "t1c": (0.50, -0.20, 0.75), "t2": (0.20, 0.55, 0.85)}
    for i, name in enumerate(MODALITIES[1:], start=1):
        rng = np.random.default_rng(seed + i)
        base, slope, les = contrast[name]
        tissue = base + slope * (flair - 0.375) + 0.10 * (_smooth_noise(rng, shape, coarse=16) - 0.5)
        vol = np.where(lesion, np.float32(les) + 0.05 * (_smooth_noise(rng, shape, coarse=8) - 0.5), tissue)
        vol = vol + rng.normal(0.0, 0.02, size=shape).astype(np.float32)
        out[name] = np.where(head, np.clip(vol, 0.01, 1.0), 0.0).astype(np.float32)
    return out


def brats_batch(seeds, shape=BRATS_SHAPE) -> np.ndarray:
    return np.stack([brats_like(int(s), shape) for s in seeds])


def maze(n_cells: int = 512, seed: int = 0):
    """Perfect maze by randomised DFS on n_cells x n_cells cells -> RGB uint8 image of
    (2n+1)^2 pixels: walls black, corridors white, start cell blue, exit cell green
    (the reference's maze PNG is not in the mount)."""
    rng = np.random.default_rng(seed)
    n = n_cells
    img = np.zeros((2 * n + 1, 2 * n + 1, 3), np.uint8)
    seen = np.zeros((n, n), bool)
    stack = [(0, 0)]
    seen[0, 0] = True
    img[1, 1] = 255
    dirs = [(0, 1), (1, 0), (0, -1), (-1, 0)]
    while stack:
        cy, cx = stack[-1]
        order = rng.permutation(4)
        for k in order:
            dy, dx = dirs[k]
            yy, xx = cy + dy, cx + dx
            if 0 <= yy < n and 0 <= xx < n and not seen[yy, xx]:
                seen[yy, xx] = True
                img[2 * cy + 1 + dy, 2 * cx + 1 + dx] = 255
                img[2 * yy + 1, 2 * xx + 1] = 255
                stack.append((yy, xx))
                break
        else:
            stack.pop()
    img[1, 1] = (0, 0, 255)
    img[2 * n - 1, 2 * n - 1] = (0, 255, 0)
    return img


def bernoulli_volume(shape, p: float, seed: int = 7) -> np.ndarray:
    return (np.random.default_rng(seed).random(shape, dtype=np.float32) < p).astype(np.uint8)


def blob_volume(shape, seed: int = 7, coarse: int = 24, q: float = 0.6) -> np.ndarray:
    s = _smooth_noise(np.random.default_rng(seed), shape, coarse)
    return (s > np.quantile(s[::4, ::4, ::4], q)).astype(np.uint8)


def task_dag_spec(n_tasks: int, seed: int = 0, window: int = 6) -> str:
    """ImgQL text of a generated formula that evaluates to exactly ``n_tasks`` operator tasks on
    the image loaded as "vol" (Greta.pdf p.79/Image420 times formulas of 11 ... 8195 tasks; the
    formulas themselves are not published, so this is a seeded stand-in): every `let` applies ONE
    operator to results of the last ``window`` statements of the right type, boolean and numeric
    results alternate through thresholds and `mask`, and no (operator, operands) pair repeats, so
    memoisation cannot shrink the task count.  Tasks = intensity + the lets + the final volume."""
    if n_tasks < 4:
        raise ValueError("need at least 4 tasks")
    rng = np.random.default_rng(seed)
    out = ['load vol = "vol"', "let f0 = intensity(vol)", "let b0 = f0 >. 0.5"]
    fs, bs, seen = ["f0"], ["b0"], {"f0 >. 0.5"}
    # an operand is one of the last `window` results or, one time in four, the input itself (a formula
    # built only from its own recent results decays to the empty or the full image within a few dozen steps)
    pick = lambda pool: pool[0] if rng.random() < 0.25 else pool[-1 - int(rng.integers(0, min(window, len(pool))))]
    for i in range(1, n_tasks - 2):
        while True:
            u = rng.random()
            a, b, fa, fb = pick(bs), pick(bs), pick(fs), pick(fs)
            c = round(float(rng.uniform(0.05, 0.95)), 3)
            if u < 0.14:   term, kind = f"near({a})", "b"
            elif u < 0.24: term, kind = f"interior({a})", "b"
            elif u < 0.34: term, kind = f"{a} & {b}", "b"
            elif u < 0.46: term, kind = f"{a} | {b}", "b"
            elif u < 0.52: term, kind = f"!{a}", "b"
            elif u < 0.55: term, kind = f"distlt(1.5, {a})", "b"
            elif u < 0.57: term, kind = f"through({a}, {b})", "b"
            elif u < 0.69: term, kind = f"{fa} >. {c}", "b"
            elif u < 0.77: term, kind = f"{fa} + {fb}", "f"
            elif u < 0.83: term, kind = f"{fa} * {fb}", "f"
            elif u < 0.91: term, kind = f"{fa} *. {c}", "f"
            else:          term, kind = f"mask({fa}, {a})", "f"
            if term not in seen and not (("&" in term or "|" in term or "through" in term) and a == b):
                break
        seen.add(term)
        name = f"{kind}{i}"
        out.append(f"let {name} = {term}")
        (bs if kind == "b" else fs).append(name)
    out.append(f'print "volume" volume({bs[-1]})')
    return "\n".join(out) + "\n"
