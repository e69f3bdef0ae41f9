"""The fused distance-interval test of voxlogica_b200/csrc/vx_edt.cu restated in numpy:
    dist(x, A) < r   <=>   g3(x, A) < T(r)   <=>   x lies in the dilation of A by the discrete ball
    { offset : g3(offset) < T(r) },   T(r) = the smallest binary32 whose sqrtf reaches r,
with g3 the squared distance accumulated in binary32, one rounding per step.  Checked against the
oracle's distance operators (which evaluate distances, not balls) for strict and non-strict
comparisons, anisotropic spacings and radii that sit exactly on lattice distances."""
import numpy as np
import pytest


def f32(x):
    return np.float32(x)


def threshold_for(r, strict):
    r = f32(r)
    ok = (lambda t: np.sqrt(f32(t)) > r) if strict else (lambda t: np.sqrt(f32(t)) >= r)
    if ok(0.0):
        return f32(0.0)
    t = f32(np.float64(r) * np.float64(r))
    if not t > 0:
        t = np.nextafter(f32(0), f32(1))
    while not ok(t):
        t = np.nextafter(t, f32(np.inf))
    while t > 0 and ok(np.nextafter(t, f32(0))):
        t = np.nextafter(t, f32(0))
    return t


def g3(dx, dy, dz, sp):
    sx, sy, sz = (f32(s) for s in sp)
    a = f32(f32(dx) * sx); a = f32(a * a)
    b = f32(f32(dy) * sy); b = f32(b * b)
    c = f32(f32(dz) * sz); c = f32(c * c)
    return f32(f32(a + b) + c)


def ball_dilate(a, T, sp):
    nz, ny, nx = a.shape
    out = np.zeros_like(a)
    if not T > 0:
        return out
    wx = next(d for d in range(1, nx + 2) if not g3(d, 0, 0, sp) < T or d >= nx) if nx > 1 else 0
    wy = next(d for d in range(1, ny + 2) if not g3(0, d, 0, sp) < T or d >= ny) if ny > 1 else 0
    wz = next(d for d in range(1, nz + 2) if not g3(0, 0, d, sp) < T or d >= nz) if nz > 1 else 0
    for dz in range(-wz, wz + 1):
        for dy in range(-wy, wy + 1):
            for dx in range(-wx, wx + 1):
                if not g3(dx, dy, dz, sp) < T:
                    continue
                src = a[max(0, -dz):nz - max(0, dz), max(0, -dy):ny - max(0, dy), max(0, -dx):nx - max(0, dx)]
                out[max(0, dz):nz - max(0, -dz), max(0, dy):ny - max(0, -dy), max(0, dx):nx - max(0, -dx)] |= src
    return out


@pytest.mark.parametrize("sp", [(1.0, 1.0, 1.0), (0.5, 1.0, 2.0), (0.7, 1.0, 2.5)])
@pytest.mark.parametrize("r", [0.0, 0.5, 1.0, 1.4142135, 2.0, 2.5, 3.0, 3.2])
def test_ball_dilation_equals_distance_threshold(oracle, sp, r):
    rng = np.random.default_rng(int(r * 1000) + int(sp[0] * 10))
    a = (rng.random((5, 9, 14)) < 0.03).astype(np.uint8)
    a[2, 4, 7] = 1
    lt = ball_dilate(a, threshold_for(r, strict=False), sp)        # d <  r  <=>  g3 < T(non-strict)
    le = ball_dilate(a, threshold_for(r, strict=True), sp)         # d <= r  <=>  g3 < T(strict)
    assert np.array_equal(lt, oracle.distlt(r, a, sp))
    assert np.array_equal(le, oracle.distleq(r, a, sp))
    assert np.array_equal(1 - lt, oracle.distgeq(r, a, sp))
    assert np.array_equal(1 - le, oracle.distgt(r, a, sp))
