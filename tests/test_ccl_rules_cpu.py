"""The union rules of the CUDA connected-component merge (voxlogica_b200/csrc/vx_ccl.cu), restated
word for word on Python integers and checked against scipy.ndimage.label on random volumes.

This is the correctness argument of the kernel as a CPU test: one union per pair of adjacent
x-runs, placed where the later of the two runs starts (`first` / `second` masks), anchors chained
along a row by the init pass, and the diagonal directions (y-1,z-1) / (y+1,z-1) skipped wherever a
set voxel of a face row already connects the two runs.  If any of these rules dropped a needed
union, some component would come out split."""
import numpy as np
import pytest
from scipy import ndimage as ndi

M32 = 0xffffffff


def seg_start(m, j):
    zeros_below = ~m & ((1 << j) - 1)
    return zeros_below.bit_length() if zeros_below else 0


def components_by_kernel_rules(vol, full):
    nz, ny, nx = vol.shape
    nwx = (nx + 31) // 32
    bits = np.zeros((nz, ny, nwx), dtype=np.uint64)
    for x in range(nx):
        bits[:, :, x // 32] |= vol[:, :, x].astype(np.uint64) << np.uint64(x % 32)
    parent = {}

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    def unite(a, b):
        a, b = find(a), find(b)
        if a != b:
            parent[max(a, b)] = min(a, b)

    def word(z, y, w):
        if z < 0 or z >= nz or y < 0 or y >= ny or w < 0 or w >= nwx:
            return 0
        return int(bits[z, y, w])

    # init: run starts are roots; bit 0 of a run continuing from the previous word hangs under
    # that word's last segment
    for z in range(nz):
        for y in range(ny):
            for w in range(nwx):
                m = word(z, y, w)
                prev_m = word(z, y, w - 1)
                lin0 = (z * ny + y) * nx + w * 32
                starts = m & ~(((m << 1) | (prev_m >> 31)) & M32)
                for j in range(32):
                    if (starts >> j) & 1:
                        parent[lin0 + j] = lin0 + j
                if (m & 1) and (prev_m >> 31):
                    parent[lin0] = lin0 - 32 + seg_start(prev_m, 31)
    # the chain links must point at existing anchors
    assert all(p in parent for p in parent.values())

    def row_bits(z, y, w):
        nc, np_, nn = word(z, y, w), word(z, y, w - 1), (word(z, y, w + 1) if full else 0)
        return nc, ((nc << 1) | (np_ >> 31)) & M32, ((nc >> 1) | (nn << 31)) & M32, np_

    dirs = [(-1, 0), (0, -1), (-1, -1), (1, -1)] if full else [(-1, 0), (0, -1)]
    n_unions = 0
    for z in range(nz):
        for y in range(ny):
            for w in range(nwx):
                m = word(z, y, w)
                if m == 0:
                    continue
                prev_m = word(z, y, w - 1)
                left = ((m << 1) | (prev_m >> 31)) & M32
                lin0 = (z * ny + y) * nx + w * 32
                u0 = um = up = v0 = vm = vp = 0
                for d, (dy, dz) in enumerate(dirs):
                    yy, zz = y + dy, z + dz
                    if yy < 0 or yy >= ny or zz < 0:
                        continue
                    N0, Nm, Np, npw = row_bits(zz, yy, w)
                    nlin0 = lin0 + (dz * ny + dy) * nx
                    if full:
                        first = m & Np & ~N0 & M32
                        second = m & ~left & (N0 | Nm) & M32
                        if d == 0:
                            u0, um, up = N0, Nm, Np
                        if d == 1:
                            v0, vm, vp = N0, Nm, Np
                        if d >= 2:
                            c0, cm, cp = u0 | v0, um | vm, up | vp
                            if d == 3:
                                w0, wm, wp, _ = row_bits(z, y + 1, w) if y + 1 < ny else (0, 0, 0, 0)
                                c0, cm, cp = v0 | w0, vm | wm, vp | wp
                            first &= ~(c0 | cp) & M32
                            second &= ~(c0 | (~N0 & cm)) & M32
                    else:
                        first = 0
                        second = m & N0 & ~(left & Nm) & M32
                    for j in range(32):
                        if (first >> j) & 1:
                            unite(lin0 + seg_start(m, j), nlin0 + j + 1)
                            n_unions += 1
                    for j in range(32):
                        if (second >> j) & 1:
                            if (N0 >> j) & 1:
                                b = nlin0 + seg_start(N0, j)
                            elif j > 0:
                                b = nlin0 + seg_start(N0, j - 1)
                            else:
                                b = nlin0 - 32 + seg_start(npw, 31)
                            unite(lin0 + seg_start(m, j), b)
                            n_unions += 1
    # label of a voxel = root of the anchor of its run segment inside its word
    lab = np.zeros(vol.shape, np.int64)
    for z in range(nz):
        for y in range(ny):
            for w in range(nwx):
                m = word(z, y, w)
                lin0 = (z * ny + y) * nx + w * 32
                for j in range(min(32, nx - w * 32)):
                    if (m >> j) & 1:
                        lab[z, y, w * 32 + j] = find(lin0 + seg_start(m, j)) + 1
    return lab, n_unions


def same_partition(a, b):
    fg = a > 0
    if not np.array_equal(fg, b > 0):
        return False
    pairs = np.unique(np.stack([a[fg], b[fg]], axis=1), axis=0)
    return len(np.unique(pairs[:, 0])) == len(pairs) == len(np.unique(pairs[:, 1]))


@pytest.mark.parametrize("shape", [(4, 5, 33), (3, 6, 70), (5, 4, 64), (2, 3, 5), (6, 7, 40)])
@pytest.mark.parametrize("p", [0.15, 0.45, 0.8])
@pytest.mark.parametrize("full", [True, False])
def test_union_rules_give_scipy_components(shape, p, full):
    rng = np.random.default_rng(hash((shape, p, full)) % (2 ** 32))
    vol = (rng.random(shape) < p).astype(np.uint8)
    lab, _ = components_by_kernel_rules(vol, full)
    structure = np.ones((3, 3, 3), int) if full else ndi.generate_binary_structure(3, 1)
    want, _ = ndi.label(vol, structure=structure)
    assert same_partition(lab, want)


def test_diagonal_skip_removes_unions_on_solid_blocks():
    """on a solid block every diagonal contact is covered by a face row: as many unions as with
    faces only, far fewer than four per run"""
    vol = np.ones((6, 8, 48), np.uint8)
    _, n26 = components_by_kernel_rules(vol, True)
    _, n6 = components_by_kernel_rules(vol, False)
    runs = 6 * 8 * 2          # two words per row
    assert n26 == n6 and n26 < 2 * runs
