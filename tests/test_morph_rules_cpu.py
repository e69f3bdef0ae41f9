"""The byte-lane arithmetic of morph16_kernel (voxlogica_b200/csrc/vx_morph.cu) restated on Python
integers: 16 voxels per thread as four 32-bit words of 0/1 bytes, rows y-1,y,y+1 and planes
z-1,z,z+1 combined with OR (dilation) or AND (erosion, identity 1 outside the image), the
x-neighbourhood by funnel shifts with one edge byte from each neighbouring chunk.  Checked against
scipy.ndimage binary_dilation / binary_erosion with the 6- and 26-neighbourhood."""
import numpy as np
import pytest
from scipy import ndimage as ndi

M32 = 0xffffffff


def words_of(row16):
    return [int.from_bytes(bytes(row16[4 * q:4 * q + 4].tolist()), "little") for q in range(4)]


def morph_chunked(vol, full, erode):
    nz, ny, nx = vol.shape
    assert nx % 16 == 0
    ident_w, ident_b = (0x01010101, 1) if erode else (0, 0)
    comb = (lambda a, b: a & b) if erode else (lambda a, b: a | b)

    def chunk(z, y, xc):                       # four words of the chunk, identity outside the image
        if 0 <= z < nz and 0 <= y < ny:
            return words_of(vol[z, y, 16 * xc:16 * xc + 16])
        return [ident_w] * 4

    def edge(z, y, x):                         # one voxel, identity outside
        return int(vol[z, y, x]) if (0 <= z < nz and 0 <= y < ny and 0 <= x < nx) else ident_b

    out = np.zeros_like(vol)
    for z in range(nz):
        for y in range(ny):
            for xc in range(nx // 16):
                def plane(zz):                 # FULL: rows combined; FACE: centre row and side rows apart
                    c, u, d = chunk(zz, y, xc), chunk(zz, y - 1, xc), chunk(zz, y + 1, xc)
                    lb = [edge(zz, yy, 16 * xc - 1) for yy in (y, y - 1, y + 1)]
                    rb = [edge(zz, yy, 16 * xc + 16) for yy in (y, y - 1, y + 1)]
                    if full:
                        return ([comb(comb(a, b), e) for a, b, e in zip(c, u, d)], None,
                                comb(comb(lb[0], lb[1]), lb[2]), comb(comb(rb[0], rb[1]), rb[2]))
                    return c, [comb(a, b) for a, b in zip(u, d)], lb[0], rb[0]
                pc, ps, plb, prb = plane(z - 1)
                cc, cs, clb, crb = plane(z)
                nc, ns, nlb, nrb = plane(z + 1)
                if full:
                    C = [comb(comb(a, b), e) for a, b, e in zip(pc, cc, nc)]
                    L, R, E = comb(comb(plb, clb), nlb), comb(comb(prb, crb), nrb), None
                else:
                    C, L, R = cc, clb, crb
                    E = [comb(comb(a, b), e) for a, b, e in zip(pc, nc, cs)]
                fl = lambda lo, hi: ((hi << 8) | (lo >> 24)) & M32      # __funnelshift_l(lo, hi, 8)
                fr = lambda lo, hi: ((lo >> 8) | (hi << 24)) & M32      # __funnelshift_r(lo, hi, 8)
                r = [comb(comb(C[0], ((C[0] << 8) | L) & M32), fr(C[0], C[1])),
                     comb(comb(C[1], fl(C[0], C[1])), fr(C[1], C[2])),
                     comb(comb(C[2], fl(C[1], C[2])), fr(C[2], C[3])),
                     comb(comb(C[3], fl(C[2], C[3])), ((C[3] >> 8) | (R << 24)) & M32)]
                if not full:
                    r = [comb(a, b) for a, b in zip(r, E)]
                out[z, y, 16 * xc:16 * xc + 16] = np.frombuffer(b"".join(w.to_bytes(4, "little") for w in r), np.uint8)
    return out


@pytest.mark.parametrize("shape", [(3, 4, 16), (4, 5, 32), (1, 6, 48), (5, 1, 32)])
@pytest.mark.parametrize("full", [True, False])
@pytest.mark.parametrize("erode", [False, True])
def test_byte_lane_morphology_equals_scipy(shape, full, erode):
    rng = np.random.default_rng(sum(shape) + 2 * full + erode)
    vol = (rng.random(shape) < (0.75 if erode else 0.2)).astype(np.uint8)
    st = np.ones((3, 3, 3), bool) if full else ndi.generate_binary_structure(3, 1)
    if erode:
        want = ndi.binary_erosion(vol, structure=st, border_value=1)      # outside never erodes
    else:
        want = ndi.binary_dilation(vol, structure=st, border_value=0)
    assert np.array_equal(morph_chunked(vol, full, erode), want.astype(np.uint8))
