"""Host-side logic that needs no GPU: the ImgQL parser, the C-ABI export list, the
synthetic generators."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/voxcuda.h is the contract: every VX_API function must be exported by the
    built library and bound in _abi.SIGNATURES (dlopen only, no compute)."""
    from voxlogica_b200 import _abi
    hdr = open(os.path.join(ROOT, "include", "voxcuda.h")).read()
    declared = set(re.findall(r"VX_API\s+(?:const\s+char\*|int32_t)\s+(vx_\w+)\s*\(", hdr))
    assert len(declared) > 50
    lib = _abi.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in voxcuda.h but not exported"
    bound = set(_abi.SIGNATURES) | {"vx_version", "vx_last_error"}
    assert declared == bound, (declared ^ bound)
    assert lib.vx_version() >= 100


def test_init_without_gpu_fails_loudly():
    """No CPU fallback: without a CUDA device vx_init reports VX_E_CUDA."""
    import ctypes as C
    from voxlogica_b200 import _abi
    lib = _abi.load_library()
    import torch
    if torch.cuda.is_available():
        return  # on the GPU box this is covered by the gpu tests
    h = C.c_uint64()
    rc = lib.vx_init(0, C.byref(h))
    assert rc == _abi.VX_E_CUDA
    assert b"no CPU fallback" in lib.vx_last_error()


def test_parser_gtv_spec():
    from voxlogica_b200.imgql import GTV_SPEC, STDLIB, Call, Num, Parser, Var
    st = Parser(GTV_SPEC).program()
    kinds = [s[0] for s in st]
    assert kinds.count("let") == 10 and kinds[0] == "import" and kinds[1] == "load"
    bg = [s for s in st if s[0] == "let" and s[1] == "background"][0]
    assert bg[3] == Call("touch", (Call("infix:<.", (Call("intensity", (Var("flair"),)), Num(0.1))), Var("border")))
    lib = Parser(STDLIB).program()
    assert {s[1] for s in lib} >= {"touch", "grow", "flt", "filter", "complement"}


def test_parser_precedence_and_prefix():
    from voxlogica_b200.imgql import Call, Num, Parser, Var
    (_, _, _, e), = Parser("let x = !a & b | c >. 1 + 2 * 3").program()
    assert e == Call("infix:|", (
        Call("infix:&", (Call("prefix:!", (Var("a"),)), Var("b"))),
        Call("infix:>.", (Var("c"), Call("infix:+", (Num(1.0), Call("infix:*", (Num(2.0), Num(3.0)))))))))


def test_synthetic_generators():
    from voxlogica_b200 import synth
    v = synth.brats_like(seed=3, shape=(20, 32, 36))
    assert v.dtype == np.float32 and v.shape == (20, 32, 36)
    assert v.min() == 0.0 and 0.8 <= v.max() <= 1.0
    assert (v[0, 0, :] == 0).all()                      # background is exactly zero
    assert np.array_equal(v, synth.brats_like(seed=3, shape=(20, 32, 36)))
    m = synth.maze(8, seed=0)
    assert m.shape == (17, 17, 3) and tuple(m[1, 1]) == (0, 0, 255) and tuple(m[15, 15]) == (0, 255, 0)
    from scipy import ndimage as ndi
    corridors = (m.sum(axis=2) > 0)
    assert ndi.label(corridors)[1] == 1                 # a perfect maze is one 4-connected component
    assert corridors.sum() == 2 * 64 - 1                # spanning tree: n cells + n-1 passages


def test_nifti_and_png_roundtrip(tmp_path):
    from voxlogica_b200 import io as vio
    rng = np.random.default_rng(3)
    vol = rng.random((5, 7, 9), dtype=np.float32)
    for name in ("v.nii", "v.nii.gz"):
        p = str(tmp_path / name)
        vio.write_nifti(p, vol, (0.5, 1.25, 3.0))
        back, sp = vio.read_nifti(p)
        assert np.array_equal(back, vol) and sp == (0.5, 1.25, 3.0)
    m = (vol > 0.5).astype(np.uint8)
    vio.write_nifti(str(tmp_path / "m.nii.gz"), m)
    back, sp = vio.read_nifti(str(tmp_path / "m.nii.gz"))
    assert np.array_equal(back, m.astype(np.float32)) and sp == (1.0, 1.0, 1.0)
    # big-endian file with a scale slope: byte-swap the little-endian header fields by hand
    import struct
    hdr = bytearray(open(str(tmp_path / "v.nii"), "rb").read()[:352])
    be = bytearray(352)
    struct.pack_into(">i", be, 0, 348)
    struct.pack_into(">8h", be, 40, *struct.unpack_from("<8h", hdr, 40))
    struct.pack_into(">hh", be, 70, 4, 16)
    struct.pack_into(">8f", be, 76, *struct.unpack_from("<8f", hdr, 76))
    struct.pack_into(">f", be, 108, 352.0)
    struct.pack_into(">ff", be, 112, 2.0, 1.0)
    be[344:348] = b"n+1\0"
    ints = rng.integers(-100, 100, size=vol.shape).astype(">i2")
    with open(str(tmp_path / "be.nii"), "wb") as f:
        f.write(bytes(be)); f.write(ints.tobytes())
    back, _ = vio.read_nifti(str(tmp_path / "be.nii"))
    assert np.array_equal(back, ints.astype(np.float32) * 2 + 1)
    from voxlogica_b200.synth import maze
    img = maze(8, seed=1)
    vio.write_png(str(tmp_path / "maze.png"), img)
    assert np.array_equal(vio.read_png(str(tmp_path / "maze.png")), img)
    ch = vio.load_image(str(tmp_path / "maze.png"))
    assert set(ch) == {"red", "green", "blue"} and ch["red"][0].shape == (1,) + img.shape[:2]
