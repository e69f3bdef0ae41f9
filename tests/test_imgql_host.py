"""Host-side logic that needs no GPU: the ImgQL parser, the C-ABI export list, the
synthetic generators."""
import os
import re

import numpy as np
import pytest

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


def test_ctypes_table_matches_the_header():
    """voxlogica_b200/_abi.py::SIGNATURES (the tested twin of the F# table) has every export of the header
    with the same number of parameters"""
    from voxlogica_b200 import _abi
    with open(os.path.join(ROOT, "include", "voxcuda.h")) as f:
        header = f.read()
    protos = re.findall(r"VX_API\s+[\w\s\*]+?\b(vx_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)
    for name, args in protos:
        if name in ("vx_version", "vx_last_error"):        # bound by hand (non-status return types)
            continue
        assert name in _abi.SIGNATURES, name
        n_c = 0 if args.strip() in ("", "void") else len(args.split(","))
        assert n_c == len(_abi.SIGNATURES[name]), name


def test_fsharp_binding_declares_every_export():
    """integration/VoxCuda.fs (the P/Invoke table a VoxLogicA maintainer adds; never compiled here) must
    declare every VX_API function of include/voxcuda.h with the same number of parameters"""
    with open(os.path.join(ROOT, "include", "voxcuda.h")) as f:
        header = f.read()
    with open(os.path.join(ROOT, "integration", "VoxCuda.fs")) as f:
        fs = f.read()
    protos = re.findall(r"VX_API\s+[\w\s\*]+?\b(vx_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)
    decl = dict(re.findall(r"extern\s+\w+\s+(vx_\w+)\s*\(([^)]*)\)", fs))
    assert len(protos) > 80
    for name, args in protos:
        assert name in decl, f"{name} has no DllImport in VoxCuda.fs"
        n_c = 0 if args.strip() in ("", "void") else len(args.split(","))
        n_fs = 0 if not decl[name].strip() else len(decl[name].split(","))
        assert n_c == n_fs, f"{name}: {n_c} parameters in the header, {n_fs} in VoxCuda.fs"
    assert set(decl) == {n for n, _ in protos}


def test_abi_constants_match_the_header():
    """every integer `#define VX_* n` of include/voxcuda.h that voxlogica_b200/_abi.py names has the same value
    there (status codes, dtypes, options, opcodes of the fusion hook), and the options are all mirrored"""
    from voxlogica_b200 import _abi
    hdr = open(os.path.join(ROOT, "include", "voxcuda.h")).read()
    defines = {k: int(v, 0) for k, v in re.findall(r"#define\s+(VXP?_[A-Z0-9_]+)\s+\(?(-?(?:0x)?[0-9a-fA-F]+)\)?\s*(?:/\*.*)?$", hdr, flags=re.M)}
    enums = re.findall(r"\b(VXP?_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)
    defines.update({k: int(v) for k, v in enums})
    assert len(defines) > 20
    seen = 0
    for name, value in defines.items():
        if hasattr(_abi, name):
            assert getattr(_abi, name) == value, name
            seen += 1
    assert seen > 20
    for name in defines:
        if name.startswith("VX_OPT_"):
            assert hasattr(_abi, name), name


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


def test_texture_spec_parses_and_oracle_finds_the_lesion(oracle):
    """BASELINE config 4 on the CPU side: the multi-modal spec parses, the synthetic modalities share
    their anatomy, and the oracle composition of the spec segments the lesion."""
    from gtv_ref import texture_oracle
    from voxlogica_b200 import synth
    from voxlogica_b200.imgql import TEXTURE_SPEC, Parser
    st = Parser(TEXTURE_SPEC).program()
    assert [s[1] for s in st if s[0] == "load"] == ["flair", "t1", "t1c", "t2"]
    assert {"simFlair", "simT1", "simT1c", "simT2", "texture", "tumour"} <= {s[1] for s in st if s[0] == "let"}
    with open(os.path.join(ROOT, "examples", "texture.imgql")) as f:
        assert len(Parser(f.read()).program()) == len(st)
    shape = (40, 64, 80)
    mods = synth.brats_multimodal(1234, shape)
    assert tuple(mods) == synth.MODALITIES
    head = mods["flair"] > 0
    for name, v in mods.items():
        assert v.dtype == np.float32 and v.shape == shape and np.array_equal(v > 0, head), name
    r = texture_oracle(oracle, mods)
    assert r["growTum"].sum() > 0 and r["tumour"].sum() >= r["growTum"].sum()
    assert np.all(r["tumour"] <= r["brain"]) and -1.0 <= r["texture"].min() and r["texture"].max() <= 1.0


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


class _FakeImage:
    pass


def test_pointwise_fusion_compiles_correct_programs():
    """The lazy pointwise DAG of the ImgQL driver must compile to vx_fused_pointwise programs that
    compute what the operator-by-operator evaluation computes (checked with a numpy interpreter of
    the register machine of include/voxcuda.h standing in for the device)."""
    from voxlogica_b200 import _abi as A
    from voxlogica_b200 import imgql
    from voxlogica_b200.ops import Image

    class FakeImage(Image):
        __slots__ = ("arr",)
        _next = [1]

        def __init__(self, arr):
            self.ctx, self.handle, self.arr = None, FakeImage._next[0], arr
            FakeImage._next[0] += 1

        def __del__(self):
            pass

        dtype = property(lambda self: self.arr.dtype.type)
        nb = property(lambda self: self.arr.shape[0])
        shape = property(lambda self: self.arr.shape)

        def numpy(self):
            return self.arr

    launches = []
    CMP = [np.less, np.less_equal, np.greater, np.greater_equal, np.equal, np.not_equal]
    AR = [np.add, np.subtract, np.multiply, np.divide, lambda x, y: y - x, lambda x, y: y / x, np.minimum, np.maximum]

    class FakeCtx:
        def fused(self, prog, leaves, bs):
            assert len(prog) <= A.VXP_MAX_OPS and len(leaves) <= A.VXP_MAX_LEAVES
            launches.append(len(prog))
            R = {}
            for opc, dst, a, b, aux, imm in prog:
                assert 0 <= dst < A.VXP_MAX_REGS
                if opc == A.VXP_LOAD: R[dst] = leaves[a].arr
                elif opc == A.VXP_NOT: R[dst] = (1 - R[a]).astype(np.uint8)
                elif opc == A.VXP_AND: R[dst] = R[a] & R[b]
                elif opc == A.VXP_OR: R[dst] = R[a] | R[b]
                elif opc == A.VXP_CMPS: R[dst] = CMP[aux](R[a], np.float32(imm)).astype(np.uint8)
                elif opc == A.VXP_CMPSB: R[dst] = CMP[aux](R[a], bs[b][:, None, None, None]).astype(np.uint8)
                elif opc == A.VXP_CMP: R[dst] = CMP[aux](R[a], R[b]).astype(np.uint8)
                elif opc == A.VXP_ARITHS: R[dst] = AR[aux](R[a], np.float32(imm)).astype(np.float32)
                elif opc == A.VXP_ARITH: R[dst] = AR[aux](R[a], R[b]).astype(np.float32)
                elif opc == A.VXP_SELECT: R[dst] = np.where(R[a] != 0, R[b], np.float32(imm)).astype(np.float32)
                else: raise AssertionError(opc)
                last = dst
            return FakeImage(R[last])

    rng = np.random.default_rng(0)
    shp = (2, 3, 4, 5)
    x, y = rng.random(shp, dtype=np.float32), rng.random(shp, dtype=np.float32)
    m = (rng.random(shp) < 0.5).astype(np.uint8)
    it = imgql.Interpreter(FakeCtx(), loader=lambda n: {"x": FakeImage(x), "y": FakeImage(y), "m": FakeImage(m)}[n])
    it.run("""
        load x = "x"
        load y = "y"
        load m = "m"
        let a = (x >. 0.3) & !(y <. 0.6) | m
        let b = mask((x + y) * 2 - 1 / (y + 1), a & (x <. y))
        let c = a & a
        let wide = ((x >. 0.1) & (x >. 0.2) & (x >. 0.3) & (x >. 0.4) & (x >. 0.5) & (x >. 0.6) & (x >. 0.7) & (x >. 0.8)) | ((y >. 0.1) & (y >. 0.2) & (y >. 0.3) & (y >. 0.4) & (y >. 0.5) & (y >. 0.6) & (y >. 0.7) & (y >. 0.8)) | ((y <. 0.15) & (x <. 0.25) & !m)
    """)
    g = it.globals
    a_want = (((x > np.float32(0.3)) & ~(y < np.float32(0.6))) | (m != 0)).astype(np.uint8)
    assert np.array_equal(g["a"].numpy(), a_want)
    assert launches == [len(launches) and launches[0]] and launches[0] <= 10       # ONE program for the whole chain
    t = ((x + y) * np.float32(2) - (np.float32(1) / (y + np.float32(1)))).astype(np.float32)
    b_want = np.where((a_want != 0) & (x < y), t, np.float32(0))
    assert np.array_equal(g["b"].numpy(), b_want)
    assert np.array_equal(g["c"].numpy(), a_want)
    wide = np.zeros(shp, bool) | (x > np.float32(0.8)) | (y > np.float32(0.8)) | ((y < np.float32(0.15)) & (x < np.float32(0.25)) & (m == 0))
    assert np.array_equal(g["wide"].numpy(), wide.astype(np.uint8))
    # unfused evaluation is still available
    it2 = imgql.Interpreter(FakeCtx(), fuse=False)
    assert it2.fuse is False


def test_xcorr_bin_edges_reproduce_the_binning_exactly():
    """The CUDA kernels bin with a host-computed table of break points instead of one binary64
    division per voxel; the table must reproduce bin_of (oracle/vx_oracle.c) for EVERY float."""
    import ctypes as C
    from voxlogica_b200 import _abi
    lib = _abi.load_library()
    rng = np.random.default_rng(5)

    def oracle_bin(v, m, M, k):
        dv = v.astype(np.float64)
        t = ((dv - m) * k) / (M - m)
        with np.errstate(invalid="ignore"):
            i = np.minimum(t.astype(np.int64), k - 1)
        return np.where((dv >= m) & (dv < M), i, k)

    cases = [(0.0, 1.0, 100), (0.1, 0.9, 32), (-3.7, 12.25, 254), (1e-3, 1.0000001e-3, 7), (0.0, 0.97265625, 100),
             (5.0, 5.0, 10), (2.0, 1.0, 10), (-1e30, 1e30, 3), (0.0, 1e-40, 5)]
    for m, M, k in cases:
        e = (C.c_float * (k + 1))()
        assert lib.vx_xcorr_bin_edges(m, M, k, e) == 0
        e = np.frombuffer(e, dtype=np.float32)
        if not M > m:
            assert np.all(np.isinf(e))
            continue
        assert np.all(e[1:].astype(np.float64) >= e[:-1].astype(np.float64))
        lo, hi = np.float32(m), np.float32(M)
        span = float(hi) - float(lo)
        v = np.concatenate([
            rng.uniform(float(lo) - 0.1 * abs(span) - 1e-30, float(hi) + 0.1 * abs(span) + 1e-30, 200000).astype(np.float32),
            e, np.nextafter(e, np.float32(-np.inf)), np.nextafter(e, np.float32(np.inf)),
            np.array([0.0, -0.0, np.inf, -np.inf, lo, hi], dtype=np.float32)])
        got = np.searchsorted(e[1:k].astype(np.float64), v.astype(np.float64), side="right")
        got = np.where((v >= e[0]) & (v < e[k]), got, k)
        assert np.array_equal(got, oracle_bin(v, m, M, k)), (m, M, k)


def test_nifti_raw_reader_keeps_file_dtype_and_scaling(tmp_path):
    """read_nifti_raw hands out the file's own voxels + scl_slope/scl_inter (what vx_upload_scaled takes);
    read_nifti is the same data converted on the host."""
    import struct

    from voxlogica_b200 import io
    a = np.random.default_rng(3).integers(0, 4096, size=(4, 5, 6)).astype(np.int16)
    p = str(tmp_path / "v.nii.gz")
    io.write_nifti(p, a, (1.0, 1.5, 2.0))
    raw, slope, inter, sp = io.read_nifti_raw(p)
    assert raw.dtype == np.int16 and np.array_equal(raw, a) and (slope, inter) == (1.0, 0.0) and sp == (1.0, 1.5, 2.0)
    # patch scl_slope / scl_inter into the header
    import gzip
    blob = bytearray(gzip.open(p, "rb").read())
    struct.pack_into("<ff", blob, 112, 1.0 / 4095.0, 0.25)
    p2 = str(tmp_path / "w.nii")
    open(p2, "wb").write(bytes(blob))
    raw, slope, inter, _ = io.read_nifti_raw(p2)
    arr, _ = io.read_nifti(p2)
    assert np.array_equal(raw, a)
    assert np.array_equal(arr, a.astype(np.float32) * np.float32(slope) + np.float32(inter))
    struct.pack_into("<ff", blob, 112, 0.0, 5.0)      # slope 0: "no scaling" (NIfTI-1), inter ignored
    open(p2, "wb").write(bytes(blob))
    arr, _ = io.read_nifti(p2)
    assert np.array_equal(arr, a.astype(np.float32))


def test_run_file_sends_16_bit_files_as_16_bits(tmp_path):
    """run_file loads a 16-bit NIfTI file (int16 / uint16 + scl_slope / scl_inter, what scanners and BraTS
    store) with upload_scaled -- the file's own voxels and the header's scaling -- and everything else with
    upload; write_nifti stores the scaling it is given.  (A recording stand-in for the device.)"""
    from voxlogica_b200 import imgql, io
    from voxlogica_b200.ops import Image

    calls = []

    class FakeImage(Image):
        __slots__ = ()

        def __init__(self):
            self.ctx, self.handle = None, 1

        def __del__(self):
            pass

    class FakeCtx:
        def upload_scaled(self, raw, slope, inter=0.0, spacing=None):
            calls.append(("scaled", raw.dtype, raw.shape, slope, inter, tuple(spacing)))
            return FakeImage()

        def upload(self, arr, spacing=None, batched=None):
            calls.append(("f32", arr.dtype, arr.shape, tuple(spacing)))
            return FakeImage()

    rng = np.random.default_rng(5)
    spec = 'load a = "a.nii.gz"\nlet x = intensity(a)\n'
    (tmp_path / "s.imgql").write_text(spec)
    for dtype, slope, inter, want in ((np.int16, 1.0 / 4095.0, 0.0, "scaled"), (np.uint16, 0.5, -3.0, "scaled"),
                                      (np.float32, 1.0, 0.0, "f32"), (np.uint8, 1.0, 0.0, "f32")):
        vol = rng.integers(0, 200, size=(3, 4, 5)).astype(dtype)
        io.write_nifti(str(tmp_path / "a.nii.gz"), vol, (1.0, 2.0, 0.5), slope=slope, inter=inter)
        raw, s2, i2, sp = io.read_nifti_raw(str(tmp_path / "a.nii.gz"))
        assert np.array_equal(raw, vol) and s2 == np.float32(slope) and i2 == np.float32(inter) and sp == (1.0, 2.0, 0.5)
        calls.clear()
        imgql.run_file(FakeCtx(), str(tmp_path / "s.imgql"))
        assert len(calls) == 1 and calls[0][0] == want, calls
        if want == "scaled":
            assert calls[0][1] == dtype and calls[0][2] == (3, 4, 5) and calls[0][3] == np.float32(slope) \
                   and calls[0][4] == np.float32(inter) and calls[0][5] == (1.0, 2.0, 0.5)


# ------------------------------------------------------------------ GC policy of the driver (no GPU)
class _CountedImage:
    """numpy-backed stand-in for an HBM image that counts how many are alive"""
    is_image = True
    live = 0
    peak = 0

    def __init__(self, a):
        self.a = a
        _CountedImage.live += 1
        _CountedImage.peak = max(_CountedImage.peak, _CountedImage.live)

    def __del__(self):
        _CountedImage.live -= 1

    nb = 1


class _NumpyCtx:
    """the operators task_dag_spec uses, on numpy arrays (scipy.ndimage for the spatial ones)"""
    def __init__(self):
        from scipy import ndimage as ndi
        self.ndi, self.s26 = ndi, np.ones((3, 3, 3), bool)
    def _w(self, a): return _CountedImage(a)
    def not_(self, x): return self._w(~x.a)
    def and_(self, x, y): return self._w(x.a & y.a)
    def or_(self, x, y): return self._w(x.a | y.a)
    def cmp_scalar(self, x, k, s): return self._w({">": np.greater, "<": np.less}[k](x.a, np.float32(s)))
    def arith_scalar(self, x, op, s): return self._w({"*": np.multiply, "+": np.add}[op](x.a, np.float32(s)))
    def arith(self, x, op, y): return self._w({"*": np.multiply, "+": np.add}[op](x.a, y.a))
    def mask(self, x, m, fill): return self._w(np.where(m.a, x.a, np.float32(fill)))
    def near(self, x, conn): return self._w(self.ndi.binary_dilation(x.a, self.s26))
    def interior(self, x, conn): return self._w(self.ndi.binary_erosion(x.a, self.s26, border_value=1))
    def distlt(self, r, x): return self._w(self.ndi.distance_transform_edt(~x.a) < r if x.a.any() else np.zeros_like(x.a))
    def through(self, x, y, conn):
        lab, _ = self.ndi.label(y.a, self.s26)
        return self._w(np.isin(lab, np.unique(lab[x.a & y.a])) & y.a)
    def volume(self, x): return np.array([float(x.a.sum())])


@pytest.mark.parametrize("n_tasks", [11, 67, 259, 1027])
def test_driver_gc_policy_bounds_live_images(n_tasks):
    """Greta.pdf p.79: the reference's GPU branch needs garbage collection to survive 4099 tasks.
    The driver's policy (weak memo + names dropped after their last use) keeps the number of live
    images bounded by the formula's working set, not by its size, and changes no result."""
    import gc as pygc
    from voxlogica_b200 import imgql
    from voxlogica_b200.synth import task_dag_spec
    spec = task_dag_spec(n_tasks, seed=n_tasks)
    vol = np.random.default_rng(5).random((6, 10, 12)).astype(np.float32)
    res = {}
    for mode in (False, True):
        pygc.collect()
        _CountedImage.live = _CountedImage.peak = 0
        it = imgql.run_spec(_NumpyCtx(), spec, {"vol": _CountedImage(vol)}, fuse=False, gc=mode)
        res[mode] = (float(it.printed["volume"][0]), it.n_tasks, _CountedImage.peak)
        del it
    assert res[True][0] == res[False][0]
    assert res[True][1] == res[False][1] == n_tasks          # nothing was recomputed, nothing memo-merged
    assert res[False][2] >= n_tasks - 4                      # without GC every result stays alive
    assert res[True][2] <= 20                                # with GC: the working set (two windows of 6 + operands)
