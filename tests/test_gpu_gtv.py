"""End-to-end: the GTV segmentation spec of Greta.pdf p.48 run by the ImgQL driver on the
GPU, against the oracle composition of the same spec, on small BraTS-like volumes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES_BOOL = ("background", "brain", "hyperIntense", "veryIntense", "growTum", "tumStatCC", "gtv")


@pytest.mark.parametrize("shape,seeds", [((48, 80, 96), (1234, 7)), ((40, 70, 90), (3,))])
def test_gtv_spec_matches_oracle(ctx, oracle, shape, seeds):
    from gtv_ref import gtv_oracle
    from voxlogica_b200.imgql import run_gtv
    from voxlogica_b200.synth import brats_batch
    vols = brats_batch(seeds, shape)
    got = run_gtv(ctx, ctx.upload(vols))
    for i in range(len(seeds)):
        want = gtv_oracle(oracle, vols[i])
        for name in NAMES_BOOL:
            assert np.array_equal(got[name].numpy()[i], want[name]), (name, seeds[i])
        for name in ("normFlair", "tumSim"):
            assert np.array_equal(got[name].numpy()[i].view(np.uint32), want[name].view(np.uint32)), name
        assert got["_printed"]["gtvVolume"][i] == want["gtv"].sum()
    # the spec must find the lesion: a non-trivial GTV strictly inside the brain
    g = got["gtv"].numpy()
    assert g.sum() > 0 and np.all(g <= got["brain"].numpy())
    assert got["_tasks"] < 60     # hash-consing: shared sub-terms evaluated once


@pytest.mark.parametrize("shape", [(48, 80, 96), (33, 50, 70)])
def test_gtv_spec_on_16bit_upload_matches_oracle(ctx, oracle, shape):
    """the bench's input path: int16 voxels + scl_slope uploaded with vx_upload_scaled.  The image keeps its
    16-bit codes, so min/max come from the recorded code range, percentiles are counted and the
    crossCorrelation bins are tabulated per code -- every named intermediate must still equal the oracle
    run on the f32 image, bit for bit, with the quantised paths on and off."""
    from gtv_ref import gtv_oracle
    from voxlogica_b200 import _abi as A
    from voxlogica_b200.imgql import run_gtv
    from voxlogica_b200.synth import brats_like_int16, int16_to_f32
    raws, slope = zip(*[brats_like_int16(sd, shape) for sd in (1234, 9)])
    raw = np.stack(raws)
    f32 = int16_to_f32(raw, slope[0])
    wants = [gtv_oracle(oracle, f32[i]) for i in range(len(raws))]
    for opt in (1, 0):
        ctx.set_option(A.VX_OPT_QUANTISED_PATHS, opt)
        try:
            got = run_gtv(ctx, ctx.upload_scaled(raw, slope[0]))
        finally:
            ctx.set_option(A.VX_OPT_QUANTISED_PATHS, 1)
        for i, want in enumerate(wants):
            for name in NAMES_BOOL:
                assert np.array_equal(got[name].numpy()[i], want[name]), (name, i, opt)
            for name in ("normFlair", "tumSim"):
                assert np.array_equal(got[name].numpy()[i].view(np.uint32), want[name].view(np.uint32)), (name, i, opt)
            assert got["_printed"]["gtvVolume"][i] == want["gtv"].sum()


def test_texture_spec_matches_oracle(ctx, oracle):
    """BASELINE config 4: crossCorrelation texture similarity on four co-registered modalities,
    batched, against the oracle composition (bit-exact, the f32 coefficient maps included)."""
    from gtv_ref import texture_oracle
    from voxlogica_b200.imgql import run_texture
    from voxlogica_b200.synth import MODALITIES, brats_multimodal
    shape, seeds = (40, 64, 80), (1234, 21)
    subj = [brats_multimodal(s, shape) for s in seeds]
    imgs = {m: ctx.upload(np.stack([sj[m] for sj in subj])) for m in MODALITIES}
    got = run_texture(ctx, imgs)
    for i, sj in enumerate(subj):
        want = texture_oracle(oracle, sj)
        for name in ("growTum", "similar", "tumour"):
            assert np.array_equal(got[name].numpy()[i], want[name]), (name, seeds[i])
        for name in ("simFlair", "simT1", "simT1c", "simT2", "texture"):
            assert np.array_equal(got[name].numpy()[i].view(np.uint32), want[name].view(np.uint32)), (name, seeds[i])
        assert got["_printed"]["tumourVolume"][i] == want["tumour"].sum()
    assert got["tumour"].numpy().sum() > 0


def test_maze_reachability_2d(ctx, oracle):
    """BASELINE configs[0]: 2D maze, colour thresholds -> near -> through (reach)."""
    from voxlogica_b200.imgql import run_spec
    from voxlogica_b200.synth import maze
    img = maze(40, seed=0)
    r, g, b = (img[..., k].astype(np.float32) for k in range(3))
    spec = '''
    import "stdlib.imgql"
    load r = "r"
    load g = "g"
    load b = "b"
    let corridor = (r >. 127) & (g >. 127) & (b >. 127)
    let start = (b >. 127) & (r <. 127)
    let exit = (g >. 127) & (r <. 127)
    let path = corridor | start | exit
    let fromStart = through(start, path)
    let reachesExit = volume(fromStart & exit)
    let dead = corridor & !through(near(exit), corridor)
    '''
    for conn in (6, 26):
        it = run_spec(ctx, spec, {k: ctx.upload(v) for k, v in (("r", r), ("g", g), ("b", b))}, conn=conn)
        path = ((img.sum(axis=2) > 0)).astype(np.uint8)
        start = np.zeros_like(path); start[1, 1] = 1
        want = oracle.through(start, path, conn)
        assert np.array_equal(it.globals["fromStart"].numpy()[0, 0], want)
        assert it.globals["reachesExit"].v[0] == 1.0     # a perfect maze connects start and exit
        assert it.globals["dead"].numpy().sum() == 0


def test_specs_from_files(ctx, oracle, tmp_path):
    """examples/*.imgql run unchanged from disk: PNG and NIfTI in, PNG and NIfTI out."""
    import os
    import shutil
    from gtv_ref import gtv_oracle
    from voxlogica_b200 import io as vio
    from voxlogica_b200.imgql import run_file
    from voxlogica_b200.synth import brats_like, maze
    ex = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")
    for name in ("maze.imgql", "gtv.imgql"):
        shutil.copy(os.path.join(ex, name), tmp_path / name)
    img = maze(20, seed=3)
    vio.write_png(str(tmp_path / "maze.png"), img)
    it = run_file(ctx, str(tmp_path / "maze.imgql"), conn=6)
    assert it.printed["solvable"][0] == 1.0
    path = (img.sum(axis=2) > 0).astype(np.uint8)
    start = np.zeros_like(path); start[1, 1] = 1
    want = oracle.through(start, path, 6)
    assert np.array_equal(vio.read_png(str(tmp_path / "fromStart.png")), want * 255)
    vol = brats_like(seed=11, shape=(40, 64, 72))
    vio.write_nifti(str(tmp_path / "flair.nii.gz"), vol)
    it = run_file(ctx, str(tmp_path / "gtv.imgql"))
    ref = gtv_oracle(oracle, vol)
    got, sp = vio.read_nifti(str(tmp_path / "gtv.nii.gz"))
    assert np.array_equal(got.astype(np.uint8), ref["gtv"]) and it.printed["gtvVolume"][0] == ref["gtv"].sum()
    # the same spec from a 16-bit file with scl_slope, as a scanner stores it: run_file sends the int16 voxels
    # (vx_upload_scaled), the operators take their code paths, the result is that of the f32 values
    from voxlogica_b200.synth import brats_like_int16, int16_to_f32
    raw, slope = brats_like_int16(seed=12, shape=(40, 64, 72))
    vio.write_nifti(str(tmp_path / "flair.nii.gz"), raw, slope=float(slope))
    it = run_file(ctx, str(tmp_path / "gtv.imgql"))
    ref = gtv_oracle(oracle, int16_to_f32(raw, slope))
    got, sp = vio.read_nifti(str(tmp_path / "gtv.nii.gz"))
    assert np.array_equal(got.astype(np.uint8), ref["gtv"]) and it.printed["gtvVolume"][0] == ref["gtv"].sum()
    assert ref["gtv"].sum() > 0
