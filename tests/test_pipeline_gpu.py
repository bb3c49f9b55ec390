"""End-to-end drop-in API on the GPU: mvlm.pipeline.create_pipeline(...).predict_one_file(path).

The CNN cannot be compared end to end under random weights (bf16 rounding is chaotic and ~6 % of
arg-max positions jump, SURVEY.md 7 hard part 2), so the landmark check is teacher-forced exactly
as north_star words it: identical peaks in -> oracle rays / consensus / snap -> landmarks within
1e-3 x mesh bounding-box diagonal (measured ~1e-9).
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from mvlm_b200 import synth
from mvlm_b200.io_obj import load_obj
from mvlm_b200.weights import seeded_state_dict
from oracle import native, stages

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scan(tmp_path_factory):
    d = tmp_path_factory.mktemp("scan")
    v, uv, t = synth.face_mesh(grid=100, seed=11)
    synth.write_obj(d / "face.obj", v, uv, t, synth.face_texture(256, seed=11))
    return d / "face.obj"


@pytest.mark.parametrize("name,n_lm", [("dtu3d", 73), ("BU3DFE", 84)])
def test_predict_one_file_teacher_forced(lib, scan, name, n_lm):
    import mvlm

    sd = seeded_state_dict(n_lm, "RGB+depth", seed=3)
    dm = mvlm.pipeline.create_pipeline(name, n_views=12, weights=sd, seed=7, n_hypotheses=16, verbose=False,
                                       image_size=(128, 128))
    assert dm.get_lm_count() == n_lm
    tr = synth.random_view_transforms(12, seed=21)
    dm.renderer_3d.transforms = tr
    lm = dm.predict_one_file(scan)
    assert lm.shape == (n_lm, 3) and lm.dtype == np.float64 and np.isfinite(lm).all()

    # teacher-forced oracle chain from the CUDA peaks
    mesh = load_obj(scan)
    dmesh = dm.renderer_3d.upload(mesh)
    u8 = dm.renderer_3d.render_device(dmesh, tr)["u8"]
    peaks = dm.predictor_2d.predict_landmarks_device(u8).cpu().numpy()
    starts, ends = stages.landmark_lines(128, peaks, tr)
    draws = dm.estimator_3d.seeded_draws(n_lm)
    ref_lm, ref_err, _ = stages.landmarks_from_lines(peaks, starts, ends, draws)
    ref_snap, _ = native.snap_to_mesh(mesh.verts, mesh.tris, ref_lm)
    diag = mesh.bbox_diagonal
    assert np.abs(lm - ref_snap).max() <= 1e-3 * diag
    assert np.abs(lm - ref_snap).max() <= 1e-6  # what is actually achieved
    assert abs(dm.last_error - ref_err) <= 1e-6 * max(1.0, abs(ref_err))
    # landmarks lie on the surface: snapping again is the identity
    again, _ = native.snap_to_mesh(mesh.verts, mesh.tris, lm)
    assert np.abs(again - lm).max() <= 1e-6

    # seam-by-seam path (reference flow with numpy between stages) gives the same landmarks
    lm2 = dm._predict_seams(scan, 0.0)
    assert np.abs(lm2 - lm).max() <= 1e-6


def test_error_behaviour(lib, scan, tmp_path):
    import mvlm

    with pytest.raises(ValueError, match="Unknown pipeline"):
        mvlm.pipeline.create_pipeline("nope")
    sd = seeded_state_dict(73, "RGB+depth", seed=3)
    dm = mvlm.pipeline.create_pipeline("DTU3D", weights=sd, verbose=False, image_size=(64, 64))
    assert dm.n_views == 8 and dm.renderer_3d.generate_3d_transformations().shape == (8, 6)
    assert dm.predict_one_file(tmp_path / "missing.obj") is None           # general_pipeline.py:78-80
    bad = tmp_path / "scan.ply"
    bad.write_text("ply")
    with pytest.raises(ValueError, match="not an .obj"):
        dm.predict_one_file(bad)                                          # render3d.py:185-186
    empty = tmp_path / "empty.obj"
    empty.write_text("# nothing\n")
    with pytest.raises(ValueError, match="does not contain any points"):
        dm.predict_one_file(empty)                                        # utils3d.py:20-21
    dm.predictor_2d = None
    with pytest.raises(ValueError, match="not initialized"):
        dm.predict_one_file(scan)                                         # general_pipeline.py:74-75
    with pytest.raises(ValueError, match="not initialized"):
        dm.get_lm_count()


def test_reference_rng_replay(lib, scan):
    """With seed=None the estimator consumes the GLOBAL numpy RNG exactly like the reference
    (np.random.choice(range(n), 8) per landmark with >= 3 lines, estimator3d.py:105)."""
    import mvlm

    sd = seeded_state_dict(73, "RGB+depth", seed=3)
    dm = mvlm.pipeline.create_pipeline("dtu3d", n_views=8, weights=sd, verbose=False, image_size=(64, 64))
    np.random.seed(99)
    a = dm.predict_one_file(scan)
    state_after = np.random.get_state()[1][:4].copy()
    np.random.seed(99)
    b = dm.predict_one_file(scan)
    assert np.array_equal(a, b)
    # 73 landmarks x one choice(range(n), 8) call were consumed
    np.random.seed(99)
    for _ in range(73):
        np.random.choice(range(4), 8, replace=True)
    assert np.array_equal(np.random.get_state()[1][:4], state_after)


def test_predict_files_prefetching_driver(lib, scan, tmp_path):
    """Pipeline.predict_files (background native loader, scans in order) == a loop of predict_one_file
    (the reference's main.py:50-62), incl. a missing file -> None."""
    import shutil

    import mvlm

    sd = seeded_state_dict(73, "RGB+depth", seed=3)
    dm = mvlm.pipeline.create_pipeline("dtu3d", n_views=8, weights=sd, seed=5, verbose=False, image_size=(64, 64))
    v, uv, t = synth.face_mesh(grid=60, seed=12)
    synth.write_obj(tmp_path / "other.obj", v, uv, t, synth.face_texture(128, seed=12))
    shutil.copy(scan, tmp_path / "face.obj")
    shutil.copy(scan.with_suffix(".jpg"), tmp_path / "face.jpg")
    files = [tmp_path / "face.obj", tmp_path / "other.obj", tmp_path / "missing.obj", tmp_path / "face.obj", tmp_path / "other.obj"]
    serial = [dm.predict_one_file(f) for f in files]
    batch = dm.predict_files(files, prefetch=2)
    assert len(batch) == len(files) and batch[2] is None and serial[2] is None
    for a, b in zip(serial, batch):
        assert (a is None and b is None) or np.array_equal(a, b)
    assert not np.array_equal(batch[0], batch[1])


def test_predict_meshes_pipelined_equals_blocking_calls(lib, scan):
    """Pipeline.predict_meshes (scans in flight back to back on one stream, buffers reused in order) == one blocking
    predict_mesh per scan, for different scans in one batch and with a None entry."""
    import mvlm

    sd = seeded_state_dict(73, "RGB+depth", seed=3)
    dm = mvlm.pipeline.create_pipeline("dtu3d", n_views=8, weights=sd, seed=5, verbose=False, image_size=(64, 64))
    a = load_obj(scan)
    v, uv, t = synth.face_mesh(grid=50, seed=13)
    from mvlm_b200.io_obj import Mesh

    b = Mesh(verts=v, tris=t, uvs=uv, texture=synth.face_texture(64, seed=13))
    batch = [a, b, None, b, a, a, b]
    blocking = [None if m is None else dm.predict_mesh(m) for m in batch]
    for depth in (1, 2, 4):
        got = dm.predict_meshes(batch, depth=depth)
        assert len(got) == len(batch) and got[2] is None
        for x, y in zip(blocking, got):
            assert (x is None and y is None) or np.array_equal(x, y)


def test_nvjpeg_texture_decoder(lib, scan, tmp_path):
    """Opt-in GPU texture decode (csrc/jpeg_decode.cu): same image as the host decoder up to decoder arithmetic
    (mean |diff| of a few LSB on this small, busy texture, < 1 LSB at 1024^2 and above; single pixels on sharp chroma
    edges differ more: libjpeg-turbo interpolates the
    sub-sampled chroma planes, nvJPEG replicates them), the pipeline runs from files with it, and the default stays
    the host decoder."""
    import mvlm

    a = load_obj(scan)                                   # PIL / libjpeg-turbo
    b = load_obj(scan, texture_decoder="nvjpeg")
    assert isinstance(b.texture, torch.Tensor) and b.texture.is_cuda and b.texture_ready is not None
    b.texture_ready.synchronize()
    got = b.texture.cpu().numpy().astype(np.int32)
    ref = a.texture.astype(np.int32)
    assert got.shape == ref.shape
    diff = np.abs(got - ref)
    # a channel swap, a vertical flip or a wrong pitch would give mean differences of tens of LSB
    assert diff.mean() <= 5.0 and (diff > 32).mean() <= 0.02, (diff.max(), diff.mean(), (diff > 32).mean())
    assert np.array_equal(a.verts, b.verts) and np.array_equal(a.tris, b.tris)
    with pytest.raises(ValueError, match="Unknown texture decoder"):
        load_obj(scan, texture_decoder="nope")
    sd = seeded_state_dict(73, "RGB+depth", seed=3)
    dm = mvlm.pipeline.create_pipeline("dtu3d", n_views=8, weights=sd, seed=5, verbose=False, image_size=(64, 64),
                                       texture_decoder="nvjpeg")
    one = dm.predict_one_file(scan)
    many = dm.predict_files([scan, scan, scan])
    assert one.shape == (73, 3) and np.isfinite(one).all()
    assert all(np.array_equal(one, m) for m in many)     # deterministic decode, same result through the batch driver
    assert mvlm.pipeline.create_pipeline("dtu3d", weights=sd, verbose=False, image_size=(64, 64)).texture_decoder == "pil"


def test_pipeline_is_safe_to_share_between_threads(lib, scan):
    """One pipeline shared by a thread pool (the reference's FastAPI server does that without a lock,
    3DMD_server.py:24-31): calls are serialised, every thread gets the landmarks of ITS scan."""
    from concurrent.futures import ThreadPoolExecutor

    import mvlm
    from mvlm_b200.io_obj import Mesh

    sd = seeded_state_dict(73, "RGB+depth", seed=3)
    dm = mvlm.pipeline.create_pipeline("dtu3d", n_views=8, weights=sd, seed=5, verbose=False, image_size=(64, 64))
    meshes = [load_obj(scan)]
    for seed in (21, 22, 23):
        v, uv, t = synth.face_mesh(grid=40 + seed, seed=seed)
        meshes.append(Mesh(verts=v, tris=t, uvs=uv, texture=synth.face_texture(64, seed=seed)))
    serial = [dm.predict_mesh(m) for m in meshes]
    jobs = [i % len(meshes) for i in range(24)]
    with ThreadPoolExecutor(max_workers=4) as pool:
        got = list(pool.map(lambda i: dm.predict_mesh(meshes[i]), jobs))
    for i, g in zip(jobs, got):
        assert np.array_equal(g, serial[i])


def test_predict_folder_writes_reference_txt_files(lib, scan, tmp_path):
    """Batch loop of the reference's main.py:50-62: `<stem>_<pipeline>.txt` written with np.savetxt(delimiter=",")."""
    import shutil

    import mvlm

    src = tmp_path / "in"
    src.mkdir()
    for name in ("b_face", "a_face"):
        shutil.copy(scan, src / f"{name}.obj")
        shutil.copy(scan.with_suffix(".jpg"), src / f"{name}.jpg")
    sd = seeded_state_dict(84, "RGB+depth", seed=3)
    dm = mvlm.pipeline.create_pipeline("BU3DFE", n_views=8, weights=sd, seed=5, verbose=False, image_size=(64, 64))
    res = dm.predict_folder(src, tmp_path / "out")
    assert [f.name for f in res] == ["a_face.obj", "b_face.obj"]
    for f, lm in res.items():
        txt = tmp_path / "out" / f"{f.stem}_bu3dfe.txt"
        assert txt.exists()
        back = np.loadtxt(txt, delimiter=",")
        assert back.shape == (84, 3) and np.array_equal(back, lm)      # %.18e round-trips float64 exactly
    with pytest.raises(ValueError, match="does not contain any .obj"):
        dm.predict_folder(tmp_path / "out")


def test_pre_align_runs_the_path_on_the_aligned_scan_and_maps_back(lib, scan):
    """pre_align (configs/*.json:59-84, estimator3d.py:186-248,257,286): landmarks of a pipeline with pre_align on scan X
    == landmarks of the plain pipeline on the pre-aligned copy of X mapped back with the inverse transform; the identity
    configuration changes nothing; the file path (fused and seam-by-seam) agrees with the array path."""
    import dataclasses

    import mvlm
    from mvlm_b200.utils import prealign

    path, mesh = scan, load_obj(scan)
    tr = synth.random_view_transforms(8, seed=11)
    kw = dict(n_views=8, weights=seeded_state_dict(73, "RGB+depth", 3), seed=4, n_hypotheses=4, verbose=False,
              image_size=(128, 128), transforms=tr)
    cfg = {"align_center_of_mass": True, "rot_x": 10.0, "rot_y": -15.0, "rot_z": 5.0, "scale": 0.9}
    plain = mvlm.pipeline.create_pipeline("dtu3d", **kw)
    aligned = mvlm.pipeline.create_pipeline("dtu3d", pre_align=cfg, **kw)
    a, b = prealign.affine(mesh.verts, cfg)
    moved = dataclasses.replace(mesh, verts=prealign.apply(mesh.verts, a, b))
    want = prealign.invert(plain.predict_mesh(moved), a, b)
    got = aligned.predict_mesh(mesh)
    assert np.array_equal(got, want)
    assert not np.allclose(got, plain.predict_mesh(mesh))            # the alignment does change what the CNN sees
    ident = mvlm.pipeline.create_pipeline("dtu3d", pre_align={"rot_x": 0, "scale": 1}, **kw)
    assert ident.pre_align is None and np.array_equal(ident.predict_mesh(mesh), plain.predict_mesh(mesh))
    from_file = aligned.predict_one_file(path)
    assert np.abs(from_file - got).max() <= 1e-9
    seam = aligned._predict_seams(path, 0.0)                           # the reference's stage-by-stage flow
    assert np.abs(seam - got).max() <= 1e-6
