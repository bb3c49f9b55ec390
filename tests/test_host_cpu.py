"""CPU: C-ABI library loads and exports every declared symbol; host-side logic without a GPU."""
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from mvlm_b200 import synth
from mvlm_b200.io_obj import load_obj
from oracle import stages

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "mvlm_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(mvlm_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/mvlm_b200.h but not exported"
    assert lib.mvlm_version() >= 1
    assert lib.mvlm_last_error() is not None
    # size queries are pure host functions (no compute)
    # packed keys of all pixels + one float4 per (view, vertex)
    assert lib.mvlm_raster_workspace_bytes(100, 256, 256, 50176) == 100 * 256 * 256 * 8 + 100 * 50176 * 16
    # a million-vertex scan is processed in chunks of views: 256 MB of transformed vertices at most
    assert lib.mvlm_raster_workspace_bytes(200, 512, 512, 1002001) <= 200 * 512 * 512 * 8 + (256 << 20)
    assert 2 * 2 ** 30 < lib.mvlm_hourglass_workspace_bytes(73, 4, 100, 256, 256) <= 6e9  # packed (20.3 GB unpacked)
    assert lib.mvlm_consensus_workspace_bytes(84, 200, 16384) > 0
    assert lib.mvlm_snap_workspace_bytes(73, 100000) > 0
    assert lib.mvlm_hourglass_workspace_bytes(73, 4, 1, 100, 100) == 0  # not a multiple of 64 -> error
    assert b"multiple of 64" in lib.mvlm_last_error()


def test_hourglass_flops_match_survey(lib):
    """Algorithmic FLOPs/view of the planned network == SURVEY.md 8(d) (reference census minus conv8)."""
    for (l, cin), want in {(73, 4): 146.106, (73, 3): 146.031, (84, 4): 150.635, (84, 2): 150.484}.items():
        assert abs(lib.mvlm_hourglass_flops_per_view(l, cin, 256, 256) / 1e9 - want) < 1.5e-3
    assert abs(lib.mvlm_hourglass_flops_per_view(73, 4, 512, 512) / lib.mvlm_hourglass_flops_per_view(73, 4, 256, 256) - 4) < 1e-9


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle."""
    for p in (ROOT / "mvlm_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p
    for p in (ROOT / "mvlm_b200" / "csrc").glob("*"):
        assert "oracle_native" not in p.read_text().replace("oracle/csrc/oracle_native.c", ""), p


def test_missing_library_fails_loudly(tmp_path):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from mvlm_b200 import _lib\n"
            "from pathlib import Path\n"
            "_lib.LIB_PATH = Path(%r) / 'nope.so'\n"
            "try:\n    _lib.load()\nexcept _lib.MvlmError as e:\n    print('LOUD', e)\n") % (str(ROOT), str(tmp_path))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "LOUD" in out and "no CPU fallback" in out


def test_rotation_and_views_match_oracle():
    from mvlm_b200.utils.render3d import ObjRenderer3D, fixed_eight_views, rotation_matrices

    for tr in (fixed_eight_views(), synth.random_view_transforms(9, seed=3)):
        assert np.array_equal(rotation_matrices(tr), stages.rotation_matrices(tr))
    r = ObjRenderer3D.__new__(ObjRenderer3D)
    r.__dict__.update(n_views=8, transforms=None)
    assert r.generate_3d_transformations().dtype == np.float32
    r2 = ObjRenderer3D(n_views=5, device="cpu")
    np.random.seed(4)
    a = r2.generate_3d_transformations()
    assert a.shape == (5, 6) and a.dtype == np.float64
    assert np.array_equal(a, synth.random_view_transforms(5, seed=4))   # same draw order as render3d.py:79-89
    assert (a[:, 0] >= -40).all() and (a[:, 0] < 40).all() and (a[:, 1] >= -80).all() and (a[:, 2] < 20).all()


def test_obj_roundtrip_and_vertex_duplication(tmp_path):
    v, uv, t = synth.face_mesh(grid=12, seed=0)
    synth.write_obj(tmp_path / "a.obj", v, uv, t, synth.face_texture(32, 0))
    m = load_obj(tmp_path / "a.obj")
    assert m.verts.shape == v.shape and m.tris.shape == t.shape and m.texture.shape == (32, 32, 3)
    assert np.allclose(m.verts[m.tris], v[t], atol=1e-5) and np.allclose(m.uvs[m.tris], uv[t], atol=1e-5)
    # a position used with two different vt indices is duplicated (vtkOBJReader behaviour); quads are fanned
    (tmp_path / "b.obj").write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 0.5 0.5\n"
                                    "f 1/1 2/2 3/3 4/4\nf 1/5 3/3 2/2\n")
    m = load_obj(tmp_path / "b.obj")
    assert m.tris.shape == (3, 3) and m.verts.shape[0] == 5
    (tmp_path / "c.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\nf -3 -2 -1\n")
    m = load_obj(tmp_path / "c.obj")
    assert m.uvs is None and m.tris.tolist() == [[0, 1, 2], [0, 1, 2]]
    (tmp_path / "d.obj").write_text("# empty\n")
    with pytest.raises(ValueError, match="does not contain any points"):
        load_obj(tmp_path / "d.obj")
    # FLAME models take `mean_texture.jpg` of their folder instead of `<stem>.jpg` (utils3d.py:39-51)
    from PIL import Image

    synth.write_obj(tmp_path / "flame_head.obj", v, uv, t, synth.face_texture(32, 0))
    Image.fromarray(synth.face_texture(16, 5)).save(tmp_path / "mean_texture.jpg")
    assert load_obj(tmp_path / "flame_head.obj").texture.shape == (16, 16, 3)
    (tmp_path / "mean_texture.jpg").unlink()
    assert load_obj(tmp_path / "flame_head.obj").texture is None


def test_native_obj_loader_matches_python_restatement(tmp_path):
    """csrc/obj_loader.cu (multi-threaded C++ parse behind the C-ABI) == the numpy restatement, bit for bit:
    index forms a, a/b, a/b/c, a//c, negative (relative) indices between interleaved v / f blocks, polygons,
    CRLF line ends, '+' signs and exponents, vertices shared by several vt, unreferenced positions."""
    from oracle.obj_ref import load_obj_python

    rng = np.random.RandomState(7)
    cases = {
        "forms.obj": "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 2 2 2\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
                     "f 1/1/1 2/2/1 3/3/1 4/4/1\nf 1//1 2//1 3//1\nf 1/3 2/2 4/1\nf 4 3 2 1\n",
        "relative.obj": "v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nf -3/-3 -2/-2 -1/-1\n"
                        "v 5 5 5\nv 6 5 5\nv 5 6 5\nvt 0.5 0.5\nf -3/-1 -2/2 -1/-4\nf 1/1 5/4 6/2\n",
        "crlf.obj": "# comment\r\nv +1.5e0 -2.25E-1 3\r\nv 1e-3 2 3\r\nv 0.1 0.2 0.3\r\n\r\nf 1 2 3\r\n",
        "nouv.obj": "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 9 9 9\nvt 0 0\nf 1 2 3\n",
    }
    v, uv, t = synth.face_mesh(grid=40, seed=3)
    synth.write_obj(tmp_path / "grid.obj", v, uv, t, None)
    # a big random soup: many positions with several vt each, spans several parser threads
    n_v, n_t, n_f = 3000, 2500, 9000
    lines = [f"v {a:.6f} {b:.6f} {c:.6f}" for a, b, c in rng.randn(n_v, 3)] + [f"vt {a:.5f} {b:.5f}" for a, b in rng.rand(n_t, 2)]
    for _ in range(n_f):
        k = rng.randint(3, 6)
        lines.append("f " + " ".join(f"{rng.randint(1, n_v + 1)}/{rng.randint(1, n_t + 1)}" for _ in range(k)))
    (tmp_path / "soup.obj").write_text("\n".join(lines) + "\n")
    for name, text in cases.items():
        (tmp_path / name).write_text(text)
    for name in list(cases) + ["grid.obj", "soup.obj"]:
        ref = load_obj_python(tmp_path / name)
        for threads in (1, 3, 0):
            got = load_obj(tmp_path / name, n_threads=threads)
            assert np.array_equal(got.verts, ref.verts) and np.array_equal(got.tris, ref.tris), name
            assert (got.uvs is None) == (ref.uvs is None) and (ref.uvs is None or np.array_equal(got.uvs, ref.uvs)), name
    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    with pytest.raises(ValueError, match="references vertex"):
        load_obj(tmp_path / "bad.obj")
    with pytest.raises(ValueError, match="does not exist"):
        load_obj(tmp_path / "missing.obj")


def test_create_pipeline_name_handling():
    import mvlm

    with pytest.raises(ValueError, match="Unknown pipeline: nope"):
        mvlm.pipeline.create_pipeline("NoPe")
    with pytest.raises(ValueError, match="Unknown pipeline"):
        mvlm.pipeline.create_pipeline("mediapipe")
    assert mvlm.pipeline is sys.modules["mvlm.pipeline"]
    from mvlm.prediction import BU3DFEPredictor, DTU3DPredictor, Predictor2D  # noqa: F401
    from mvlm.utils import Estimator3D, ObjVTKRenderer3D  # noqa: F401


def test_estimator_reference_draws_replay_global_rng():
    from mvlm_b200.utils.estimator3d import Estimator3D

    peaks, _, _, _ = synth.synthetic_rays(n_landmarks=5, n_views=20, seed=1)
    peaks[3, :, 2] = 0.3  # all equal -> no lines -> no draw consumed
    e = Estimator3D(device="cpu")
    np.random.seed(11)
    d = e.reference_draws(peaks)
    np.random.seed(11)
    for lm in range(5):
        n = int(stages.line_filter_mask(peaks[lm, :, 2], "quantile", 0.5, 0.5).sum())
        if n >= 3:
            assert np.array_equal(d[lm, 0], np.random.choice(range(n), 8, replace=True))
        else:
            assert (d[lm] == 0).all()
    e.mode = "bogus"
    with pytest.raises(ValueError, match="Unknown mode for line matching"):
        e.reference_draws(peaks)


def test_checkpoint_name_matching():
    """Weight auto-discovery picks the reference's file of exactly this model / mode (paulsenpredictor.py:15-39)."""
    from mvlm_b200.prediction.paulsenpredictor import PaulsenModel

    names = ["MVLMModel_DTU3D_RGB_07092019_only_state_dict-c0255a70.pth", "MVLMModel_DTU3D_Depth_19092019_only_state_dict-95b89b63.pth",
             "MVLMModel_DTU3D_geometry_only_state_dict-41851074.pth", "MVLMModel_DTU3D_geometry+depth_20102019_15epoch_only_state_dict-73b20e31.pth",
             "MVLMModel_DTU3D_RGB+depth_20092019_only_state_dict-e3c12463a9.pth", "MVLMModel_BU_3DFE_RGB+depth_05102019_5epoch-90e29350.pth"]

    class M(PaulsenModel):
        def __init__(self, model_type, mode):
            self.model_type, self.image_mode = model_type, mode

        def get_lm_count(self):
            return 73

    for mode, want in (("RGB", 0), ("depth", 1), ("geometry", 2), ("geometry+depth", 3), ("RGB+depth", 4)):
        hits = [i for i, n in enumerate(names) if M("MVLMModel_DTU3D", mode)._checkpoint_matches(n)]
        assert hits == [want], (mode, hits)
    assert [i for i, n in enumerate(names) if M("MVLMModel_BU_3DFE", "RGB+depth")._checkpoint_matches(n)] == [5]


def test_pre_align_transform_matches_vtk_call_order():
    """mvlm_b200/utils/prealign.py == the 4x4 composition of the legacy vtkTransform calls (oracle/stages.py), its
    inverse returns the points, and the order is the one of estimator3d.py:198-209 (translate first, scale last)."""
    from mvlm_b200.utils import prealign
    from oracle import stages

    rng = np.random.RandomState(5)
    verts = (rng.rand(500, 3) * 100 + [10, -20, 30]).astype(np.float32)
    cfg = {"align_center_of_mass": True, "rot_x": 12.0, "rot_y": -35.0, "rot_z": 7.5, "scale": 1.25}
    a, b = prealign.affine(verts, cfg)
    m = stages.pre_align_matrix(verts, cfg)
    assert np.abs(m[:3, :3] - a).max() <= 1e-14 and np.abs(m[:3, 3] - b).max() <= 1e-11
    moved = prealign.apply(verts, a, b)
    assert moved.dtype == np.float32
    want = (np.c_[verts.astype(np.float64), np.ones(len(verts))] @ m.T)[:, :3]
    assert np.abs(moved - want).max() <= 1e-4                       # float32 storage of ~100 mm coordinates
    assert np.abs(moved.astype(np.float64).mean(0)).max() <= 1e-4    # centre of mass at the origin
    assert np.abs(prealign.invert(want, a, b) - verts).max() <= 1e-9
    # a point on the z axis: Rz leaves it, Rx(90) takes it to -y, then scale 2
    a2, b2 = prealign.affine(verts, {"rot_x": 90.0, "rot_z": 45.0, "scale": 2.0})
    assert np.allclose(a2 @ [0, 0, 1.0] + b2, [0, -2.0, 0], atol=1e-12)
    assert prealign.is_identity(None) and prealign.is_identity({"rot_x": 0, "scale": 1, "align_center_of_mass": False})
    assert not prealign.is_identity({"align_center_of_mass": True})
    with pytest.raises(ValueError):
        prealign.affine(verts, {"rot_w": 1})


def test_obj_loader_cases_documented_from_vtkobjreader(tmp_path):
    """Scan files as vtkOBJReader / obj_to_actor (utils3d.py:10-36) treat them, checked against HAND-WRITTEN per-corner
    arrays (independent of oracle/obj_ref.py, which restates the same rules): statements other than v / vt / f are
    ignored (o, g, s, usemtl, mtllib, vn, comments); polygons become a triangle fan in corner order; a position used
    with different vt indices appears once per (position, vt) pair so that the texture seam survives; a file whose faces
    carry no vt loads without texture coordinates."""
    text = """# exported by a scanner
mtllib scan.mtl
o head
v 0 0 0
v 2 0 0
v 2 2 0
v 0 2 0
v 1 1 5
vn 0 0 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vt 0.25 0.75
g front
usemtl skin
s 1
f 1/1/1 2/2/1 3/3/1 4/4/1
g seam
s off
f 1/5 5/3 2/2
"""
    (tmp_path / "s.obj").write_text(text)
    m = load_obj(tmp_path / "s.obj")
    want_pos = np.array([[[0, 0, 0], [2, 0, 0], [2, 2, 0]],      # fan of the quad: (c0, c1, c2), (c0, c2, c3)
                         [[0, 0, 0], [2, 2, 0], [0, 2, 0]],
                         [[0, 0, 0], [1, 1, 5], [2, 0, 0]]], np.float32)
    want_uv = np.array([[[0, 0], [1, 0], [1, 1]],
                        [[0, 0], [1, 1], [0, 1]],
                        [[0.25, 0.75], [1, 1], [1, 0]]], np.float32)
    assert m.tris.shape == (3, 3) and m.tris.dtype == np.int32
    assert np.array_equal(m.verts[m.tris], want_pos) and np.array_equal(m.uvs[m.tris], want_uv)
    # position 1 carries vt 1 and vt 5 -> two vertices; position 3 carries vt 3 twice (both faces) -> one; 6 in total
    assert m.verts.shape == (6, 3) and m.uvs.shape == (6, 2)
    assert m.texture is None                                            # no s.jpg next to it: plain white actor (:61-64)
    # a file whose faces carry no vt at all: no texture coordinates, even if vt lines exist
    (tmp_path / "n.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 1\nusemtl m\nf 1 2 3\n")
    n = load_obj(tmp_path / "n.obj")
    assert n.uvs is None and n.verts.shape == (3, 3) and n.tris.tolist() == [[0, 1, 2]]
    # a pentagon: fan of three triangles in corner order
    (tmp_path / "p.obj").write_text("v 0 0 0\nv 1 0 0\nv 2 1 0\nv 1 2 0\nv 0 1 0\nf 1 2 3 4 5\n")
    p = load_obj(tmp_path / "p.obj")
    assert p.tris.tolist() == [[0, 1, 2], [0, 2, 3], [0, 3, 4]]
