"""Size-independent properties at the FULL headline size (BASELINE.json configs[2]: 100 views of 256^2, 50k-vertex scan),
where the CPU oracle is too slow to run everything:
  * raster: tri-id / depth maps of three of the 100 views == the C oracle's (bit-exact), coverage sane;
  * CNN: graph replay == plain launches == a second run (determinism), fused arg-max keys == arg-max of the heat maps
    the same plan materialises (bit-exact index and value) on every view;
  * tail: landmarks finite, on the surface (snapping is idempotent), identical between two runs.
"""
import numpy as np
import pytest
import torch

from mvlm_b200 import synth
from mvlm_b200.io_obj import Mesh
from mvlm_b200.weights import seeded_state_dict
from oracle import native, stages

pytestmark = pytest.mark.gpu

V, S = 100, 256


@pytest.fixture(scope="module")
def setup(lib):
    from mvlm_b200.pipeline import create_pipeline

    v, uv, t = synth.face_mesh(grid=224, seed=1234)
    mesh = Mesh(verts=v, tris=t, uvs=uv, texture=synth.face_texture(1024, seed=1234))
    tr = synth.random_view_transforms(V, seed=1234)
    dm = create_pipeline("dtu3d", n_views=V, weights=seeded_state_dict(73, "RGB+depth", 1234), seed=1234, n_hypotheses=1,
                         verbose=False, image_size=(S, S), transforms=tr)
    return dm, mesh, tr


def test_fullsize_raster_matches_oracle_on_sampled_views(setup):
    dm, mesh, tr = setup
    dmesh = dm.renderer_3d.upload(mesh)
    out = dm.renderer_3d.render_device(dmesh, tr, want_tri=True, want_z=True)
    tri = out["tri"].cpu().numpy()
    z = out["z"].cpu().numpy()
    views = [0, 37, 99]
    rot = stages.rotation_matrices(tr)[views]
    _, ref_tri, ref_z = native.raster_multiview(mesh.verts, mesh.uvs, mesh.tris, mesh.texture, rot, S, S)
    assert np.array_equal(tri[views], ref_tri)           # north_star asks >= 99.9 %; the frozen rules give 100 %
    assert np.array_equal(z[views], ref_z)
    cover = (tri >= 0).mean()
    assert 0.15 < cover < 0.6                             # the face fills a plausible part of every view
    assert tri.max() < len(mesh.tris)


def test_fullsize_cnn_determinism_and_fused_argmax(setup):
    dm, mesh, tr = setup
    dmesh = dm.renderer_3d.upload(mesh)
    u8 = dm.renderer_3d.render_device(dmesh, tr)["u8"]
    net = dm.predictor_2d.network(V, S, S)
    p_graph = net.forward(u8, graph=True)[0].clone()
    p_graph2 = net.forward(u8, graph=True)[0].clone()
    p_plain, hm = net.forward(u8, want_heatmaps=True)
    torch.cuda.synchronize()
    assert torch.equal(p_graph, p_graph2) and torch.equal(p_graph, p_plain)
    assert torch.isfinite(hm).all()
    flat = hm.view(V, 73, -1)
    val, idx = flat.max(-1)
    assert torch.equal(p_plain[..., 0], (idx // S).T.float() - 1.0)       # row - 1   (paulsenpredictor.py:127)
    assert torch.equal(p_plain[..., 1], (idx % S).T.float() - 0.5)        # col - 0.5
    assert torch.equal(p_plain[..., 2], val.T)
    # first maximum in row-major order, like np.argmax: no earlier pixel holds the same value
    first = (flat == val.unsqueeze(-1)).float().argmax(-1)
    assert torch.equal(first, idx)


def test_fullsize_landmarks_on_surface_and_repeatable(setup):
    dm, mesh, tr = setup
    a = dm.predict_mesh(mesh)
    b = dm.predict_mesh(mesh)
    assert a.shape == (73, 3) and np.isfinite(a).all() and np.array_equal(a, b)
    again, _ = native.snap_to_mesh(mesh.verts, mesh.tris, a)
    assert np.abs(again - a).max() <= 1e-6
    lo, hi = mesh.verts.min(0), mesh.verts.max(0)
    assert (a >= lo - 1e-6).all() and (a <= hi + 1e-6).all()
