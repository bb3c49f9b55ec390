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


def test_fullsize_cnn_matches_oracle_on_sampled_views(setup):
    """The 100-view plan against the reference-pinned oracle (oracle/hourglass_ref.py, pinned to MVLMModel by
    tests/golden/cnn_*.npz) on views 0, 57 and 99: the plan takes code paths the small parity cases never see (32-row
    tiles everywhere, arg-max tile ranges spanning images, workspace offsets beyond 2^31 bytes, packed buffers).
    Bars = DESIGN.md section 5: vs the fp32 oracle mean|err| <= 1.2 % and max|err| <= 12 % of the heat maps' std, and no
    worse than 1.25 x an ideal same-rounding-points bf16 evaluation."""
    from oracle.hourglass_ref import HourglassOracle

    dm, mesh, tr = setup
    dmesh = dm.renderer_3d.upload(mesh)
    out = dm.renderer_3d.render_device(dmesh, tr, want_f32=True)
    net = dm.predictor_2d.network(V, S, S)
    _, hm = net.forward(out["u8"], want_heatmaps=True)
    torch.cuda.synchronize()
    views = [0, 57, 99]
    x = out["f32"][views].permute(0, 3, 1, 2).contiguous().cpu()
    sd = dm.predictor_2d._state_dict
    ref32 = HourglassOracle(sd).forward(x)
    ref16 = HourglassOracle(sd, emulate_bf16=True).forward(x)
    got = hm[views].cpu()
    std = ref32.std().item()
    e32 = (got - ref32).abs()
    ideal = (ref16 - ref32).abs().mean().item()
    print(f"full-size heat maps (views {views}): std {std:.3f}, vs fp32 oracle max {e32.max().item():.4f} mean {e32.mean().item():.5f}; "
          f"ideal bf16 evaluation mean {ideal:.5f}")
    assert e32.max().item() <= 0.12 * std and e32.mean().item() <= 0.012 * std, (e32.max().item(), e32.mean().item(), std)
    assert e32.mean().item() <= 1.25 * ideal + 1e-4 * std
    # ORACLE heat maps through the standalone CUDA peak kernel at 256^2: index and value bit-exact (R5), both methods'
    # arg-max part; "moment" within 1e-4 px of the numpy restatement
    from mvlm_b200 import ops

    ref_np = ref32.numpy()
    pk = ops.heatmap_peaks(ref32.cuda(), "simple").cpu().numpy()
    assert np.array_equal(pk, stages.heatmap_peaks(ref_np, "simple"))
    pm = ops.heatmap_peaks(ref32.cuda(), "moment").cpu().numpy()
    want = stages.heatmap_peaks(ref_np, "moment")
    assert np.array_equal(pm[..., 2], want[..., 2]) and np.abs(pm[..., :2] - want[..., :2]).max() <= 1e-4
