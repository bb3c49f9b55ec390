"""GPU parity of the non-CNN stages against the CPU oracle (oracle/), through the C-ABI.

Bars: raster tri-ID / depth byte / texel bit-exact (north_star asks >= 99.9 % tri-ID agreement);
peak indices bit-exact; rays and fp64 consensus within stated absolute tolerances.
"""
import numpy as np
import pytest
import torch

from mvlm_b200 import synth
from oracle import native, stages

pytestmark = pytest.mark.gpu


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ------------------------------------------------------------------------------ raster
@pytest.mark.parametrize("size,mode", [(128, "RGB+depth"), (256, "RGB+depth"), (128, "geometry+depth"),
                                       (128, "RGB"), (128, "depth"), (64, "geometry")])
def test_raster_matches_oracle(lib, size, mode):
    from mvlm_b200 import ops

    verts, uvs, tris = synth.face_mesh(grid=120, seed=5)
    tex = synth.face_texture(256, seed=5)
    tr = synth.random_view_transforms(6, seed=3)
    tr[0, :3] = 0  # frontal view too
    rot = stages.rotation_matrices(tr)
    ref_img, ref_tri, ref_z = native.raster_multiview(verts, uvs, tris, tex, rot, size, size, mode)
    out = ops.raster_multiview(cuda(verts), cuda(uvs), cuda(tris), cuda(tex), cuda(rot.reshape(-1, 9)), size, size, mode,
                               want_f32=True, want_tri=True, want_z=True)
    torch.cuda.synchronize()
    tri = out["tri"].cpu().numpy()
    agree = (tri == ref_tri).mean()
    assert agree >= 0.999, agree
    assert agree == 1.0, f"tri-id maps differ on {(tri != ref_tri).sum()} pixels"
    assert np.array_equal(out["z"].cpu().numpy().view(np.uint32), ref_z.view(np.uint32))
    assert np.array_equal(out["f32"].cpu().numpy(), ref_img)
    # the packed u8 image the CNN stem reads is the same image, times 255
    c = ref_img.shape[-1]
    u8 = out["u8"].cpu().numpy()[..., :c]
    assert np.array_equal(u8, np.rint(ref_img * 255).astype(np.uint8))
    assert (ref_tri >= 0).mean() > 0.05  # the mesh is actually drawn


def test_raster_rgba_texture_and_view_chunks(lib):
    """RGBA texture (one 4-byte load per texel) == RGB texture; a million-vertex scan whose transformed vertices exceed the
    256 MB budget is rendered in chunks of views with the same result as the oracle (tri-ids, depth, image bit-exact)."""
    from mvlm_b200 import ops

    verts, uvs, tris = synth.face_mesh(grid=90, seed=9)
    tex = synth.face_texture(128, seed=9)
    tr = synth.random_view_transforms(5, seed=1)
    rot = stages.rotation_matrices(tr)
    args = (cuda(verts), cuda(uvs), cuda(tris))
    rgb = ops.raster_multiview(*args, cuda(tex), cuda(rot.reshape(-1, 9)), 128, 128, want_f32=True, want_tri=True)
    tex4 = np.concatenate([tex, np.full(tex.shape[:2] + (1,), 7, np.uint8)], axis=-1)
    rgba = ops.raster_multiview(*args, cuda(tex4), cuda(rot.reshape(-1, 9)), 128, 128, want_f32=True, want_tri=True)
    torch.cuda.synchronize()
    assert torch.equal(rgb["u8"], rgba["u8"]) and torch.equal(rgb["f32"], rgba["f32"]) and torch.equal(rgb["tri"], rgba["tri"])
    # 1 002 001 vertices x 20 views x 16 B = 320 MB > 256 MB -> two chunks of views
    verts, uvs, tris = synth.face_mesh(grid=1001, seed=2)
    tr = synth.random_view_transforms(20, seed=6)
    rot = stages.rotation_matrices(tr)
    assert lib.mvlm_raster_workspace_bytes(20, 64, 64, len(verts)) < 20 * 64 * 64 * 8 + 20 * len(verts) * 16
    out = ops.raster_multiview(cuda(verts), cuda(uvs), cuda(tris), cuda(tex), cuda(rot.reshape(-1, 9)), 64, 64,
                               want_f32=True, want_tri=True, want_z=True)
    torch.cuda.synchronize()
    ref_img, ref_tri, ref_z = native.raster_multiview(verts, uvs, tris, tex, rot, 64, 64)
    assert np.array_equal(out["tri"].cpu().numpy(), ref_tri)
    assert np.array_equal(out["z"].cpu().numpy().view(np.uint32), ref_z.view(np.uint32))
    assert np.array_equal(out["f32"].cpu().numpy(), ref_img)


def test_raster_untextured_is_white(lib):
    from mvlm_b200 import ops

    verts, uvs, tris = synth.face_mesh(grid=40, seed=2)
    rot = stages.rotation_matrices(np.zeros((1, 6)))
    out = ops.raster_multiview(cuda(verts), None, cuda(tris), None, cuda(rot.reshape(-1, 9)), 64, 64, want_f32=True,
                               want_tri=True)
    img = out["f32"].cpu().numpy()
    ref, ref_tri, _ = native.raster_multiview(verts, None, tris, None, rot, 64, 64)
    assert np.array_equal(img, ref)
    assert (img[..., :3] == 1.0).all()           # actor colour white, ambient only (utils3d.py:61-64)
    assert img[0, 0, 0, 3] == np.float32(1 / 255)  # background depth byte 1 (wrapped -255)


# ------------------------------------------------------------------------------ peaks
def test_peaks_simple_bit_exact(lib):
    from mvlm_b200 import ops

    rng = np.random.RandomState(0)
    hm = rng.randn(3, 7, 64, 64).astype(np.float32)
    hm[0, 0] = 0.25                      # all-equal map -> index 0 -> (-1, -0.5, v)
    hm[0, 1, 10, 5] = hm[0, 1, 40, 60] = 9.0  # tie -> first in row-major order
    hm[1, 2, 33, 17] = np.nan            # NaN wins
    hm[1, 2, 50, 1] = np.nan
    hm[2, 3, 63, 63] = 50.0              # last element
    ref = stages.heatmap_peaks(hm, "simple")
    got = ops.heatmap_peaks(cuda(hm), "simple").cpu().numpy()
    assert np.array_equal(got[..., :2], ref[..., :2])
    assert np.array_equal(got[..., 2].view(np.uint32), ref[..., 2].view(np.uint32))
    assert tuple(got[0, 0, :2]) == (-1.0, -0.5)
    assert tuple(got[1, 0, :2]) == (9.0, 4.5)
    assert tuple(got[2, 1, :2]) == (32.0, 16.5)


def test_peaks_moment(lib):
    from mvlm_b200 import ops

    rng = np.random.RandomState(1)
    hm = (0.05 * rng.rand(2, 5, 64, 64)).astype(np.float32)
    yy, xx = np.mgrid[0:64, 0:64]
    for v in range(2):
        for l in range(5):
            cy, cx = rng.uniform(5, 59, 2)
            hm[v, l] += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 18.0).astype(np.float32)
    ref = stages.heatmap_peaks(hm, "moment")
    got = ops.heatmap_peaks(cuda(hm), "moment").cpu().numpy()
    # index part bit-exact, sub-pixel moment within 1e-4 px (fp64 vs numpy's mixed fp32/fp64 sums)
    assert np.allclose(got[..., :2], ref[..., :2], atol=1e-4, rtol=0)
    assert np.array_equal(got[..., 2], ref[..., 2])
    assert np.abs(ref[..., :2] - stages.heatmap_peaks(hm, "simple")[..., :2]).max() > 0.05  # refinement happened


# ------------------------------------------------------------------------------ rays
@pytest.mark.parametrize("preset", [False, True])
def test_rays(lib, preset):
    from mvlm_b200 import ops
    from mvlm_b200.utils.render3d import fixed_eight_views

    rng = np.random.RandomState(2)
    tr = fixed_eight_views() if preset else synth.random_view_transforms(13, seed=9)
    v = tr.shape[0]
    peaks = np.stack([rng.uniform(-1, 255, (9, v)), rng.uniform(-0.5, 255.5, (9, v)), rng.rand(9, v)], -1).astype(np.float32)
    rs, re = stages.landmark_lines(256, peaks, tr)
    rot = stages.rotation_matrices(tr)
    gs, ge = ops.rays_from_peaks(cuda(peaks), cuda(rot.reshape(-1, 9)), 256)
    assert np.abs(gs.cpu().numpy() - rs).max() <= 1e-9
    assert np.abs(ge.cpu().numpy() - re).max() <= 1e-9


# ------------------------------------------------------------------------------ consensus
def _run_consensus(ops, peaks, starts, ends, draws, **kw):
    lm, err, nl = ops.consensus(cuda(peaks), cuda(starts), cuda(ends), cuda(draws.view(np.int32)), **kw)
    return lm.cpu().numpy(), err.cpu().numpy(), nl.cpu().numpy()


@pytest.mark.parametrize("n_hyp", [1, 8, 200])
@pytest.mark.parametrize("mode", ["quantile", "absolute"])
def test_consensus_matches_oracle(lib, n_hyp, mode):
    from mvlm_b200 import ops

    peaks, starts, ends, truth = synth.synthetic_rays(n_landmarks=11, n_views=50, outlier_frac=0.3, seed=4)
    draws = synth.hypothesis_table(11, n_hyp, seed=8)
    ref_lm, ref_mean, ref_err = stages.landmarks_from_lines(peaks, starts, ends, draws, mode=mode)
    lm, err, nl = _run_consensus(ops, peaks, starts, ends, draws, mode=mode)
    assert np.abs(lm - ref_lm).max() <= 1e-8, np.abs(lm - ref_lm).max()
    assert np.allclose(err, ref_err, rtol=1e-9, atol=1e-9)
    for l in range(11):
        m = stages.line_filter_mask(peaks[l, :, 2], mode, 0.5, 0.5)
        assert nl[l] == m.sum()
    if n_hyp >= 200:
        bbox = 160.0 * np.sqrt(3)
        assert np.abs(lm - truth).max() < 1e-2 * bbox  # robust estimate is near the truth


def test_consensus_edge_cases(lib):
    from mvlm_b200 import ops

    peaks, starts, ends, _ = synth.synthetic_rays(n_landmarks=6, n_views=9, outlier_frac=0.2, seed=6)
    peaks[0, :, 2] = 0.7                    # all equal -> quantile keeps none -> zeros, error 0
    peaks[1, :, 2] = 0.1; peaks[1, 3, 2] = 0.9; peaks[1, 5, 2] = 0.8   # absolute: 2 lines -> plain LSQ
    peaks[2, 4, 2] = np.nan                 # NaN -> np.quantile is NaN -> nothing kept
    draws = synth.hypothesis_table(6, 4, seed=1)
    for mode, q in (("quantile", 0.5), ("absolute", 0.5), ("quantile", 0.3), ("quantile", 0.77), ("quantile", 1.0), ("quantile", 0.0)):
        ref_lm, _, ref_err = stages.landmarks_from_lines(peaks, starts, ends, draws, mode=mode, threshold_quantile=q)
        lm, err, nl = _run_consensus(ops, peaks, starts, ends, draws, mode=mode, threshold_quantile=q)
        for l in range(6):
            assert nl[l] == stages.line_filter_mask(peaks[l, :, 2], mode, q, 0.5).sum(), (mode, q, l)
        assert np.abs(lm - ref_lm).max() <= 1e-8, (mode, q)
        assert np.allclose(err, ref_err, rtol=1e-9, atol=1e-9), (mode, q)
    with pytest.raises(ValueError):
        ops.consensus(cuda(peaks), cuda(starts), cuda(ends), cuda(draws.view(np.int32)), mode="bogus")


def test_lsq_recovers_concurrent_point(lib):
    """Closed-form property: noise-free concurrent lines meet in the point (reference probe: 4e-6)."""
    from mvlm_b200 import ops

    peaks, starts, ends, truth = synth.synthetic_rays(n_landmarks=20, n_views=30, outlier_frac=0.0, seed=12)
    # remove the jitter: rebuild rays exactly through the truth
    d = ends - starts
    starts = truth[:, None, :] - 0.5 * d
    ends = truth[:, None, :] + 0.5 * d
    peaks[:, :, 2] = 1.0
    draws = synth.hypothesis_table(20, 1, seed=0)
    lm, err, nl = _run_consensus(ops, peaks, starts, ends, draws, mode="absolute", threshold_absolute=0.5)
    assert (nl == 30).all()
    assert np.abs(lm - truth).max() < 1e-9
    assert err.max() < 1e-15


# ------------------------------------------------------------------------------ snap
def test_snap_matches_oracle(lib):
    from mvlm_b200 import ops

    verts, _, tris = synth.face_mesh(grid=90, seed=3)
    rng = np.random.RandomState(5)
    lm = rng.uniform(-100, 100, (37, 3))
    lm[:5] = verts[rng.randint(0, len(verts), 5)]  # exactly on vertices
    ref, ref_tri = native.snap_to_mesh(verts, tris, lm)
    out, tid = ops.snap_to_mesh(cuda(verts), cuda(tris), cuda(lm))
    out = out.cpu().numpy()
    assert np.abs(out - ref).max() <= 1e-9
    # idempotence: a snapped point snaps to itself
    out2, _ = ops.snap_to_mesh(cuda(verts), cuda(tris), cuda(out))
    assert np.abs(out2.cpu().numpy() - out).max() <= 1e-6
    d_ref = np.linalg.norm(ref - lm, axis=1)
    d_got = np.linalg.norm(out - lm, axis=1)
    assert np.allclose(d_ref, d_got, atol=1e-9)


def _snap_cases(rng, verts, n):
    """Landmarks near the surface (the pipeline's case), on vertices, inside the bounding box, far outside, non-finite."""
    ext = verts.max(0) - verts.min(0)
    near = verts[rng.randint(0, len(verts), n)] + rng.normal(0, 0.01, (n, 3)) * ext
    inside = rng.uniform(verts.min(0), verts.max(0), (n // 2, 3))
    far = rng.uniform(-30, 30, (8, 3)) * ext + verts.mean(0)
    on = verts[rng.randint(0, len(verts), 6)].astype(np.float64)
    odd = np.array([[np.nan, 0, 0], [np.inf, 0, 0], [1e30, -1e30, 0]])
    return np.concatenate([near, inside, far, on, odd]).astype(np.float64)


@pytest.mark.parametrize("case", ["face", "face_big_tris", "slivers", "two_tris", "degenerate", "duplicates"])
def test_snap_grid_equals_brute_force(lib, case):
    """the grid is only an accelerator (like vtkCellLocator, estimator3d.py:258-262) -- same
    triangle id and bit-identical point as the full scan, for every kind of landmark and mesh."""
    from mvlm_b200 import ops

    rng = np.random.RandomState(11)
    if case in ("face", "face_big_tris", "duplicates"):
        verts, _, tris = synth.face_mesh(grid=140, seed=4)
        if case == "face_big_tris":  # a few huge triangles -> oversize list
            extra = np.array([[-500, -500, -80], [500, -500, -80], [0, 600, -90], [300, 300, 300]], np.float32)
            tris = np.concatenate([tris, len(verts) + np.array([[0, 1, 2], [1, 2, 3]], np.int32)])
            verts = np.concatenate([verts, extra])
        if case == "duplicates":  # exact ties between coincident triangles -> lowest id must win
            tris = np.concatenate([tris, tris[::7]])
    elif case == "slivers":
        n = 4000
        a = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
        verts = np.concatenate([a, a + rng.normal(0, 30, (n, 3)).astype(np.float32), a + rng.normal(0, 0.01, (n, 3)).astype(np.float32)])
        tris = np.stack([np.arange(n), np.arange(n) + n, np.arange(n) + 2 * n], 1).astype(np.int32)
    elif case == "two_tris":
        verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.5]], np.float32)
        tris = np.array([[0, 1, 2], [1, 3, 2]], np.int32)
    else:  # all vertices identical / collinear: zero-area triangles, zero-extent bounding box
        verts = np.concatenate([np.zeros((3, 3), np.float32), np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2]], np.float32)])
        tris = np.array([[0, 1, 2], [3, 4, 5], [0, 0, 0]], np.int32)
    lm = _snap_cases(rng, verts, 60)
    dv, dt, dl = cuda(verts), cuda(tris), cuda(lm)
    ref, ref_tri = ops.snap_to_mesh(dv, dt, dl)
    grid = ops.SnapGrid(dv, dt)
    out, tid, stats = grid.query(dl, want_stats=True)
    info = grid.describe()
    finite = np.isfinite(lm).all(1)
    assert torch.equal(tid, ref_tri), (info, (tid != ref_tri).nonzero().flatten().tolist())
    assert np.array_equal(out.cpu().numpy()[finite].view(np.int64), ref.cpu().numpy()[finite].view(np.int64))
    st = stats.cpu().numpy()
    print(case, info, "tests/landmark: median", int(np.median(st[:, 0])), "max", int(st[:, 0].max()), "of", len(tris),
          "; full-scan fallbacks", int((st[:, 1] < 0).sum()))
    if case == "face":
        near = st[:60]
        assert (near[:, 1] >= 0).all() and np.median(near[:, 0]) < 0.05 * len(tris)  # the index actually prunes
    if case == "face_big_tris":
        assert info["n_oversize"] >= 2
    # CPU oracle agrees too
    o_ref, _ = native.snap_to_mesh(verts, tris, lm[finite])
    assert np.abs(out.cpu().numpy()[finite] - o_ref).max() <= 1e-9 * max(1.0, np.abs(lm[finite]).max())


def test_snap_grid_auto_selection(lib, monkeypatch):
    """ops.snap_to_mesh(grid="auto") and DeviceMesh.snap_grid() switch on the triangle count; same answer either way."""
    from mvlm_b200 import ops
    from mvlm_b200.io_obj import Mesh
    from mvlm_b200.utils.render3d import DeviceMesh

    verts, uvs, tris = synth.face_mesh(grid=60, seed=2)
    lm = cuda(verts[::97].astype(np.float64) + 0.3)
    dm = DeviceMesh(Mesh(verts=verts, uvs=uvs, tris=tris, texture=None), torch.device("cuda"))
    assert dm.snap_grid() is None  # small mesh: brute force
    a, ta = ops.snap_to_mesh(dm.verts, dm.tris, lm, grid="auto")
    monkeypatch.setattr(ops, "SNAP_GRID_MIN_TRIS", 1000)
    dm2 = DeviceMesh(dm.mesh, torch.device("cuda"))
    assert isinstance(dm2.snap_grid(), ops.SnapGrid) and dm2.snap_grid() is dm2.snap_grid()
    b, tb = ops.snap_to_mesh(dm2.verts, dm2.tris, lm, grid=dm2.snap_grid())
    c, tc = ops.snap_to_mesh(dm2.verts, dm2.tris, lm, grid="auto")
    assert torch.equal(a, b) and torch.equal(ta, tb) and torch.equal(a, c) and torch.equal(ta, tc)
