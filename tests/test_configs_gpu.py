"""BASELINE.json configs as parity cases (bench.py only times configs[2]):
  C2  BU3DFE geometry+depth pipeline (84 landmarks)            -> teacher-forced end to end
  C4  one scan, 512^2 views split over ranks + ray all-gather   -> view-split path == single-rank path
  C5  consensus micro-benchmark: 84 landmarks x 200 views, 30 % outliers, 16 384 seeded hypotheses
"""
import os

import numpy as np
import pytest
import torch

from mvlm_b200 import synth
from mvlm_b200.io_obj import Mesh
from mvlm_b200.weights import seeded_state_dict
from oracle import native, stages

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _mesh(grid=80, seed=3):
    v, uv, t = synth.face_mesh(grid=grid, seed=seed)
    return Mesh(verts=v, tris=t, uvs=uv, texture=synth.face_texture(256, seed=seed))


def test_c2_bu3dfe_geometry_depth(lib):
    import mvlm

    mesh = _mesh()
    tr = synth.random_view_transforms(10, seed=8)
    dm = mvlm.pipeline.create_pipeline("bu3dfe", n_views=10, weights=seeded_state_dict(84, "geometry+depth", 5), seed=3,
                                       n_hypotheses=4, verbose=False, image_size=(128, 128), transforms=tr,
                                       channel_mode="geometry+depth")
    lm = dm.predict_mesh(mesh)
    assert lm.shape == (84, 3) and np.isfinite(lm).all()
    dmesh = dm.renderer_3d.upload(mesh)
    out = dm.renderer_3d.render_device(dmesh, tr, want_f32=True)
    ref_img, _, _ = native.raster_multiview(mesh.verts, mesh.uvs, mesh.tris, mesh.texture, stages.rotation_matrices(tr),
                                            128, 128, "geometry+depth")
    assert np.array_equal(out["f32"].cpu().numpy(), ref_img)        # 2-channel (shade, depth) stack, bit-exact
    peaks = dm.predictor_2d.predict_landmarks_device(out["u8"]).cpu().numpy()
    s, e = stages.landmark_lines(128, peaks, tr)
    ref_lm, _, _ = stages.landmarks_from_lines(peaks, s, e, dm.estimator_3d.seeded_draws(84))
    ref, _ = native.snap_to_mesh(mesh.verts, mesh.tris, ref_lm)
    assert np.abs(ref - lm).max() <= 1e-3 * mesh.bbox_diagonal


def test_c4_view_split_equals_single_rank(lib):
    """World size 1 here (one GPU); tools/run_view_split.py runs the same check on 2+ GPUs over NCCL."""
    import torch.distributed as dist

    import mvlm
    from mvlm_b200.sharding import predict_mesh_view_split

    mesh = _mesh(grid=60)
    tr = synth.random_view_transforms(6, seed=2)
    dm = mvlm.pipeline.create_pipeline("dtu3d", n_views=6, weights=seeded_state_dict(73, "RGB+depth", 5), seed=3,
                                       n_hypotheses=4, verbose=False, image_size=(512, 512), transforms=tr)
    single = dm.predict_mesh(mesh)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        split = predict_mesh_view_split(dm, mesh, tr)
    finally:
        if created:
            dist.destroy_process_group()
    assert np.array_equal(single, split)


def test_view_split_key_handoff_equals_plain_peaks(lib):
    """The view-split hand-off on one GPU: three ranks' blocks of views (3 + 2 + 2 of 7) written by forward_keys into the
    slots of the gather buffer, one kernel from the gathered keys -> peaks of all views, bit-identical to the peaks of one
    plan over all 7 views (a view's CNN result does not depend on which other views share its launch)."""
    from mvlm_b200 import ops
    from mvlm_b200.sharding import split_views

    sd = seeded_state_dict(73, "RGB+depth", seed=7)
    g = torch.Generator().manual_seed(2)
    v, size, world = 7, 64, 3
    img = torch.randint(0, 256, (v, size, size, 4), generator=g, dtype=torch.uint8).cuda()
    whole, _ = ops.Hourglass(sd, 73, 4, v, size, size).forward(img)
    slot = split_views(v, 0, world)[1]
    kb = torch.full((world, slot, 73), -1, dtype=torch.int64, device="cuda")  # garbage in the rows no rank owns
    nets = {}
    for r in range(world):
        s0, c = split_views(v, r, world)
        net = nets.setdefault(c, ops.Hourglass(sd, 73, 4, c, size, size))
        for _ in range(2):  # second call replays the CUDA graph captured for this key buffer
            net.forward_keys(img[s0:s0 + c].contiguous(), kb[r, :c])
    peaks = ops.peaks_from_gathered_keys(kb, v, size)
    torch.cuda.synchronize()
    assert torch.equal(peaks.view(torch.int32), whole.view(torch.int32))


def test_c5_consensus_microbench_parity(lib):
    from mvlm_b200 import ops

    peaks, starts, ends, truth = synth.synthetic_rays(n_landmarks=84, n_views=200, outlier_frac=0.3, seed=1234)
    draws = synth.hypothesis_table(84, 16384, seed=1234)
    lm, err, nl = ops.consensus(cuda(peaks), cuda(starts), cuda(ends), cuda(draws.view(np.int32)))
    lm, err, nl = lm.cpu().numpy(), err.cpu().numpy(), nl.cpu().numpy()
    assert (nl == 100).all()                                  # median filter keeps exactly half of 200 distinct values
    # full-size oracle on a landmark subset (numpy evaluates ~5 000 hypotheses/s)
    for l in (0, 41, 83):
        ref_lm, _, ref_err = stages.landmarks_from_lines(peaks[l:l + 1], starts[l:l + 1], ends[l:l + 1], draws[l:l + 1])
        assert np.abs(lm[l] - ref_lm[0]).max() <= 1e-8
        assert abs(err[l] - ref_err[0]) <= 1e-9 * max(1.0, ref_err[0])
    # all landmarks at H = 64: the first 64 hypotheses only
    lm64, err64, _ = ops.consensus(cuda(peaks), cuda(starts), cuda(ends), cuda(np.ascontiguousarray(draws[:, :64]).view(np.int32)))
    ref_lm, _, ref_err = stages.landmarks_from_lines(peaks, starts, ends, draws[:, :64])
    assert np.abs(lm64.cpu().numpy() - ref_lm).max() <= 1e-8
    assert np.allclose(err64.cpu().numpy(), ref_err, rtol=1e-9, atol=1e-9)
    # size-independent properties at full size: more hypotheses never increase the error; robust to 30 % outliers
    assert (err <= err64.cpu().numpy() + 1e-12).all()
    assert np.abs(lm - truth).max() < 1.0                     # mm, rays carry 0.5 mm jitter


def test_view_axis_slicing_equals_one_launch(lib, monkeypatch):
    """Stacks beyond one plan's limits (32-bit pixel indices / half of the free memory) run as slices of the view axis:
    same peaks.  With the packed workspace 100 views of 512^2 (17.6 GB) are ONE plan on a B200."""
    import mvlm
    from mvlm_b200.prediction.paulsenpredictor import PaulsenModel

    mesh = _mesh(grid=60)
    tr = synth.random_view_transforms(12, seed=4)
    dm = mvlm.pipeline.create_pipeline("dtu3d", n_views=12, weights=seeded_state_dict(73, "RGB+depth", 5), seed=3,
                                       verbose=False, image_size=(64, 64), transforms=tr)
    dmesh = dm.renderer_3d.upload(mesh)
    u8 = dm.renderer_3d.render_device(dmesh, tr)["u8"]
    whole = dm.predictor_2d.predict_landmarks_device(u8).clone()
    assert dm.predictor_2d.max_views_per_launch(100, 512, 512) == 100     # was 25 with the unpacked 20.6 GB / 100-view layout
    assert dm.predictor_2d.max_views_per_launch(40000, 512, 512) <= (2 ** 32 - 1) // (512 * 512)
    monkeypatch.setattr(PaulsenModel, "max_views_per_launch", lambda self, v, h, w: min(v, 4))
    dm.predictor_2d._nets.clear()                                          # forget the 12-view plan: slices of 4 now
    sliced = dm.predictor_2d.predict_landmarks_device(u8)
    assert torch.equal(whole, sliced)
    assert list(dm.predictor_2d._nets) == [(4, 64, 64)]
