"""Dataflow segments (csrc/conv_flow.cu) vs the per-layer plan (csrc/conv_umma.cu).

Both run the same tiles through the same tcgen05 pipeline and the same epilogue code
(csrc/conv_epilogue.cuh), so every intermediate tensor, the heat maps and the fused peaks must be
BIT-identical; only the order in which tiles of different layers / views are executed differs.  The
per-layer plan is in turn checked against the reference-pinned oracle in tests/test_hourglass_gpu.py.
"""
import os

import pytest
import torch

from mvlm_b200.weights import IMAGE_CHANNELS, seeded_state_dict

pytestmark = pytest.mark.gpu

PROBES = ("x1", "y3", "r3", "hg1", "sum_temp", "x10")


def _build(env, *args):
    from mvlm_b200 import ops

    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return ops.Hourglass(*args, keep_probes=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("n_landmarks,mode,size,views,lo,hi,tiles,k", [
    (73, "RGB+depth", 128, 7, 1, 32, 64, 3),      # the default window: hourglass levels of <= 32 rows (tiles of 32 .. 4 rows)
    (73, "RGB+depth", 128, 7, 32, 128, 64, 3),    # the large layers as segments, ragged last batch
    (84, "geometry+depth", 128, 3, 1, 1024, 16, 2),  # everything the kernel supports in segments
    (73, "RGB+depth", 256, 5, 1, 32, 64, 3),      # the headline geometry (segment layout of the 100-view plan)
    (73, "RGB", 256, 4, 64, 256, 128, 1),         # no interleave: every group waits for the one right before it
])
def test_flow_equals_per_layer_plan(lib, n_landmarks, mode, size, views, lo, hi, tiles, k):
    sd = seeded_state_dict(n_landmarks, mode, seed=1234)
    cin = IMAGE_CHANNELS[mode]
    g = torch.Generator().manual_seed(11)
    img = torch.randint(0, 256, (views, size, size, 4), generator=g, dtype=torch.uint8)
    img[..., cin:] = 0
    img = img.cuda()
    args = (sd, n_landmarks, cin, views, size, size)
    ref = _build({"MVLM_FLOW": "0"}, *args)
    flow = _build({"MVLM_FLOW": "1", "MVLM_FLOW_LO": str(lo), "MVLM_FLOW_HI": str(hi), "MVLM_FLOW_TILES": str(tiles),
                   "MVLM_FLOW_K": str(k)}, *args)
    assert flow.num_segments > 0 and ref.num_segments == 0
    pk_ref, hm_ref = ref.forward(img, want_heatmaps=True)
    for rep in range(3):  # counters are reset per launch: repeated calls must agree too
        pk, hm = flow.forward(img, want_heatmaps=True)
        torch.cuda.synchronize()
        for name in PROBES:
            a, b = ref.probe(name), flow.probe(name)
            assert torch.equal(a.view(torch.int16), b.view(torch.int16)), (name, rep)
        assert torch.equal(hm.view(torch.int32), hm_ref.view(torch.int32)), rep
        assert torch.equal(pk.view(torch.int32), pk_ref.view(torch.int32)), rep
    # CUDA-graph replay of the segmented plan
    pk_g, _ = flow.forward(img, graph=True)
    pk_g2, _ = flow.forward(img, graph=True)
    torch.cuda.synchronize()
    assert torch.equal(pk_g2.view(torch.int32), pk_ref.view(torch.int32))


def test_workspace_packing_is_bit_identical_and_small(lib):
    """Buffer reuse (hourglass.cu, ws_alloc / assign_offsets) changes addresses only: peaks and heat maps of the packed
    plan equal those of the plan where every buffer has its own memory; the headline plan fits 6 GB."""
    from mvlm_b200 import ops

    assert lib.mvlm_hourglass_workspace_bytes(73, 4, 100, 256, 256) <= 6e9
    sd = seeded_state_dict(73, "RGB+depth", seed=1234)
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (6, 128, 128, 4), generator=g, dtype=torch.uint8).cuda()
    args = (sd, 73, 4, 6, 128, 128)
    flat = _build({"MVLM_HG_NO_REUSE": "1"}, *args)
    packed = ops.Hourglass(*args)
    assert packed.workspace.numel() < 0.35 * flat.workspace.numel()
    pk0, hm0 = flat.forward(img, want_heatmaps=True)
    for _ in range(2):
        pk1, hm1 = packed.forward(img, want_heatmaps=True)
        torch.cuda.synchronize()
        assert torch.equal(pk0.view(torch.int32), pk1.view(torch.int32))
        assert torch.equal(hm0.view(torch.int32), hm1.view(torch.int32))
    with pytest.raises(Exception):
        packed.probe("r3")


@pytest.mark.parametrize("n_landmarks,mode,size,views", [(73, "RGB+depth", 128, 3), (84, "geometry+depth", 64, 2),
                                                        (73, "RGB", 256, 2)])
def test_fused_pool_and_act_copies_equal_standalone_passes(lib, n_landmarks, mode, size, views):
    """The hourglass max-pools (paulsenpredictor.py:309-329) and the BatchNorm+ReLU copy of conv4's resample branch
    (:263-265) written by their producers' epilogues (ConvEpilogue::aux_mode, opt-in: it is not faster) == the stand-alone
    pool2_act / bn_relu passes of the default plan: every probe tensor, the heat maps and the peaks bit-identical."""
    sd = seeded_state_dict(n_landmarks, mode, seed=77)
    cin = IMAGE_CHANNELS[mode]
    g = torch.Generator().manual_seed(13)
    img = torch.randint(0, 256, (views, size, size, 4), generator=g, dtype=torch.uint8)
    img[..., cin:] = 0
    img = img.cuda()
    args = (sd, n_landmarks, cin, views, size, size)
    unfused = _build({"MVLM_HG_ELT_FUSION": "0"}, *args)
    fused = _build({"MVLM_HG_ELT_FUSION": "1"}, *args)
    assert fused.num_launches == unfused.num_launches - 11   # ten max-pools and one BatchNorm+ReLU pass are gone
    pk0, hm0 = unfused.forward(img, want_heatmaps=True)
    pk1, hm1 = fused.forward(img, want_heatmaps=True)
    torch.cuda.synchronize()
    for name in PROBES:
        assert torch.equal(unfused.probe(name).view(torch.int16), fused.probe(name).view(torch.int16)), name
    assert torch.equal(hm0.view(torch.int32), hm1.view(torch.int32))
    assert torch.equal(pk0.view(torch.int32), pk1.view(torch.int32))
    # the dataflow kernel runs the same variants (low levels in segments)
    flow = _build({"MVLM_HG_ELT_FUSION": "1", "MVLM_FLOW": "1", "MVLM_FLOW_LO": "1", "MVLM_FLOW_HI": "32"}, *args)
    pk2, hm2 = flow.forward(img, want_heatmaps=True)
    torch.cuda.synchronize()
    assert torch.equal(hm0.view(torch.int32), hm2.view(torch.int32)) and torch.equal(pk0.view(torch.int32), pk2.view(torch.int32))
