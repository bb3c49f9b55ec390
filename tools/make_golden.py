"""Generates tests/golden/*.npz by running the REFERENCE's own code (imported in place from
/root/reference, see oracle/ref_loader.py) on seeded inputs.  Run in the build container only:

    python tools/make_golden.py

The reference ships no tests or fixtures of its own; these vectors are the pin
for the CPU oracle (tests/test_oracle_golden.py), which in turn is the checker for the CUDA path.
Stages that run inside VTK (renderer, surface snap) cannot be executed here and have no golden.
"""
from __future__ import annotations

import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from mvlm_b200 import synth  # noqa: E402
from mvlm_b200.weights import seeded_state_dict  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = ROOT / "tests" / "golden"


def cnn_golden(tag, n_landmarks, mode, size, seed):
    sd = seeded_state_dict(n_landmarks, mode, seed)
    model = ref_loader.make_model(n_landmarks, mode, sd)
    cin = model.in_channels
    g = torch.Generator().manual_seed(seed + 1)
    img_u8 = torch.randint(0, 256, (2, size, size, cin), generator=g, dtype=torch.uint8)
    x = (img_u8.float() / 255.0).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        out = model(x)          # (2 stages, B, L, H, W)
    hm = out[-1].numpy()        # what predict_landmarks_from_images consumes (:204-205)
    keep = sorted({0, 1, n_landmarks // 2, n_landmarks - 1})
    flat = hm.reshape(hm.shape[0], hm.shape[1], -1)
    np.savez_compressed(OUT / f"cnn_{tag}.npz", n_landmarks=n_landmarks, mode=mode, seed=seed, img_u8=img_u8.numpy(),
                        channels=np.array(keep), heatmaps_subset=hm[:, keep].astype(np.float32),
                        argmax=flat.argmax(-1).astype(np.int64), maxval=flat.max(-1).astype(np.float32),
                        mean=hm.mean((2, 3)).astype(np.float32), std=np.float32(hm.std()))
    print(tag, hm.shape, "std", hm.std())


def stages_golden():
    ns = ref_loader.load()
    pp, est_mod, u3d = ns.paulsenpredictor, ns.estimator3d, ns.utils3d
    rng = np.random.RandomState(7)
    # ---- peaks (find_maxima_in_batch_of_heatmaps, :160-165) -------------------------------------
    hm = (0.05 * rng.rand(3, 6, 64, 64)).astype(np.float32)
    yy, xx = np.mgrid[0:64, 0:64]
    for v in range(3):
        for l in range(6):
            cy, cx = rng.uniform(3, 61, 2)
            hm[v, l] += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 20.0).astype(np.float32)
    hm[0, 0] = 0.25
    hm[0, 1, 10, 5] = hm[0, 1, 40, 60] = 9.0
    hm[2, 3, 63, 63] = 50.0
    hm_nan = hm.copy()
    hm_nan[1, 2, 33, 17] = np.nan
    hm_nan[1, 2, 50, 1] = np.nan
    pred = pp.DTU3DPredictor.__new__(pp.DTU3DPredictor)
    res = {}
    for method in ("simple", "moment"):
        pred.selection_method = method
        for name, arr in (("", hm), ("_nan", hm_nan)):
            if method == "moment" and name == "_nan":
                continue
            out = np.empty((6, 3, 3), dtype=np.float32)
            pred.find_maxima_in_batch_of_heatmaps(torch.from_numpy(arr), out)
            res[f"peaks_{method}{name}"] = out
    # ---- rays (estimate_landmark_lines, estimator3d.py:31-90) ------------------------------------
    est = est_mod.Estimator3D()
    tr64 = synth.random_view_transforms(7, seed=5)
    tr32 = np.array([[30, 15, 0, 0, 0, 0], [30, -15, 0, 0, 0, 0], [30, 45, 0, 0, 0, 0], [30, -45, 0, 0, 0, 0],
                     [-30, 15, 0, 0, 0, 0], [-30, -15, 0, 0, 0, 0], [-30, 45, 0, 0, 0, 0], [-30, -45, 0, 0, 0, 0]],
                    dtype=np.float32)
    img = np.zeros((1, 256, 256, 4), np.float32)
    for name, tr in (("f64", tr64), ("f32", tr32)):
        pk = np.stack([rng.uniform(-1, 255, (5, len(tr))), rng.uniform(-0.5, 255.5, (5, len(tr))), rng.rand(5, len(tr))],
                      -1).astype(np.float32)
        s, e = est.estimate_landmark_lines(np.broadcast_to(img, (len(tr), 256, 256, 4)), pk, tr)
        res[f"rays_{name}_peaks"], res[f"rays_{name}_tr"] = pk, tr
        res[f"rays_{name}_starts"], res[f"rays_{name}_ends"] = s, e
    # ---- LSQ (utils3d.py:99-124) --------------------------------------------------------------------
    peaks, starts, ends, truth = synth.synthetic_rays(n_landmarks=6, n_views=40, outlier_frac=0.3, seed=3)
    res["lsq_pa"], res["lsq_pb"] = starts[0], ends[0]
    res["lsq_out"] = u3d.compute_intersection_between_lines(starts[0], ends[0])
    same = np.repeat(starts[0][:1], 8, 0), np.repeat(ends[0][:1], 8, 0)          # degenerate: one line 8x
    res["lsq_degenerate_out"] = u3d.compute_intersection_between_lines(*same)
    # ---- RANSAC (estimator3d.py:92-137) with an explicit hypothesis list --------------------------
    draws = synth.hypothesis_table(6, 12, seed=2)
    orig_choice = np.random.choice
    ran_p, ran_e = [], []
    for lm in range(6):
        pa, pb = starts[lm], ends[lm]
        best_p, best_e = None, None
        for h in range(draws.shape[1]):
            idx = (draws[lm, h].astype(np.uint64) % np.uint64(len(pa))).astype(np.int64)
            np.random.choice = lambda *a, **k: idx
            try:
                p, e = est.compute_intersection_between_lines_ransac(pa, pb)
            finally:
                np.random.choice = orig_choice
            if e != 100000000 and (best_e is None or e < best_e):
                best_p, best_e = p, e
        if best_p is None:
            best_p, best_e = u3d.compute_intersection_between_lines(pa, pb), 100000000
        ran_p.append(best_p)
        ran_e.append(best_e)
    res["ransac_starts"], res["ransac_ends"], res["ransac_draws"] = starts, ends, draws
    res["ransac_points"], res["ransac_errors"] = np.array(ran_p), np.array(ran_e, dtype=np.float64)
    # ---- estimate_landmarks_from_lines (:158-183), single reference draw, seeded global RNG -----
    for mode in ("quantile", "absolute"):
        est.mode = mode
        np.random.seed(123)
        with contextlib.redirect_stdout(io.StringIO()):
            lm_out, err = est.estimate_landmarks_from_lines(peaks, starts, ends)
        res[f"efl_{mode}_landmarks"], res[f"efl_{mode}_error"] = lm_out, np.float64(err)
    res["efl_peaks"] = peaks
    np.savez_compressed(OUT / "stages.npz", heatmaps=hm, heatmaps_nan=hm_nan, **res)
    print("stages:", sorted(res))


def views_golden():
    """View lists of the reference renderer (render3d.py:79-112), global numpy RNG seeded: the 8-view preset and the
    random lists the pipelines draw for any other n_views."""
    cls = ref_loader.load_renderer_class()
    res = {}
    r8 = cls(n_views=8)
    res["views_8"] = np.asarray(r8.generate_3d_transformations())
    for n, seed in ((5, 4), (100, 1234), (200, 77)):
        np.random.seed(seed)
        res[f"views_{n}_seed{seed}"] = np.asarray(cls(n_views=n).generate_3d_transformations())
    np.random.seed(11)
    res["random_transform_3_seed11"] = np.asarray(cls(n_views=3).random_transform(3))
    np.savez_compressed(OUT / "views.npz", **res)
    print("views:", {k: (v.shape, str(v.dtype)) for k, v in res.items()})


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not present: goldens can only be regenerated in the build container")
    OUT.mkdir(parents=True, exist_ok=True)
    cnn_golden("dtu3d_rgbd_64", 73, "RGB+depth", 64, 1234)
    cnn_golden("bu3dfe_geod_64", 84, "geometry+depth", 64, 99)
    stages_golden()
    views_golden()


if __name__ == "__main__":
    main()
