"""Per-stage timing and roofline fraction of the non-CNN stages (SURVEY.md 8d algorithmic bytes).
GPU box:  python tools/bench_stages.py  -> one JSON line per stage (CUDA events, 3 warm-ups, 20 timed launches)."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from mvlm_b200 import build, ops, synth  # noqa: E402
from mvlm_b200.utils.render3d import rotation_matrices  # noqa: E402

build.build()
pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
HBM = pk["hbm_gbs"]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(stage, ms, alg_bytes, note=""):
    gbs = alg_bytes / ms / 1e6
    print(json.dumps({"stage": stage, "ms": round(ms, 4), "algorithmic_MB": round(alg_bytes / 1e6, 2), "achieved_GBps": round(gbs, 1),
                      "hbm_peak_GBps": HBM, "frac": round(gbs / HBM, 4), "note": note}), flush=True)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


V, S, L = 100, 256, 73
verts, uvs, tris = synth.face_mesh(224, 1234)
tex = synth.face_texture(1024, 1234)
tr = synth.random_view_transforms(V, 1234)
rot = cuda(rotation_matrices(tr).reshape(-1, 9))
dv, du, dt, dx = cuda(verts), cuda(uvs), cuda(tris), cuda(tex)
zbuf = ops.raster_workspace(V, S, S, dv.shape[0], "cuda")
u8 = torch.empty((V, S, S, 4), dtype=torch.uint8, device="cuda")
ms = timeit(lambda: ops.raster_multiview(dv, du, dt, dx, rot, S, S, zbuf=zbuf, out_u8=u8))
cov = float((ops.raster_multiview(dv, du, dt, dx, rot, S, S, want_tri=True)["tri"] >= 0).float().mean())
alg = verts.nbytes + uvs.nbytes + tris.nbytes + V * S * S * (8 + 8 + 4) + cov * V * S * S * 3
report("raster (all views)", ms, alg, f"zbuf clear+atomics+resolve, u8 NHWC4 out; coverage {cov:.3f}; practical limiter is atomics/ALU")

hm = torch.randn((V, L, S, S), device="cuda")
ms = timeit(lambda: ops.heatmap_peaks(hm, "simple"), 10)
report("peaks standalone (simple)", ms, hm.numel() * 4, "fused path: 0 extra bytes (arg-max in the conv11 epilogue)")
ms = timeit(lambda: ops.heatmap_peaks(hm, "moment"), 10)
report("peaks standalone (moment)", ms, hm.numel() * 4)
del hm

peaks = torch.rand((L, V, 3), device="cuda")
ms = timeit(lambda: ops.rays_from_peaks(peaks, rot, S))
report("rays", ms, L * V * (12 + 48))

pk5, st5, en5, _ = synth.synthetic_rays(84, 200, 0.3, 1234)
for H in (1, 16384):
    draws = cuda(synth.hypothesis_table(84, H, 1234).view(np.int32))
    a, b, c = cuda(pk5), cuda(st5), cuda(en5)
    ms = timeit(lambda: ops.consensus(a, b, c, draws), 10)
    report(f"consensus C5 (84 lm x 200 views, H={H})", ms, 84 * 200 * 60 + 84 * H * 32,
           f"{84 * H * (200 + 100) / ms / 1e6:.1f} M point-line distances/ms; ALU(fp64)-bound")

lm = cuda(np.random.RandomState(0).uniform(-80, 80, (L, 3)))
ms = timeit(lambda: ops.snap_to_mesh(dv, dt, lm))
report("snap (73 lm x 99k tris)", ms, tris.nbytes + verts.nbytes, f"{L * len(tris) / ms / 1e6:.1f} M point-triangle tests/ms; mesh L2-resident")
