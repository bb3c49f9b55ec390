"""Where does a config-4 scan spend its time on one rank?  (2M triangles, 25 views of 512^2 = one rank's share of 200)"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import build, ops, synth
from mvlm_b200.io_obj import Mesh
from mvlm_b200.pipeline import create_pipeline
from mvlm_b200.weights import seeded_state_dict
build.build()
views, size, grid = 25, 512, 1001
v, uv, t = synth.face_mesh(grid=grid, seed=1234)
mesh = Mesh(verts=v, tris=t, uvs=uv, texture=synth.face_texture(1024, seed=1234))
tr = synth.random_view_transforms(views, seed=77)
dm = create_pipeline("dtu3d", n_views=views, weights=seeded_state_dict(73, "RGB+depth", 1234), seed=5, n_hypotheses=8, verbose=False,
                     image_size=(size, size), transforms=tr)
r, p, e = dm.renderer_3d, dm.predictor_2d, dm.estimator_3d
def timed(name, fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); print(f"{name:28s} {(time.perf_counter()-t0)/n*1e3:8.2f} ms"); return out
for _ in range(4):  # every slot of the pinned staging ring allocates on first use
    r.upload(mesh)
dmesh = timed("upload (47 MB)", lambda: r.upload(mesh), n=8)
out = timed("raster 2M tris x 25 views", lambda: r.render_device(dmesh, tr))
peaks = timed("CNN 25 x 512^2", lambda: p.predict_landmarks_device(out["u8"]))
rays = timed("rays (host rotations)", lambda: e.estimate_landmark_lines_device(peaks, tr, size))
draws = timed("draws", lambda: torch.from_numpy(e.seeded_draws(73).view(np.int32)).cuda())
lm = timed("consensus", lambda: e.estimate_landmarks_from_lines_device(peaks, rays[0], rays[1], draws))
timed("snap 73 x 2M tris (scan)", lambda: ops.snap_to_mesh(dmesh.verts, dmesh.tris, lm[0]))
timed("snap 73 x 2M tris (grid build + query)", lambda: ops.SnapGrid(dmesh.verts, dmesh.tris).query(lm[0]), n=10)
g = timed("  grid alloc + build", lambda: ops.SnapGrid(dmesh.verts, dmesh.tris), n=10)
_, _, st = g.query(lm[0], want_stats=True)
print("  landmarks handed to the scan (random-weight landmarks are not near the surface):", int((st[:, 1] < 0).sum()), "of", len(st))
timed("  grid query (these landmarks)", lambda: g.query(lm[0]), n=10)
near = dmesh.verts[::13699][:73].double() + 0.5
timed("  grid query (landmarks 0.9 mm off the surface)", lambda: g.query(near), n=10)
timed("  scan (landmarks 0.9 mm off the surface)", lambda: ops.snap_to_mesh(dmesh.verts, dmesh.tris, near), n=10)
from mvlm_b200.utils import render3d
render3d._PARALLEL_COPY_BYTES = 1 << 40
timed("upload, single-thread staging copy", lambda: r.upload(mesh), n=8)
render3d._PARALLEL_COPY_BYTES = 4 << 20
timed("upload, 4-thread staging copy", lambda: r.upload(mesh), n=8)
timed("predict_mesh total", lambda: dm.predict_mesh(mesh))
