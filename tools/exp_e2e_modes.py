"""Experiment: blocking predict_mesh loop vs pipelined predict_meshes, alternating, same process (headline workload)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import build, synth
from mvlm_b200.io_obj import Mesh
from mvlm_b200.pipeline import create_pipeline
from mvlm_b200.weights import seeded_state_dict
build.build()
v, uv, t = synth.face_mesh(grid=224, seed=1234)
tex = synth.face_texture(1024, seed=1234)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
mesh = Mesh(verts=pin(v), tris=pin(t), uvs=pin(uv), texture=pin(tex))
tr = synth.random_view_transforms(100, seed=1234)
dm = create_pipeline("dtu3d", n_views=100, weights=seeded_state_dict(73, "RGB+depth", 1234), seed=1234, verbose=False, image_size=(256, 256), transforms=tr)
for _ in range(3): dm.predict_mesh(mesh)
n = 30
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): dm.predict_mesh(mesh)
    torch.cuda.synchronize(); a = n / (time.perf_counter() - t0)
    res = []
    for depth in (1, 2, 3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dm.predict_meshes([mesh] * n, depth=depth)
        torch.cuda.synchronize(); res.append(n / (time.perf_counter() - t0))
    print(f"rep {rep}: blocking {a:.1f} | pipelined depth 1/2/3: {res[0]:.1f} {res[1]:.1f} {res[2]:.1f} scans/s")
