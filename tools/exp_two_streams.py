"""Experiment: two scans in flight on two CUDA streams (two plans / workspaces) vs one, headline workload."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import build, ops, synth  # noqa: E402
from mvlm_b200.io_obj import Mesh  # noqa: E402
from mvlm_b200.pipeline import create_pipeline  # noqa: E402
from mvlm_b200.utils.render3d import rotation_matrices  # noqa: E402
from mvlm_b200.weights import seeded_state_dict  # noqa: E402

build.build()
V, S, L = 100, 256, 73
n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = 40
v, uv, t = synth.face_mesh(grid=224, seed=1234)
mesh = Mesh(verts=v, tris=t, uvs=uv, texture=synth.face_texture(1024, seed=1234))
sd = seeded_state_dict(L, "RGB+depth", 1234)
tr = synth.random_view_transforms(V, seed=1234)
dev = torch.device("cuda")
lanes = []
for i in range(n_streams):
    dm = create_pipeline("dtu3d", n_views=V, weights=sd, seed=1234, n_hypotheses=1, verbose=False, image_size=(S, S), transforms=tr)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        net = dm.predictor_2d.network(V, S, S)
        dmesh = dm.renderer_3d.upload(mesh)
        rot = torch.from_numpy(rotation_matrices(tr).reshape(-1, 9)).to(dev)
        draws = torch.from_numpy(dm.estimator_3d.seeded_draws(L).view(np.int32)).to(dev)
        zbuf = None
        u8 = torch.empty((V, S, S, 4), dtype=torch.uint8, device=dev)
    lanes.append(dict(dm=dm, st=st, net=net, dmesh=dmesh, rot=rot, draws=draws, zbuf=zbuf, u8=u8))


def step(ln):
    with torch.cuda.stream(ln["st"]):
        ops.raster_multiview(ln["dmesh"].verts, ln["dmesh"].uvs, ln["dmesh"].tris, ln["dmesh"].tex, ln["rot"], S, S, "RGB+depth",
                             zbuf=ln["zbuf"], out_u8=ln["u8"])
        peaks, _ = ln["net"].forward(ln["u8"], graph=True)
        s, e = ops.rays_from_peaks(peaks, ln["rot"], S)
        lm, err, _ = ops.consensus(peaks, s, e, ln["draws"])
        out, _ = ops.snap_to_mesh(ln["dmesh"].verts, ln["dmesh"].tris, lm)
    return out


for _ in range(3):
    for ln in lanes:
        step(ln)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(steps):
    step(lanes[i % n_streams])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{n_streams} stream(s): {steps / dt:.2f} scans/s ({dt / steps * 1e3:.2f} ms/scan)")
