"""Config 4 on N GPUs: one scan whose views are split over the ranks, peaks all-gathered over NCCL.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_view_split.py [views] [size]
Every rank checks that the view-split landmarks equal the single-rank result bit for bit and rank 0 prints timings."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import build, synth  # noqa: E402
from mvlm_b200.io_obj import Mesh  # noqa: E402
from mvlm_b200.pipeline import create_pipeline  # noqa: E402
from mvlm_b200 import sharding  # noqa: E402
from mvlm_b200.sharding import predict_mesh_view_split, upload_mesh_sharded  # noqa: E402
from mvlm_b200.weights import seeded_state_dict  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 16
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 224
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
build.build()
v, uv, t = synth.face_mesh(grid=grid, seed=1234)
mesh = Mesh(verts=v, tris=t, uvs=uv, texture=synth.face_texture(1024, seed=1234))
tr = synth.random_view_transforms(views, seed=77)
sd = seeded_state_dict(73, "RGB+depth", 1234)
dm = create_pipeline("dtu3d", n_views=views, weights=sd, seed=5, n_hypotheses=8, verbose=False, image_size=(size, size),
                     transforms=tr, device=f"cuda:{local}")
for _ in range(5):  # plan creation, graph capture and the renderer's pinned staging slots are first-use costs
    split = predict_mesh_view_split(dm, mesh, tr)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    split = predict_mesh_view_split(dm, mesh, tr)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
ok, checked = True, False
# the sharded upload (1/world of the scan per rank + NVLink all-gather) delivers the host arrays bit for bit, and timing
sharded = sharding._mesh_bytes(mesh) >= sharding.SHARDED_UPLOAD_MIN_BYTES and world > 1
up = {}
for name, fn in (("sharded", lambda: upload_mesh_sharded(dm.renderer_3d, mesh)), ("whole", lambda: dm.renderer_3d.upload(mesh))):
    for _ in range(4):
        d = fn()
    torch.cuda.synchronize()
    dist.barrier()
    t1 = time.perf_counter()
    for _ in range(5):
        d = fn()
    torch.cuda.synchronize()
    up[name] = (time.perf_counter() - t1) / 5 * 1e3
    for a, b in ((d.verts, mesh.verts), (d.tris, mesh.tris), (d.uvs, mesh.uvs), (d.tex, mesh.texture)):
        ok = ok and bool(np.array_equal(a.cpu().numpy(), b))
if views * size * size <= 64 * 256 * 256:  # single-rank reference only when it fits comfortably
    single = dm.predict_mesh(mesh)
    ok, checked = ok and bool(np.array_equal(single, split)), True
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"view-split over {world} ranks: {views} views {size}^2, {len(t)} tris: {dt * 1e3:.2f} ms/scan, "
          + (f"identical to single rank: {bool(flag.item())}" if checked else "single-rank comparison skipped at this size")
          + f"; scan upload {sharding._mesh_bytes(mesh) / 1e6:.0f} MB: whole per rank {up['whole']:.2f} ms, 1/{world} per rank + all-gather "
          f"{up['sharded']:.2f} ms ({'used' if sharded else 'not used at this size'}), arrays identical: {bool(flag.item())}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
