"""GPU box: where does SnapGrid() spend its time at 2M triangles (allocation vs the build launches)?"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvlm_b200 import ops, synth
v, _, t = synth.face_mesh(grid=1001, seed=1)
dv, dt = torch.from_numpy(v).cuda(), torch.from_numpy(t).cuda()
lib = ops._lib.load()
nb = lib.mvlm_snap_grid_bytes(len(t))
print("grid bytes", nb / 1e6, "MB")
def both(name, fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): out = fn()
    e1.record(); t_host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    print(f"{name:42s} host enqueue {t_host:7.3f} ms   device {e0.elapsed_time(e1) / n:7.3f} ms")
both("torch.empty(grid bytes)", lambda: torch.empty((nb,), dtype=torch.uint8, device="cuda"))
buf = torch.empty((nb,), dtype=torch.uint8, device="cuda")
both("build into one buffer", lambda: ops.check(lib.mvlm_snap_grid_build(ops.ptr(dv), ops.ptr(dt), len(t), ops.ptr(buf), nb, ops.cur_stream())))
both("SnapGrid()", lambda: ops.SnapGrid(dv, dt))
lm = torch.from_numpy(v[::13699][:73].astype(np.float64) + 0.5).cuda()
both("SnapGrid().query(near landmarks)", lambda: ops.SnapGrid(dv, dt).query(lm))
far = torch.from_numpy(np.random.RandomState(0).uniform(-3000, 3000, (73, 3))).cuda()
g = ops.SnapGrid(dv, dt)
both("query(landmarks up to 3 m away)", lambda: g.query(far))
both("scan (same)", lambda: ops.snap_to_mesh(dv, dt, far))
