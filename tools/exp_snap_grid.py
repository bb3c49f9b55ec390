"""GPU box: brute-force surface snap vs the uniform-grid path (build + query) over mesh sizes; finds the crossover
that ops.SNAP_GRID_MIN_TRIS encodes.  Run: python tools/exp_snap_grid.py > gpurun_out/snap_grid.txt"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mvlm_b200 import ops, synth  # noqa: E402


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    L = 73
    rng = np.random.RandomState(0)
    print(f"{'tris':>9} {'brute ms':>9} {'build ms':>9} {'query ms':>9} {'grid total':>10} {'speed-up(1 use)':>15} "
          f"{'tests/lm':>9} {'dims':>16} {'oversize':>8} identical | landmarks 10 mm / anywhere in the box: query ms (handed to the scan)")
    for g in (100, 160, 224, 320, 400, 560, 720, 1000):
        verts, _, tris = synth.face_mesh(grid=g, seed=1)
        lm = verts[rng.randint(0, len(verts), L)].astype(np.float64) + rng.normal(0, 1.0, (L, 3))  # ~1 mm off the surface
        dv, dt, dl = torch.from_numpy(verts).cuda(), torch.from_numpy(tris).cuda(), torch.from_numpy(lm).cuda()
        ws = torch.empty((ops._lib.load().mvlm_snap_workspace_bytes(L, len(tris)),), dtype=torch.uint8, device="cuda")
        t_brute = timed(lambda: ops.snap_to_mesh(dv, dt, dl, workspace=ws))
        grid = ops.SnapGrid(dv, dt)
        lib = ops._lib.load()
        t_build = timed(lambda: ops.check(lib.mvlm_snap_grid_build(ops.ptr(dv), ops.ptr(dt), len(tris), ops.ptr(grid.buf),
                                                                    grid.buf.numel(), ops.cur_stream()), "build"))
        t_query = timed(lambda: grid.query(dl))
        a, ta = ops.snap_to_mesh(dv, dt, dl, workspace=ws)
        b, tb, st = grid.query(dl, want_stats=True)
        info = grid.describe()
        same = bool(torch.equal(a, b) and torch.equal(ta, tb))
        far = []
        for kind in ("10mm", "box"):
            if kind == "10mm":
                lf = verts[rng.randint(0, len(verts), L)].astype(np.float64) + rng.normal(0, 1.0, (L, 3)) / np.sqrt(3) * 10.0
            else:
                lf = rng.uniform(verts.min(0), verts.max(0), (L, 3))
            dlf = torch.from_numpy(lf).cuda()
            t_far = timed(lambda: grid.query(dlf))
            a2, ta2 = ops.snap_to_mesh(dv, dt, dlf, workspace=ws)
            b2, tb2, st2 = grid.query(dlf, want_stats=True)
            same = same and bool(torch.equal(a2, b2) and torch.equal(ta2, tb2))
            far.append(f"{t_far:.4f} ({int((st2[:, 1] < 0).sum())})")
        print(f"{len(tris):>9} {t_brute:>9.4f} {t_build:>9.4f} {t_query:>9.4f} {t_build + t_query:>10.4f} "
              f"{t_brute / (t_build + t_query):>15.2f} {int(st[:, 0].float().median()):>9} {str(info['dims']):>16} "
              f"{info['n_oversize']:>8} {same} | {far[0]} / {far[1]}")


if __name__ == "__main__":
    main()
