"""Seeded synthetic inputs: a textured face-like mesh (OBJ + JPG), view transforms, ray sets.

The reference's sample scans are missing blobs (.MISSING_LARGE_BLOBS) and there is no network,
so every benchmark and parity test runs on these.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def face_mesh(grid: int = 224, seed: int = 1234):
    """Height-field 'face': grid x grid vertices on an ellipsoidal cap (+-90 x +-110 mm, ~80 mm deep)
    with seeded low-frequency bumps, centred so that it stays inside the +-150 mm orthographic
    window for every view angle the renderer draws.  Returns (verts f32 (N,3), uvs f32 (N,2),
    tris i32 (T,3))."""
    rng = np.random.RandomState(seed)
    a = np.linspace(-1.0, 1.0, grid)
    aa, bb = np.meshgrid(a, a, indexing="xy")
    r2 = (aa ** 2 + bb ** 2) / 2.0
    z = 80.0 * np.sqrt(np.clip(1.0 - r2, 0.0, 1.0))
    # nose / brow / chin style bumps
    for _ in range(6):
        cx, cy = rng.uniform(-0.5, 0.5, 2)
        s = rng.uniform(0.08, 0.3)
        amp = rng.uniform(-6.0, 14.0)
        z += amp * np.exp(-((aa - cx) ** 2 + (bb - cy) ** 2) / (2 * s * s))
    z += 18.0 * np.exp(-((aa) ** 2 + (bb + 0.05) ** 2) / (2 * 0.12 ** 2))  # nose
    x = 90.0 * aa
    y = 110.0 * bb
    verts = np.stack([x, y, z], -1).reshape(-1, 3)
    verts -= verts.mean(0, keepdims=True)
    # keep inside the 150 mm view sphere
    rmax = np.linalg.norm(verts, axis=1).max()
    if rmax > 145.0:
        verts *= 145.0 / rmax
    u = (aa + 1.0) / 2.0
    v = (bb + 1.0) / 2.0
    uvs = np.stack([u, v], -1).reshape(-1, 2)
    idx = np.arange(grid * grid).reshape(grid, grid)
    q00, q01, q10, q11 = idx[:-1, :-1], idx[:-1, 1:], idx[1:, :-1], idx[1:, 1:]
    t1 = np.stack([q00, q01, q11], -1).reshape(-1, 3)
    t2 = np.stack([q00, q11, q10], -1).reshape(-1, 3)
    tris = np.concatenate([t1, t2], 0)
    return verts.astype(np.float32), uvs.astype(np.float32), tris.astype(np.int32)


def face_texture(size: int = 1024, seed: int = 1234) -> np.ndarray:
    """(size,size,3) uint8: smooth seeded colour noise plus high-contrast markers."""
    rng = np.random.RandomState(seed + 1)
    t = np.linspace(0, 1, size, dtype=np.float32)
    xx, yy = np.meshgrid(t, t, indexing="xy")
    img = np.zeros((size, size, 3), np.float32)
    for c in range(3):
        acc = np.zeros((size, size), np.float32)
        for _ in range(5):
            fx, fy = rng.uniform(1, 9, 2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += np.sin(2 * np.pi * (fx * xx + fy * yy) + ph).astype(np.float32)
        img[..., c] = 0.55 + 0.12 * acc
    for _ in range(40):
        cx, cy = rng.uniform(0.05, 0.95, 2)
        r = rng.uniform(0.004, 0.02)
        col = rng.uniform(0, 1, 3)
        m = (xx - cx) ** 2 + (yy - cy) ** 2 < r * r
        img[m] = col
    return (np.clip(img, 0, 1) * 255).astype(np.uint8)


def write_obj(path: Path, verts, uvs, tris, texture: np.ndarray | None = None, jpeg_quality: int = 95) -> Path:
    """Writes `<path>.obj` (v / vt / f a/a b/b c/c) and, if given, `<stem>.jpg` beside it
    (the file the reference's obj_to_actor looks for, src/mvlm/utils/utils3d.py:26)."""
    path = Path(path)
    with open(path, "w") as f:
        f.write("# mvlm_b200 synthetic scan\n")
        np.savetxt(f, verts, fmt="v %.6f %.6f %.6f")
        if uvs is not None:
            np.savetxt(f, uvs, fmt="vt %.6f %.6f")
            t = tris + 1
            np.savetxt(f, np.stack([t[:, 0], t[:, 0], t[:, 1], t[:, 1], t[:, 2], t[:, 2]], 1), fmt="f %d/%d %d/%d %d/%d")
        else:
            np.savetxt(f, tris + 1, fmt="f %d %d %d")
    if texture is not None:
        from PIL import Image

        Image.fromarray(texture).save(path.with_suffix(".jpg"), quality=jpeg_quality)
    return path


def random_view_transforms(n_views: int, seed: int | None = None) -> np.ndarray:
    """The reference's ObjVTKRenderer3D.random_transform draw order (src/mvlm/utils/render3d.py:79-89):
    rx in [-40,40), ry in [-80,80), rz in [-20,20) integers, then the unused scale/tx/ty draws;
    (V,6) float64.  `seed` seeds the GLOBAL numpy RNG exactly like a harness around the reference would."""
    if seed is not None:
        np.random.seed(seed)
    rx = np.random.randint(-40, 40, size=n_views)
    ry = np.random.randint(-80, 80, size=n_views)
    rz = np.random.randint(-20, 20, size=n_views)
    scale = np.random.uniform(1.4, 1.9, size=n_views)
    tx = np.random.randint(-20, 20, size=n_views)
    ty = np.random.randint(-20, 20, size=n_views)
    return np.stack((rx, ry, rz, scale, tx, ty), axis=1)


def synthetic_rays(n_landmarks: int = 84, n_views: int = 200, outlier_frac: float = 0.3, seed: int = 1234):
    """Consensus micro-benchmark input (BASELINE.json config 5): per landmark a true 3D point,
    V rays through it (0.5 mm jitter), `outlier_frac` of them displaced by U(-80,80) mm.
    Returns (peaks (L,V,3) f32 [row, col, value], starts, ends (L,V,3) f64, truth (L,3))."""
    rng = np.random.RandomState(seed)
    truth = rng.uniform(-80, 80, (n_landmarks, 3))
    ang = np.deg2rad(np.stack([rng.randint(-40, 40, n_views), rng.randint(-80, 80, n_views),
                               rng.randint(-20, 20, n_views)], 1).astype(np.float64))
    starts = np.empty((n_landmarks, n_views, 3))
    ends = np.empty((n_landmarks, n_views, 3))
    for v in range(n_views):
        ax, ay, az = ang[v]
        mx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        my = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        mz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        r = my @ mx @ mz
        cam = truth @ r.T  # camera coordinates of the points
        cam = cam + rng.normal(0, 0.5, cam.shape)
        out = rng.uniform(0, 1, n_landmarks) < outlier_frac
        cam[out, :2] += rng.uniform(-80, 80, (int(out.sum()), 2))
        s = np.stack([cam[:, 0], cam[:, 1], np.full(n_landmarks, 500.0)], 1)
        e = np.stack([cam[:, 0], cam[:, 1], np.full(n_landmarks, -500.0)], 1)
        starts[:, v] = s @ r
        ends[:, v] = e @ r
    peaks = np.zeros((n_landmarks, n_views, 3), np.float32)
    peaks[:, :, 2] = rng.uniform(0, 1, (n_landmarks, n_views)).astype(np.float32)
    return peaks, starts, ends, truth


def hypothesis_table(n_landmarks: int, n_hyp: int, seed: int = 1234) -> np.ndarray:
    """Shared seeded RANSAC draws (L,H,8) uint32; line index = draw mod n_lines_after_filter."""
    return np.random.RandomState(seed).randint(0, 2 ** 32, (n_landmarks, n_hyp, 8), dtype=np.uint64).astype(np.uint32)
