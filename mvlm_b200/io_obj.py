"""Scan loader: Wavefront OBJ (+ optional `<stem>.jpg` texture) -> flat arrays.

Replaces vtkOBJReader / vtkJPEGReader as used by the reference's obj_to_actor
(src/mvlm/utils/utils3d.py:10-36): positions `v`, texture coordinates `vt`, faces
`f a[/b[/c]] ...` (polygons are fan-triangulated).  Like vtkOBJReader, a position that is
referenced with different `vt` indices is duplicated so that every output vertex has exactly
one (position, uv) pair.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np


@dataclass
class Mesh:
    verts: np.ndarray          # (Nv,3) float32
    tris: np.ndarray           # (Nt,3) int32
    uvs: np.ndarray | None     # (Nv,2) float32
    texture: np.ndarray | None  # (Th,Tw,3) uint8, row 0 = top of the image
    path: Path | None = None

    @property
    def bbox_diagonal(self) -> float:
        return float(np.linalg.norm(self.verts.max(0) - self.verts.min(0)))


def load_obj(path: Path | str, load_texture: bool = True) -> Mesh:
    path = Path(path)
    if not path.is_file():
        raise ValueError(f"File {path} does not exist.")
    v_rows, vt_rows, f_rows = [], [], []
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            if line.startswith("v "):
                v_rows.append(line[2:])
            elif line.startswith("vt "):
                vt_rows.append(line[3:])
            elif line.startswith("f "):
                f_rows.append(line[2:])
    if len(v_rows) == 0:
        raise ValueError(f"File {path} does not contain any points.")
    pos = np.loadtxt(v_rows, dtype=np.float64, ndmin=2, usecols=(0, 1, 2)).astype(np.float32)
    tex = np.loadtxt(vt_rows, dtype=np.float64, ndmin=2, usecols=(0, 1)).astype(np.float32) if vt_rows else None
    corners_v, corners_t = [], []
    for row in f_rows:
        toks = row.split()
        vi, ti = [], []
        for tok in toks:
            parts = tok.split("/")
            a = int(parts[0])
            vi.append(a - 1 if a > 0 else len(pos) + a)
            if len(parts) > 1 and parts[1] != "":
                b = int(parts[1])
                ti.append(b - 1 if b > 0 else (len(tex) + b if tex is not None else -1))
            else:
                ti.append(-1)
        for k in range(1, len(vi) - 1):
            corners_v.append((vi[0], vi[k], vi[k + 1]))
            corners_t.append((ti[0], ti[k], ti[k + 1]))
    cv = np.asarray(corners_v, dtype=np.int64).reshape(-1, 3)
    ct = np.asarray(corners_t, dtype=np.int64).reshape(-1, 3)
    if tex is None or (ct < 0).all():
        verts, uvs, tris = pos, None, cv.astype(np.int32)
    else:
        # unique (position, uv) pairs -> output vertices
        pairs = np.stack([cv.reshape(-1), ct.reshape(-1)], 1)
        uniq, inv = np.unique(pairs, axis=0, return_inverse=True)
        verts = pos[uniq[:, 0]]
        uvs = np.where(uniq[:, 1:2] >= 0, tex[np.clip(uniq[:, 1], 0, None)], 0.0).astype(np.float32)
        tris = inv.reshape(-1, 3).astype(np.int32)
    texture = None
    jpg = path.with_suffix(".jpg")
    if load_texture and uvs is not None and jpg.exists():
        try:
            from PIL import Image

            texture = np.asarray(Image.open(jpg).convert("RGB"), dtype=np.uint8)
        except Exception:  # noqa: BLE001 - same policy as the reference: ignore an unreadable texture
            texture = None
    return Mesh(verts=np.ascontiguousarray(verts), tris=np.ascontiguousarray(tris),
                uvs=None if uvs is None else np.ascontiguousarray(uvs), texture=texture, path=path)
