"""Pre-alignment of a scan before rendering (the "pre-align" block of the legacy configs,
/root/reference/configs/*.json:59-84, and the commented-out apply_pre_transformation /
transform_landmarks_to_original_space pair of src/mvlm/utils/estimator3d.py:186-248,257,286).

    t = vtkTransform(); t.Scale(s,s,s); t.RotateY(ry); t.RotateX(rx); t.RotateZ(rz); t.Translate(-cm)

vtkTransform concatenates in pre-multiply mode (each call right-multiplies the current matrix), so a point
goes through  p' = S Ry Rx Rz (p - cm):  centre of mass to the origin (vtkCenterOfMass without weights = the mean
of ALL points), rotate about z, x, y (degrees), scale.  The whole path (render, CNN, rays, consensus, snap) then
runs on the pre-aligned mesh and the snapped landmarks are mapped back with the inverse transform (:226-248,:286).
Off by default, like in the shipped configs (identity, align_center_of_mass false).
"""
from __future__ import annotations

import numpy as np

DEFAULT = {"align_center_of_mass": False, "rot_x": 0.0, "rot_y": 0.0, "rot_z": 0.0, "scale": 1.0}


def is_identity(cfg: dict | None) -> bool:
    if not cfg:
        return True
    c = {**DEFAULT, **cfg}
    return (not c["align_center_of_mass"]) and c["rot_x"] == 0 and c["rot_y"] == 0 and c["rot_z"] == 0 and c["scale"] == 1


def affine(verts: np.ndarray, cfg: dict) -> tuple[np.ndarray, np.ndarray]:
    """(A (3,3), b (3,)) float64 with p' = A p + b for this scan."""
    c = {**DEFAULT, **cfg}
    unknown = set(cfg) - set(DEFAULT) - {"write_pre_aligned"}
    if unknown:
        raise ValueError(f"unknown pre-align keys: {sorted(unknown)}")
    rx, ry, rz = (np.deg2rad(float(c[k])) for k in ("rot_x", "rot_y", "rot_z"))
    mx = np.array([[1, 0, 0], [0, np.cos(rx), -np.sin(rx)], [0, np.sin(rx), np.cos(rx)]])
    my = np.array([[np.cos(ry), 0, np.sin(ry)], [0, 1, 0], [-np.sin(ry), 0, np.cos(ry)]])
    mz = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]])
    a = float(c["scale"]) * (my @ mx @ mz)
    cm = np.asarray(verts, np.float64).mean(axis=0) if c["align_center_of_mass"] else np.zeros(3)
    return a, -(a @ cm)


def apply(verts: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """float32 points through the double-precision transform, stored as float32 (vtkTransformPolyDataFilter)."""
    return (np.asarray(verts, np.float64) @ a.T + b).astype(np.float32)


def invert(points: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Landmarks of the pre-aligned space back to the scan's own space (t.GetInverse(), :238)."""
    return np.linalg.solve(a, (np.asarray(points, np.float64) - b).T).T
