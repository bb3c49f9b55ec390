"""ctypes wrapper of oracle/csrc/oracle_native.c (CPU oracle: renderer + surface snap)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle_native.so"
_lib = None

CHANNEL_MODES = {"RGB+depth": 0, "geometry+depth": 1, "RGB": 2, "depth": 3, "geometry": 4}
MODE_CHANNELS = {0: 4, 1: 2, 2: 3, 3: 1, 4: 1}


def build() -> Path:
    src = _HERE / "csrc" / "oracle_native.c"
    if not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.oracle_raster_multiview.restype = C.c_int
        _lib.oracle_snap_to_mesh.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def raster_multiview(verts, uvs, tris, tex, rot, h, w, channel_mode="RGB+depth"):
    """verts (Nv,3) f32, uvs (Nv,2) f32|None, tris (Nt,3) i32, tex (Th,Tw,3) u8|None, rot (V,3,3) f64.
    Returns (image (V,H,W,C) f32, tri_id (V,H,W) i32, zbuf (V,H,W) f32)."""
    lib = load()
    verts = np.ascontiguousarray(verts, np.float32)
    tris = np.ascontiguousarray(tris, np.int32)
    rot = np.ascontiguousarray(rot, np.float64).reshape(-1, 9)
    uvs = None if uvs is None else np.ascontiguousarray(uvs, np.float32)
    tex = None if tex is None else np.ascontiguousarray(tex, np.uint8)
    mode = CHANNEL_MODES[channel_mode]
    v = rot.shape[0]
    img = np.empty((v, h, w, MODE_CHANNELS[mode]), np.float32)
    tri = np.empty((v, h, w), np.int32)
    z = np.empty((v, h, w), np.float32)
    th, tw = (tex.shape[0], tex.shape[1]) if tex is not None else (0, 0)
    rc = lib.oracle_raster_multiview(_p(verts), _p(uvs), C.c_int(len(verts)), _p(tris), C.c_int(len(tris)),
                                     _p(tex), C.c_int(th), C.c_int(tw), _p(rot), C.c_int(v), C.c_int(h),
                                     C.c_int(w), C.c_int(mode), _p(img), _p(tri), _p(z))
    if rc != 0:
        raise RuntimeError(f"oracle_raster_multiview failed: {rc}")
    return img, tri, z


def snap_to_mesh(verts, tris, landmarks):
    """Exact closest point on the mesh per landmark, fp64.  Returns ((L,3) f64, tri ids (L,) i32)."""
    lib = load()
    verts = np.ascontiguousarray(verts, np.float32)
    tris = np.ascontiguousarray(tris, np.int32)
    lm = np.ascontiguousarray(landmarks, np.float64)
    out = np.empty_like(lm)
    tid = np.empty((lm.shape[0],), np.int32)
    rc = lib.oracle_snap_to_mesh(_p(verts), _p(tris), C.c_int(len(tris)), _p(lm), C.c_int(lm.shape[0]), _p(out), _p(tid))
    if rc != 0:
        raise RuntimeError(f"oracle_snap_to_mesh failed: {rc}")
    return out, tid
