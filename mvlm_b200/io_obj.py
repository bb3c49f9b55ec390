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
    texture: object | None     # (Th,Tw,3) uint8, row 0 = top of the image: numpy array, or a CUDA tensor (nvJPEG decode)
    path: Path | None = None
    texture_ready: object | None = None  # CUDA event recorded after a device-side decode (texture is a CUDA tensor)

    @property
    def bbox_diagonal(self) -> float:
        return float(np.linalg.norm(self.verts.max(0) - self.verts.min(0)))


def _texture_file(path: Path):
    """`<stem>.jpg` next to the scan (utils3d.py:26); FLAME models (`flame*.obj`) use `mean_texture.jpg` of their folder
    instead (utils3d.py:39-51).  Returns None when there is no texture file."""
    if path.name.startswith("flame"):
        print("Loading flame texture")
        jpg = path.parent / "mean_texture.jpg"
        if not jpg.exists():
            print("Could not load flame texture", jpg)
            return None
        return jpg
    jpg = path.with_suffix(".jpg")
    return jpg if jpg.exists() else None


def _load_texture(path: Path):
    jpg = _texture_file(path)
    if jpg is None:
        return None
    try:
        from PIL import Image

        return np.asarray(Image.open(jpg).convert("RGB"), dtype=np.uint8)
    except Exception:  # noqa: BLE001 - same policy as the reference: ignore an unreadable texture
        return None


_tls = None


def _decode_texture_nvjpeg(path: Path, device):
    """`<stem>.jpg` -> (Th,Tw,3) uint8 CUDA tensor + the event recorded behind the decode (mvlm_jpeg_decode_rgb on a
    per-thread side stream: the loader threads of Pipeline.predict_files decode while the main stream computes)."""
    import threading

    import torch

    from . import _lib

    global _tls
    if _tls is None:
        _tls = threading.local()
    jpg = _texture_file(path)
    if jpg is None:
        return None, None
    data = jpg.read_bytes()
    lib = _lib.load()
    import ctypes as C

    w, h = C.c_int(), C.c_int()
    if lib.mvlm_jpeg_info(data, len(data), C.byref(w), C.byref(h)) != 0:
        return None, None  # same policy as the reference: ignore an unreadable texture
    device = torch.device(device)
    if getattr(_tls, "stream", None) is None or _tls.device != device:
        _tls.stream, _tls.device = torch.cuda.Stream(device), device
    with torch.cuda.device(device), torch.cuda.stream(_tls.stream):
        tex = torch.empty((h.value, w.value, 3), dtype=torch.uint8, device=device)
        if lib.mvlm_jpeg_decode_rgb(data, len(data), tex.data_ptr(), w.value, h.value, _tls.stream.cuda_stream) != 0:
            return None, None
        ready = torch.cuda.Event()
        ready.record(_tls.stream)
    return tex, ready


def load_obj(path: Path | str, load_texture: bool = True, n_threads: int = 0, texture_decoder: str = "pil",
             device="cuda") -> Mesh:
    """Native loader (csrc/obj_loader.cu through the C-ABI, multi-threaded parse): what the pipeline uses.
    texture_decoder: "pil" (host decode, libjpeg-turbo: the decoder the parity tests share with the oracle) or
    "nvjpeg" (csrc/jpeg_decode.cu: decoded on the GPU into device memory; values may differ by a few LSB)."""
    import ctypes as C

    from . import _lib

    path = Path(path)
    if not path.is_file():
        raise ValueError(f"File {path} does not exist.")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.mvlm_obj_load(str(path).encode(), n_threads, C.byref(h))
    if rc != 0:
        raise ValueError(lib.mvlm_last_error().decode("utf-8", "replace"))
    try:
        nv, nt, has_uv = C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.mvlm_obj_counts(h, C.byref(nv), C.byref(nt), C.byref(has_uv)), "mvlm_obj_counts")
        verts = np.empty((nv.value, 3), dtype=np.float32)
        tris = np.empty((nt.value, 3), dtype=np.int32)
        uvs = np.empty((nv.value, 2), dtype=np.float32) if has_uv.value else None
        _lib.check(lib.mvlm_obj_copy(h, verts.ctypes.data, None if uvs is None else uvs.ctypes.data, tris.ctypes.data),
                   "mvlm_obj_copy")
    finally:
        lib.mvlm_obj_free(h)
    texture, ready = None, None
    if load_texture and uvs is not None:
        if texture_decoder == "nvjpeg":
            texture, ready = _decode_texture_nvjpeg(path, device)
        elif texture_decoder == "pil":
            texture = _load_texture(path)
        else:
            raise ValueError(f"Unknown texture decoder: {texture_decoder}")
    return Mesh(verts=verts, tris=tris, uvs=uvs, texture=texture, path=path, texture_ready=ready)
