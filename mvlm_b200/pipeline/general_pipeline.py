"""Drop-in boundary: the reference's Pipeline (src/mvlm/pipeline/general_pipeline.py:39-146).

`predict_one_file(path)` keeps the reference's signature, error behaviour and stage prints.  When
the three components are the B200 ones (the default), the scan stays on the GPU from mesh upload
to the final (L,3) landmarks ("fused" path); if a user swapped in their own Predictor2D /
renderer / estimator, the reference's seam-by-seam numpy flow is used instead.
"""
from __future__ import annotations

__all__ = ["Pipeline"]

import abc
import dataclasses
import time
from pathlib import Path

import numpy as np
import torch

from ..io_obj import Mesh, load_obj
from ..prediction.paulsenpredictor import PaulsenModel
from ..utils import Estimator3D, ObjRenderer3D, prealign


class _DeviceLock:
    """Re-entrant lock that also makes the pipeline's device the current CUDA device while it is held: every kernel
    of the library is launched on the current stream of the CURRENT device, so create_pipeline(device="cuda:1") must
    not depend on what the calling thread's current device happens to be."""

    def __init__(self, device: torch.device):
        import threading

        self._lock = threading.RLock()
        self._device = device
        self._guards = []

    def __enter__(self):
        self._lock.acquire()
        guard = torch.cuda.device(self._device) if self._device.type == "cuda" else None
        if guard is not None:
            guard.__enter__()
        self._guards.append(guard)
        return self

    def __exit__(self, *exc):
        guard = self._guards.pop()
        try:
            if guard is not None:
                guard.__exit__(*exc)
        finally:
            self._lock.release()
        return False


class TimeMixin:
    def __init__(self):
        self.start_time = time.time()
        self.end_time = None

    def tic(self):
        self.start_time = time.time()

    def toc(self):
        self.end_time = time.time()
        return self.end_time - self.start_time

    def toc_p(self):
        return self.p_time(self.toc())

    def p_time(self, t):
        return f"{t:08.6f} s"


class Pipeline(abc.ABC, TimeMixin):
    def __init__(self, render_image_stack: bool = False, offscreen: bool = True, n_views: int = 8,
                 render_image_folder: Path | None = None, visualize_rays: bool = False,
                 screenshot_folder: Path | None = None, *, image_size: tuple = (256, 256),
                 channel_mode: str = "RGB+depth", n_hypotheses: int = 1, seed: int | None = None,
                 transforms: np.ndarray | None = None, device: str = "cuda", verbose: bool = True,
                 texture_decoder: str = "pil", pre_align: dict | None = None):
        TimeMixin.__init__(self)
        self.render_image_stack = render_image_stack
        self.render_image_folder = render_image_folder
        self.n_views = n_views
        self.visualize_rays = visualize_rays  # accepted for compatibility; the VTK ray viewer is out of scope
        self.screenshot_folder = screenshot_folder
        self.verbose = verbose
        self.device = torch.device(device)
        # "pil": host decode (shared with the parity oracle); "nvjpeg": decode on the GPU into device memory
        self.texture_decoder = texture_decoder
        # legacy "pre-align" block (configs/*.json:59-84): {"align_center_of_mass", "rot_x", "rot_y", "rot_z", "scale"};
        # None / identity = off, see utils/prealign.py
        self.pre_align = None if prealign.is_identity(pre_align) else dict(pre_align)

        self.renderer_3d = ObjRenderer3D(image_size=image_size, offscreen=offscreen, n_views=n_views,
                                         channel_mode=channel_mode, device=device)
        self.renderer_3d.transforms = transforms
        self.renderer_3d.verbose = verbose
        self.renderer_3d.pre_align = self.pre_align
        self.estimator_3d = Estimator3D(n_hypotheses=n_hypotheses, seed=seed, device=device)
        self.estimator_3d.verbose = verbose
        # One pipeline object = one set of device buffers (renderer images, CNN workspace, CUDA graphs): calls from
        # several threads (the reference's FastAPI server shares one pipeline across its thread pool without a lock,
        # 3DMD_server.py:24-31) are serialised here.
        self._lock = _DeviceLock(self.device)
        self.predictor_2d = None  # assigned by the subclasses
        self.last_error = None    # "Landmarks [Error]" of the last scan (general_pipeline.py:109)

    def _print(self, *a):
        if self.verbose:
            print(*a)

    def get_lm_count(self) -> int:
        if self.predictor_2d is None:
            raise ValueError("Predictor2D is not initialized.")
        return self.predictor_2d.get_lm_count()

    # ------------------------------------------------------------------ general_pipeline.py:67-131
    def predict_one_file(self, file_name: Path, landmark_indices: list[int] | None = None,
                         view_indices: list[int] | None = None, clip_rays_to_mesh: bool = True):
        with self._lock:
            if self.predictor_2d is None:
                raise ValueError("Predictor2D is not initialized.")
            file_name = Path(file_name)
            full_s = time.time()
            if not file_name.exists():
                print(f"File {file_name} does not exist")
                return None
            fused = (isinstance(self.predictor_2d, PaulsenModel) and type(self.renderer_3d) is ObjRenderer3D
                     and type(self.estimator_3d) is Estimator3D and not self.render_image_stack
                     and self.predictor_2d.selection_method in ("simple", "moment"))
            if fused:
                r = self.renderer_3d
                if not file_name.is_file():
                    raise FileNotFoundError(f"File {file_name} is not a file")
                if not file_name.suffix == ".obj":
                    raise ValueError(f"File {file_name} is not an .obj file. Only .obj files are supported.")
                self.tic()
                mesh = load_obj(file_name, texture_decoder=self.texture_decoder, device=self.device)
                self._print("Render [1] - Setup time: ", self.toc_p())
                landmarks = self.predict_mesh(mesh)
                self._print("Landmarks 3D Total: ", self.p_time(time.time() - full_s))
                return landmarks
            return self._predict_seams(file_name, full_s)

    def predict_files(self, file_names, prefetch: int = 2) -> list:
        """Batch driver (the reference's main.py:50-62 loop is strictly serial): the native loader parses / decodes the
        next `prefetch` scans on background threads (the C parser and the JPEG decoder release the GIL) and the copies
        and launches of scan i+1 are enqueued while the GPU still works on scan i (see predict_meshes).
        Returns one (L,3) array (or None for a missing file) per input, in order."""
        with self._lock:
            from concurrent.futures import ThreadPoolExecutor

            if self.predictor_2d is None:
                raise ValueError("Predictor2D is not initialized.")
            files = [Path(f) for f in file_names]

            def load(f: Path):
                if not f.exists():
                    return None
                if not f.is_file():
                    raise FileNotFoundError(f"File {f} is not a file")
                if not f.suffix == ".obj":
                    raise ValueError(f"File {f} is not an .obj file. Only .obj files are supported.")
                # four parser threads per scan: with more, the loaders of `prefetch` scans oversubscribe the host and the
                # thread that feeds the GPU gets descheduled (measured on the 16-core box: 36 scans/s with 16, 51 with 4)
                return load_obj(f, n_threads=4, texture_decoder=self.texture_decoder, device=self.device)

            def meshes():
                with ThreadPoolExecutor(max_workers=max(1, prefetch)) as pool:
                    pending = [pool.submit(load, f) for f in files[:prefetch]]
                    for i, f in enumerate(files):
                        mesh = pending.pop(0).result()
                        if i + prefetch < len(files):
                            pending.append(pool.submit(load, files[i + prefetch]))
                        if mesh is None:
                            print(f"File {f} does not exist")
                        yield mesh

            return self.predict_meshes(meshes())

    def predict_folder(self, path, out=None) -> dict:
        """The reference's batch loop (main.py:23-62) for this pipeline: every `*.obj` of a folder (sorted), or one
        file; landmarks are written as `<out>/<stem>_<pipeline name>.txt` with `np.savetxt(..., delimiter=",")`, files
        whose landmarks could not be predicted are skipped.  Returns {file: landmarks | None}."""
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(f"{path.as_posix()} does not exist.")
        if path.is_file():
            if path.suffix.lower() != ".obj":
                raise ValueError(f"{path.as_posix()} is not an .obj file.")
            files, out_dir = [path], Path(out) if out is not None else path.parent
        else:
            files, out_dir = sorted(path.glob("*.obj")), Path(out) if out is not None else path
            if len(files) == 0:
                raise ValueError("Given folder does not contain any .obj files.")
        out_dir.mkdir(parents=True, exist_ok=True)
        pname = getattr(self, "name", type(self).__name__.lower())
        results = dict(zip(files, self.predict_files(files)))
        for f, lm in results.items():
            if lm is None:
                print(f"Landmarks for {f} could not be predicted -> skipping file [{f.stem}] for pipeline {pname}")
                continue
            np.savetxt((out_dir / f"{f.stem}_{pname}.txt").as_posix(), lm, delimiter=",")
        return results

    def _enqueue_mesh(self, mesh: Mesh, transforms: np.ndarray | None = None):
        """Enqueues the whole hot path of one scan on the current stream WITHOUT waiting for it: host -> device copies
        of the scan, raster, CNN, rays, consensus, snap, device -> pinned-host copy of the (L*3 + 1) result.
        Returns (pinned host tensor, CUDA event recorded after the copy, objects to keep alive until then)."""
        r, p, e = self.renderer_3d, self.predictor_2d, self.estimator_3d
        if transforms is None:
            transforms = r.generate_3d_transformations()
        transforms = np.asarray(transforms)
        back = None
        if self.pre_align is not None:
            # the whole path runs on the pre-aligned scan; the snapped landmarks are mapped back in _finish
            a, b = prealign.affine(mesh.verts, self.pre_align)
            mesh = dataclasses.replace(mesh, verts=prealign.apply(mesh.verts, a, b))
            back = (a, b)
        dmesh = r.upload(mesh)
        out = r.render_device(dmesh, transforms)
        peaks = p.predict_landmarks_device(out["u8"])
        starts, ends = e.estimate_landmark_lines_device(peaks, transforms, r.image_size[0], rot=r.rotations_device(transforms))
        if e.seed is not None:
            draws_d = e.seeded_draws_device(peaks.shape[0])
        else:
            # reference RNG replay needs the per-landmark line counts -> one small D2H of the peak values (synchronises)
            draws = e.reference_draws(peaks.cpu().numpy())
            draws_d = torch.from_numpy(draws.view(np.int32)).to(self.device)
        lm, err, _ = e.estimate_landmarks_from_lines_device(peaks, starts, ends, draws_d)
        from .. import ops

        snapped, _ = ops.snap_to_mesh(dmesh.verts, dmesh.tris, lm, grid=dmesh.snap_grid())
        result = torch.cat([snapped.reshape(-1), err.sum().reshape(1) / err.numel()])
        # ring of page-locked result buffers, allocated together on first use (cudaHostAlloc synchronises the device)
        ring = self.__dict__.setdefault("_result_ring", [])
        if not ring or ring[0].numel() != result.numel():
            ring.clear()
            ring.extend(torch.empty(result.shape, dtype=result.dtype, pin_memory=True) for _ in range(8))
        self._result_i = (getattr(self, "_result_i", -1) + 1) % len(ring)
        host = ring[self._result_i]
        host.copy_(result, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        return host, done, (dmesh, result, peaks, starts, ends, lm, back)

    def _finish(self, host: torch.Tensor, done, keep=None) -> np.ndarray:
        done.synchronize()
        result = host.numpy().copy()
        self.last_error = float(result[-1])
        self._print("Landmarks [Error]: ", f"{self.last_error:08.6f}", " mm")
        lm = result[:-1].reshape(-1, 3)
        back = keep[-1] if keep else None
        return lm if back is None else prealign.invert(lm, *back)

    def predict_mesh(self, mesh: Mesh, transforms: np.ndarray | None = None) -> np.ndarray:
        """Fused device path for an already loaded scan (host arrays in, (L,3) float64 out)."""
        with self._lock:
            host, done, keep = self._enqueue_mesh(mesh, transforms)
            return self._finish(host, done, keep)

    def predict_meshes(self, meshes, depth: int = 2) -> list:
        """Batch form of predict_mesh: same results, but up to `depth` scans are in flight -- the copies and launches of
        the next scan are enqueued (one stream, so buffers are reused in order) before the host waits for the landmarks
        of the previous one, which keeps host-side latency off the GPU's critical path.  A None entry yields None."""
        with self._lock:
            assert depth < 8, "depth is bounded by the ring of result buffers"
            results, inflight = [], []
            for mesh in meshes:
                if mesh is None:
                    inflight.append(None)
                else:
                    inflight.append(self._enqueue_mesh(mesh))
                while len([x for x in inflight if x is not None]) > depth or (inflight and inflight[0] is None):
                    head = inflight.pop(0)
                    results.append(None if head is None else self._finish(*head))
            for head in inflight:
                results.append(None if head is None else self._finish(*head))
            return results

    def _predict_seams(self, file_name: Path, full_s: float):
        self.tic()
        image_stack, transform_stack, pd = self.renderer_3d.multiview_render(file_name)
        self._print("Render [Total]: ", self.toc_p())
        if self.render_image_stack:
            self.visualize_image_stack(image_stack, file_name)
        self.tic()
        landmark_stack, valid = self.predictor_2d.predict_landmarks_from_images(image_stack)
        self._print("Prediction [Total]: ", self.toc_p())
        landmark_stack = landmark_stack[:, valid, :]
        transform_stack = transform_stack[valid]
        image_stack = image_stack[valid]
        self.tic()
        lines_s, lines_e = self.estimator_3d.estimate_landmark_lines(image_stack, landmark_stack, transform_stack)
        self._print("Landmarks [0] - From Heatmaps: ", self.toc_p())
        self.tic()
        landmarks, error = self.estimator_3d.estimate_landmarks_from_lines(landmark_stack, lines_s, lines_e)
        self._print("Landmarks [1] - From View Lines: ", self.toc_p())
        self.tic()
        landmarks = self.estimator_3d.project_landmarks_to_surface(pd, landmarks)
        if self.pre_align is not None:  # pd is the pre-aligned scan: back to the scan's own space (estimator3d.py:286)
            landmarks = prealign.invert(landmarks, *self.renderer_3d.last_pre_align)
        self._print("Landmarks [2] - Project to Surface: ", self.toc_p())
        self.last_error = error
        self._print("Landmarks [Error]: ", f"{error:08.6f}", " mm")
        self._print("Landmarks 3D Total: ", self.p_time(time.time() - full_s))
        return landmarks

    # general_pipeline.py:133-146
    def visualize_image_stack(self, image_stack: np.ndarray, file_name: Path):
        from PIL import Image

        save_folder = self.render_image_folder or file_name.parent
        if not Path(save_folder).exists():
            raise ValueError(f"Folder for --visualize-method flag [{save_folder}] does not exist.")
        for i in range(self.n_views):
            single_image = np.uint8(image_stack[i, :, :, 0:3] * 255)
            Image.fromarray(single_image).save(Path(save_folder) / f"{file_name.stem}_{i:02d}.png")
