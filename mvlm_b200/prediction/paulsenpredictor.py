"""Host-side mirror of the reference's PaulsenModel / DTU3DPredictor / BU3DFEPredictor
(src/mvlm/prediction/paulsenpredictor.py:42-243) on top of the CUDA stacked-hourglass plan
(csrc/hourglass.cu) and the peak kernels (csrc/peaks.cu, fused arg-max in csrc/conv_umma.cu).

Weights: the reference downloads `models_urls[...]` with torch.hub (:94-102).  Here, in order:
`weights=` (a state_dict or a .pth path), `<model_dir>/<name>.pth`, `$MVLM_B200_WEIGHTS_DIR`, and
otherwise -- only if `$MVLM_B200_RANDOM_INIT` is set or weights="random" -- a seeded random init.
"""
from __future__ import annotations

import abc
import os
from pathlib import Path

import numpy as np
import torch

from .. import ops
from ..weights import IMAGE_CHANNELS, seeded_state_dict
from .predictor2d import Predictor2D

__all__ = ["BU3DFEPredictor", "DTU3DPredictor", "PaulsenModel"]


class PaulsenModel(Predictor2D):
    def __init__(self, model_type: str, image_mode: str, n_gpus=1, batch_size=2, selection_method="simple",
                 weights=None, seed: int = 1234, device: str = "cuda"):
        super().__init__()
        self.batch_size = batch_size          # kept for API compatibility; all views run as one batch
        self.selection_method = selection_method
        self.model_type = model_type
        self.image_mode = image_mode
        self.n_gpus = n_gpus
        self.device = torch.device(device)
        self.verbose = True
        self._state_dict = self._load_weights(weights, seed)
        self._nets: dict = {}

    @abc.abstractmethod
    def get_lm_count(self) -> int:
        pass

    # ------------------------------------------------------------------ weights
    def _load_weights(self, weights, seed):
        name = f"{self.model_type}-{self.image_mode}"
        if isinstance(weights, dict):
            return weights
        if isinstance(weights, (str, Path)) and str(weights) != "random":
            ckpt = torch.load(str(weights), map_location="cpu")
            return ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        if weights is None:
            for d in (Path(__file__).parent / "models", Path(os.environ.get("MVLM_B200_WEIGHTS_DIR", "/nonexistent"))):
                for cand in sorted(d.glob("*.pth")):
                    if self._checkpoint_matches(cand.name):
                        ckpt = torch.load(str(cand), map_location="cpu")
                        return ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        if weights == "random" or os.environ.get("MVLM_B200_RANDOM_INIT"):
            return seeded_state_dict(self.get_lm_count(), self.image_mode, seed)
        raise RuntimeError(
            f"no weights for {name}: pass weights=<state_dict|path>, put a checkpoint in prediction/models/ or "
            "$MVLM_B200_WEIGHTS_DIR, or set MVLM_B200_RANDOM_INIT=1 for a seeded random init (no network access)")

    def _checkpoint_matches(self, file_name: str) -> bool:
        """The reference's checkpoint files (paulsenpredictor.py:15-39) are named <model_type>_<mode>_<date...>.pth with
        the mode in either case ("..._DTU3D_Depth_19092019..."); "RGB" must not match "RGB+depth_..." nor "geometry"
        "geometry+depth_...": the mode has to be followed by a separator.  (A wrong file would still be rejected by the
        tensor-size check of mvlm_hourglass_create.)"""
        low = file_name.lower()
        for sep in ("_", "-"):
            prefix = f"{self.model_type}{sep}{self.image_mode}".lower()
            if low.startswith(prefix) and len(low) > len(prefix) and low[len(prefix)] in "_-.":
                return True
        return False

    # ------------------------------------------------------------------ network cache
    def network(self, n_views: int, h: int, w: int) -> ops.Hourglass:
        key = (n_views, h, w)
        if key not in self._nets:
            # at most two plans alive (4.4 GB of workspace each at 100 views of 256^2): the current stack shape and, when
            # a stack is sliced along the view axis, the shorter last slice
            while len(self._nets) >= 2:
                self._nets.pop(next(iter(self._nets)))
            if self._nets and self.device.type == "cuda":
                free, _ = torch.cuda.mem_get_info(self.device)
                if free < self.workspace_bytes(n_views, h, w) * 1.2:
                    self._nets.clear()
            self._nets[key] = ops.Hourglass(self._state_dict, self.get_lm_count(), IMAGE_CHANNELS[self.image_mode],
                                            n_views, h, w, device=self.device)
        return self._nets[key]

    def workspace_bytes(self, n_views: int, h: int, w: int) -> int:
        from .. import _lib

        return int(_lib.load().mvlm_hourglass_workspace_bytes(self.get_lm_count(), IMAGE_CHANNELS[self.image_mode], n_views, h, w))

    def max_views_per_launch(self, v: int, h: int, w: int) -> int:
        """Largest slice of the view axis one plan may take: 32-bit pixel indices and, on a GPU, half of the free
        memory for the (packed) workspace.  Prefers a divisor of v so that every slice reuses the same plan.
        200 views of 512^2 (config C4) are one plan of 35 GB."""
        limit = (2 ** 32 - 1) // (h * w)
        if self.device.type == "cuda":
            free, _ = torch.cuda.mem_get_info(self.device)
            cached = sum(n.workspace.numel() for n in self._nets.values())
            per_view = max(1, self.workspace_bytes(min(v, 64), h, w) // min(v, 64))
            limit = min(limit, max(1, int((free + cached) * 0.5) // per_view))
        if v <= limit:
            return v
        best = max(d for d in range(1, limit + 1) if v % d == 0)
        return best if best * 2 > limit else limit

    # ------------------------------------------------------------------ paulsenpredictor.py:167-217
    def predict_landmarks_from_images(self, image_stack: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        n_views = image_stack.shape[0]
        valid = np.ones((n_views), dtype=bool)
        img = torch.from_numpy(np.ascontiguousarray(image_stack, dtype=np.float32)).to(self.device)
        peaks = self.predict_landmarks_device(img)
        return peaks.cpu().numpy(), valid

    def predict_landmarks_device(self, img: torch.Tensor) -> torch.Tensor:
        """img: (V,H,W,4) uint8 from the rasteriser or (V,H,W,C) float32 in [0,1] on the device.
        Returns peaks (L,V,3) float32 on the device."""
        cin = IMAGE_CHANNELS[self.image_mode]
        if img.dtype == torch.float32 and img.shape[3] > cin:
            img = img[..., :cin].contiguous()   # e.g. an RGB model fed the 4-channel stack
        v, h, w = img.shape[0], img.shape[1], img.shape[2]
        # View batches: the conv kernel indexes pixels with 32 bits (V*H*W < 2^32) and the plan's workspace grows with V
        # (about 0.7 KB per pixel); larger stacks run as equal slices of the view axis.
        chunk = v if (v, h, w) in self._nets else self.max_views_per_launch(v, h, w)  # a cached plan fits by construction
        if chunk < v:
            return torch.cat([self.predict_landmarks_device(img[i:i + chunk]).clone() for i in range(0, v, chunk)], dim=1)
        net = self.network(v, h, w)
        if self.selection_method in ("simple", "moment"):
            # the rasteriser's persistent u8 image has a stable address: replay the captured CUDA graph.  "moment" is
            # fused too: its 31x31 windows are re-evaluated around the fused arg-max (csrc/peaks.cu), no heat maps
            peaks, _ = net.forward(img, want_heatmaps=False, want_peaks=True, graph=img.dtype == torch.uint8,
                                   selection_method=self.selection_method)
            return peaks
        # the reference leaves the coordinates at zero for unknown methods (:118,:129)
        return torch.zeros((self.get_lm_count(), v, 3), dtype=torch.float32, device=self.device)

    def can_write_keys(self, v: int, h: int, w: int) -> bool:
        """True when `predict_keys_device` applies: the default arg-max selection and one plan for the whole block."""
        return self.selection_method == "simple" and ((v, h, w) in self._nets or self.max_views_per_launch(v, h, w) >= v)

    def predict_keys_device(self, img: torch.Tensor, out_keys: torch.Tensor) -> None:
        """View-split hand-off: the network's fused arg-max writes the (V, L) u64 keys of this block of views straight
        into `out_keys` (this rank's slot of the all-gather buffer, sharding.predict_mesh_view_split)."""
        cin = IMAGE_CHANNELS[self.image_mode]
        if img.dtype == torch.float32 and img.shape[3] > cin:
            img = img[..., :cin].contiguous()
        v, h, w = img.shape[0], img.shape[1], img.shape[2]
        self.network(v, h, w).forward_keys(img, out_keys)

    def find_maxima_in_batch_of_heatmaps(self, heatmaps, heatmap_maxima=None):
        """(V,L,H,W) float32 (numpy or torch) -> (L,V,3); paulsenpredictor.py:160-165."""
        hm = heatmaps if isinstance(heatmaps, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(heatmaps))
        if hm.dim() != 4:
            raise RuntimeError(f"Unexpected heatmap tensor shape: {tuple(hm.shape)}")
        out = ops.heatmap_peaks(hm.to(self.device, torch.float32), self.selection_method).cpu().numpy()
        if heatmap_maxima is not None:
            heatmap_maxima[...] = out
            return heatmap_maxima
        return out


class BU3DFEPredictor(PaulsenModel):
    def __init__(self, batch_size=2, selection_method="simple", n_gpus=1, image_mode="RGB+depth", **kw):
        super().__init__(model_type="MVLMModel_BU_3DFE", image_mode=image_mode, n_gpus=n_gpus, batch_size=batch_size,
                         selection_method=selection_method, **kw)

    def get_lm_count(self) -> int:
        return 84


class DTU3DPredictor(PaulsenModel):
    def __init__(self, batch_size=2, selection_method="simple", n_gpus=1, image_mode="RGB+depth", **kw):
        super().__init__(model_type="MVLMModel_DTU3D", image_mode=image_mode, n_gpus=n_gpus, batch_size=batch_size,
                         selection_method=selection_method, **kw)

    def get_lm_count(self) -> int:
        return 73
