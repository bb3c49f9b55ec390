"""src/mvlm/pipeline/__init__.py:15-43 -- the factory.  Only the two pipelines of the hot path
("dtu3d", "bu3dfe") exist here; the MediaPipe / dlib / face-alignment wrappers of third-party
black-box models are out of scope."""
__all__ = ["BU3DFEPipeline", "DTU3DPipeline", "Pipeline", "create_pipeline"]

from .general_pipeline import Pipeline
from .paulsen_pipeline import BU3DFEPipeline, DTU3DPipeline


def create_pipeline(name: str, **kwargs):
    name = name.lower()
    if name == "bu3dfe":
        p = BU3DFEPipeline(**kwargs)
        p.name = name
        return p
    elif name == "dtu3d":
        p = DTU3DPipeline(**kwargs)
        p.name = name
        return p
    elif name in ("mediapipe", "dlib", "face_alignment"):
        raise ValueError(f"Unknown pipeline: {name} (third-party predictor pipelines are not part of mvlm_b200)")
    else:
        raise ValueError(f"Unknown pipeline: {name}")
