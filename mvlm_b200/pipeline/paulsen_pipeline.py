"""src/mvlm/pipeline/paulsen_pipeline.py:7-15 -- pipelines that assign a Paulsen predictor."""
__all__ = ["BU3DFEPipeline", "DTU3DPipeline"]

from ..prediction import BU3DFEPredictor, DTU3DPredictor
from .general_pipeline import Pipeline


def _split(kwargs):
    pk = {k: kwargs.pop(k) for k in ("weights", "selection_method", "batch_size") if k in kwargs}
    return pk


class BU3DFEPipeline(Pipeline):
    def __init__(self, *args, **kwargs):
        pk = _split(kwargs)
        super().__init__(*args, **kwargs)
        self.predictor_2d = BU3DFEPredictor(image_mode=kwargs.get("channel_mode", "RGB+depth"),
                                            device=kwargs.get("device", "cuda"), **pk)
        self.predictor_2d.verbose = self.verbose


class DTU3DPipeline(Pipeline):
    def __init__(self, *args, **kwargs):
        pk = _split(kwargs)
        super().__init__(*args, **kwargs)
        self.predictor_2d = DTU3DPredictor(image_mode=kwargs.get("channel_mode", "RGB+depth"),
                                           device=kwargs.get("device", "cuda"), **pk)
        self.predictor_2d.verbose = self.verbose
