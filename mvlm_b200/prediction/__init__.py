__all__ = ["BU3DFEPredictor", "DTU3DPredictor", "PaulsenModel", "Predictor2D"]

from .paulsenpredictor import BU3DFEPredictor, DTU3DPredictor, PaulsenModel
from .predictor2d import Predictor2D
