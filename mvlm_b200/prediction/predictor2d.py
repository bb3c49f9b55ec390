"""The reference's plugin seam (src/mvlm/prediction/predictor2d.py:8-27): any 2D landmark
predictor can be assigned to `Pipeline.predictor_2d`."""
__all__ = ["Predictor2D"]

import abc

import numpy as np


class Predictor2D(abc.ABC):
    def __init__(self):
        pass

    @abc.abstractmethod
    def predict_landmarks_from_images(self, image_stack: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        """Returns (landmarks (n_landmarks, n_views, 3) [row, col, value], valid (n_views,) bool)."""

    @abc.abstractmethod
    def get_lm_count(self) -> int:
        pass
