__all__ = ["ObjRenderer3D", "ObjVTKRenderer3D", "Estimator3D", "prealign"]

from . import prealign
from .estimator3d import Estimator3D
from .render3d import ObjRenderer3D, ObjVTKRenderer3D
