"""Imports the reference's own modules IN PLACE from /root/reference (build container only).

Used by tools/make_golden.py and by tests that are skipped when the reference tree is
absent (the GPU box).  Nothing is copied: modules are executed from their original
location with stub modules for the packages that are not installable here
(`vtk`, `matplotlib`), following SURVEY.md appendix C.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

REF_ROOT = Path("/root/reference")
REF_SRC = REF_ROOT / "src"


def available() -> bool:
    return (REF_SRC / "mvlm" / "utils" / "estimator3d.py").is_file()


_cache: dict = {}


def load():
    """Returns a namespace with .paulsenpredictor, .estimator3d, .utils3d, .predictor2d modules."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present")
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot", "vtk", "vtk.util",
                                             "vtk.util.numpy_support", "mvlm", "mvlm.prediction", "mvlm.utils")}
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    vtk = types.ModuleType("vtk")
    vtk.vtkActor = object
    vtk.vtkPolyData = object
    vtk_util = types.ModuleType("vtk.util")
    vtk_ns = types.ModuleType("vtk.util.numpy_support")
    vtk_ns.vtk_to_numpy = lambda *a, **k: None
    vtk.util = vtk_util
    vtk_util.numpy_support = vtk_ns
    stubs = {"matplotlib": mpl, "matplotlib.pyplot": plt, "vtk": vtk, "vtk.util": vtk_util,
             "vtk.util.numpy_support": vtk_ns}
    sys.modules.update(stubs)
    # bare packages: do NOT execute the reference's __init__ (they import mediapipe/dlib/...)
    pkgs = {}
    for name, sub in (("mvlm", ""), ("mvlm.prediction", "prediction"), ("mvlm.utils", "utils")):
        m = types.ModuleType(name)
        m.__path__ = [str(REF_SRC / "mvlm" / sub)]
        pkgs[name] = m
    sys.modules.update(pkgs)

    def imp(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, REF_SRC / "mvlm" / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    ns = types.SimpleNamespace()
    try:
        ns.predictor2d = imp("mvlm.prediction.predictor2d", "prediction/predictor2d.py")
        ns.paulsenpredictor = imp("mvlm.prediction.paulsenpredictor", "prediction/paulsenpredictor.py")
        ns.utils3d = imp("mvlm.utils.utils3d", "utils/utils3d.py")
        ns.estimator3d = imp("mvlm.utils.estimator3d", "utils/estimator3d.py")
    finally:
        # leave no trace of the fake `mvlm`/`vtk` packages in sys.modules: the product ships its own `mvlm`
        for k in list(sys.modules):
            if k == "mvlm" or k.startswith("mvlm.") or k in stubs:
                del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    _cache["ns"] = ns
    return ns


def make_predictor(kind: str, state_dict: dict, selection_method: str = "simple", batch_size: int = 2):
    """Constructs the reference's DTU3DPredictor / BU3DFEPredictor with injected weights (no network)."""
    ns = load()
    pp = ns.paulsenpredictor
    cls = {"dtu3d": pp.DTU3DPredictor, "bu3dfe": pp.BU3DFEPredictor}[kind]
    orig = pp.load_state_dict_from_url
    pp.load_state_dict_from_url = lambda url, **kw: state_dict
    try:
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            pred = cls(batch_size=batch_size, selection_method=selection_method, n_gpus=0)
    finally:
        pp.load_state_dict_from_url = orig
    return pred


def make_model(n_landmarks: int, image_channels: str, state_dict: dict):
    """The reference's MVLMModel (torch CPU, eval) with the given weights."""
    pp = load().paulsenpredictor
    m = pp.MVLMModel(n_landmarks=n_landmarks, n_features=256, dropout_rate=0.2, image_channels=image_channels)
    m.load_state_dict(state_dict, strict=True)
    m.eval()
    return m


def load_renderer_class():
    """The reference's ObjVTKRenderer3D (utils/render3d.py) with `vtk` replaced by a permissive mock: its constructor
    only configures VTK objects, and the view-list generators `random_transform` / `generate_3d_transformations`
    (:79-112) are plain numpy -- those are what tools/make_golden.py runs verbatim.  The VTK render path itself
    cannot be executed here."""
    from unittest import mock

    if not available():
        raise RuntimeError("reference tree not present")
    names = ("vtk", "vtk.util", "vtk.util.numpy_support", "mvlm", "mvlm.utils", "mvlm.utils.utils3d")
    saved = {k: sys.modules.get(k) for k in names}
    vtk = mock.MagicMock(name="vtk")
    sys.modules.update({"vtk": vtk, "vtk.util": vtk.util, "vtk.util.numpy_support": vtk.util.numpy_support})
    pk = types.ModuleType("mvlm")
    pk.__path__ = [str(REF_SRC / "mvlm")]
    pu = types.ModuleType("mvlm.utils")
    pu.__path__ = [str(REF_SRC / "mvlm" / "utils")]
    sys.modules.update({"mvlm": pk, "mvlm.utils": pu})
    try:
        for modname, rel in (("mvlm.utils.utils3d", "utils/utils3d.py"), ("mvlm.utils.render3d", "utils/render3d.py")):
            spec = importlib.util.spec_from_file_location(modname, REF_SRC / "mvlm" / rel)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[modname] = mod
            spec.loader.exec_module(mod)
        cls = sys.modules["mvlm.utils.render3d"].ObjVTKRenderer3D
    finally:
        for k in list(sys.modules):
            if k == "mvlm" or k.startswith("mvlm.") or k in ("vtk", "vtk.util", "vtk.util.numpy_support"):
                del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    return cls
