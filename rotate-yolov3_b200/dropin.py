"""Import aliases that let the reference's own scripts resolve their hot-path imports to this package.

The reference imports (train.py / test.py / detect.py / model/loss.py):
    from model.models import *                    -> Darknet, YOLOLayer
    from utils.nms.nms import non_max_suppression -> non_max_suppression
    from utils.nms.r_nms import r_nms             -> the compiled extension utils/nms/r_nms*.so (pybind module ``r_nms``)
    from utils.utils import skew_bbox_iou         (utils/utils.py:290)
``install()`` covers both deployment shapes:
  * inside a checkout of the reference (its ``utils`` / ``model`` packages are importable): the compiled ``r_nms``
    module is provided (so ``make.sh`` never has to build the legacy THC extension) and the three Python-level names are
    re-bound on the reference's own modules -- every other reference function keeps working unchanged;
  * stand-alone (no reference on sys.path, e.g. the tests on the GPU box): minimal modules with exactly those names are
    registered so that reference-style import statements work."""
import importlib
import sys
import types


def _module(name):
    if name in sys.modules:
        return sys.modules[name], False
    try:
        return importlib.import_module(name), False
    except Exception:
        m = types.ModuleType(name)
        m.__dict__["__path__"] = []
        sys.modules[name] = m
        if "." in name:
            parent, _ = _module(name.rsplit(".", 1)[0])
            setattr(parent, name.rsplit(".", 1)[1], m)
        return m, True


def install():
    from . import Darknet, YOLOLayer, non_max_suppression, r_nms, skew_bbox_iou
    # 1. the compiled extension: always ours (the reference's cannot be built against torch >= 2)
    ext = types.ModuleType("utils.nms.r_nms")
    ext.r_nms = r_nms
    _module("utils")
    _module("utils.nms")
    sys.modules["utils.nms.r_nms"] = ext
    setattr(sys.modules["utils.nms"], "r_nms", ext)
    # 2. Python-level entry points
    nms_mod, _ = _module("utils.nms.nms")
    nms_mod.non_max_suppression = non_max_suppression
    nms_mod.r_nms = r_nms
    uu, _ = _module("utils.utils")
    uu.skew_bbox_iou = skew_bbox_iou
    mm, _ = _module("model.models")
    mm.Darknet = Darknet
    mm.YOLOLayer = YOLOLayer
    return {"utils.nms.r_nms": ext, "utils.nms.nms": nms_mod, "utils.utils": uu, "model.models": mm}
