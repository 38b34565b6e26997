"""Import shim: the product package lives in ``rotate-yolov3_b200/`` (the directory name the project
layout prescribes, which is not a valid Python identifier).  This shim makes it importable as
``rotate_yolov3_b200`` by pointing the package search path at that directory and executing its
``__init__.py`` in this module's namespace."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rotate-yolov3_b200")
__path__[:] = [_real]
_init = _os.path.join(_real, "__init__.py")
with open(_init, "r") as _f:
    exec(compile(_f.read(), _init, "exec"))
del _f, _init
