"""Registers the `tacotron-2_b200/` package directory (not a valid identifier) as module `tacotron2_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "tacotron-2_b200")

if "tacotron2_b200" not in sys.modules:
    _spec = importlib.util.spec_from_file_location(
        "tacotron2_b200", os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules["tacotron2_b200"] = _mod
    _spec.loader.exec_module(_mod)

t2 = sys.modules["tacotron2_b200"]
