"""Import shim: loads the package directory `3dgan_b200/` under the importable name `b200gan`."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "3dgan_b200")
_spec = importlib.util.spec_from_file_location("b200gan", os.path.join(_pkg, "__init__.py"),
                                               submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200gan"] = _mod
_spec.loader.exec_module(_mod)
