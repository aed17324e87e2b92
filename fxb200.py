"""Import shim: the package directory is named ``featureextraction.jl_b200`` (not a
valid Python identifier), so ``import fxb200`` loads it under this alias."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "featureextraction.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "fxb200", os.path.join(_pkg_dir, "__init__.py"),
    submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fxb200"] = _mod
_spec.loader.exec_module(_mod)
