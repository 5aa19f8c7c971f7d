"""Import shim: the product package lives in the directory ``3dgs-native_b200/`` (not a valid
Python identifier), so ``import gsb200`` loads that directory as the package ``gsb200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3dgs-native_b200")
_spec = importlib.util.spec_from_file_location("gsb200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gsb200"] = _mod
_spec.loader.exec_module(_mod)
