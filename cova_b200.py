"""Import shim: the package directory is named ``cova-web-object-detection_b200`` (not a legal
Python identifier), so ``import cova_b200`` loads it from that directory as the package ``cova_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cova-web-object-detection_b200")
_spec = importlib.util.spec_from_file_location(
    "cova_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cova_b200"] = _mod
_spec.loader.exec_module(_mod)
