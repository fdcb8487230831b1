"""Import alias: ``import mst_b200`` loads the package that lives in the directory
``diffusion-based-motion-style-transfer_b200/`` (its name mirrors the reference repository and is not
an importable identifier)."""
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_pkg_dir = os.path.join(_root, "diffusion-based-motion-style-transfer_b200")
_spec = importlib.util.spec_from_file_location(
    "mst_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mst_b200"] = _mod
_spec.loader.exec_module(_mod)
