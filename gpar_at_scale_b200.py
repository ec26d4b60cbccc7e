"""Import alias: the package directory is `gpar-at-scale_b200/` (not a valid Python identifier);
this loader registers it as the module `gpar_at_scale_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpar-at-scale_b200")
_spec = importlib.util.spec_from_file_location(
    "gpar_at_scale_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gpar_at_scale_b200"] = _mod
_spec.loader.exec_module(_mod)
