"""Import shim: makes the package directory `pillarnet-lts_b200/` importable as `pillarnet_lts_b200`
(a hyphen is not a valid module name).  `import pillarnet_lts_b200` from the repo root loads it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pillarnet-lts_b200")
_spec = importlib.util.spec_from_file_location(
    "pillarnet_lts_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pillarnet_lts_b200"] = _mod
_spec.loader.exec_module(_mod)
