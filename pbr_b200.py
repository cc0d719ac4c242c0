"""Import alias for the package directory `physically-based-rendering_b200/` (its name is not a
valid Python identifier).  `import pbr_b200` gives the package; submodules are reachable as
`pbr_b200.capi`, `pbr_b200.scenes`, ..."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "physically-based-rendering_b200")
_spec = importlib.util.spec_from_file_location(
    "pbr_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pbr_b200"] = _mod
_spec.loader.exec_module(_mod)
