"""Import shim: the package directory is named `cpu-raytracing-rt_b200` (after the reference
repo), which is not a valid Python identifier; `import rtb200` loads it under this name."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpu-raytracing-rt_b200")
_spec = importlib.util.spec_from_file_location(
    "rtb200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rtb200"] = _mod
_spec.loader.exec_module(_mod)
