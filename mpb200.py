"""Import shim: the package directory is named `motionplanning.jl_b200` (not a valid Python
identifier), so `import mpb200` loads it from that directory under the module name `mpb200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "motionplanning.jl_b200")
_spec = importlib.util.spec_from_file_location("mpb200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mpb200"] = _mod
_spec.loader.exec_module(_mod)
