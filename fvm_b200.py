"""Import shim: `import fvm_b200` loads the package in ./finitevolumemethod.jl_b200/ (whose
directory name, fixed by the project layout, is not an importable identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "finitevolumemethod.jl_b200")
_spec = importlib.util.spec_from_file_location("fvm_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fvm_b200"] = _mod
_spec.loader.exec_module(_mod)
