"""Import alias: `import khronos_b200` loads the package in ./khronos.jl_b200/
(a directory name with a dot cannot be imported with a plain import statement)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "khronos.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "khronos_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["khronos_b200"] = _mod
_spec.loader.exec_module(_mod)
