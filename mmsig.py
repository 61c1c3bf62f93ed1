"""Import alias: `import mmsig` loads the package directory `multimodalmusig.jl_b200/`
(whose name is not a valid Python identifier) under the module name `mmsig`."""
import importlib.util
import os
import sys

_p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multimodalmusig.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "mmsig", os.path.join(_p, "__init__.py"), submodule_search_locations=[_p])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mmsig"] = _mod
_spec.loader.exec_module(_mod)
