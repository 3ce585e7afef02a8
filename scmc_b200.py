"""Import shim: exposes the package directory `semiclassicalmc.jl_b200/` (not a legal Python identifier)
under the importable name `scmc_b200`.  `import scmc_b200` returns that package."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "semiclassicalmc.jl_b200")
_spec = _ilu.spec_from_file_location("scmc_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["scmc_b200"] = _mod
_spec.loader.exec_module(_mod)
