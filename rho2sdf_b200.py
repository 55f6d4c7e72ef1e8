"""Import shim: the product package lives in the directory `rho2sdf.jl_b200/` (a name Python cannot import directly).
`import rho2sdf_b200 as r2s` loads that package and re-exports its public names."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rho2sdf.jl_b200")
_spec = importlib.util.spec_from_file_location("rho2sdf_jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules.setdefault("rho2sdf_jl_b200", _mod)
_spec.loader.exec_module(_mod)
globals().update({k: getattr(_mod, k) for k in _mod.__all__})
__all__ = list(_mod.__all__)
