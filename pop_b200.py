"""Import alias: the package directory is `piecewiseorthogonalpolynomials.jl_b200/` (a dot in
the name is not importable with a plain `import`), so `import pop_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "piecewiseorthogonalpolynomials.jl_b200")
_spec = importlib.util.spec_from_file_location("pop_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pop_b200"] = _mod
_spec.loader.exec_module(_mod)
