"""Importable alias for the package directory ``stablelinearalgebra.jl_b200/`` (its name, fixed by the
project layout, is not a valid Python identifier): ``import sla_b200`` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stablelinearalgebra.jl_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
