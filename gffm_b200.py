"""Import shim: registers the package that lives in `gpufinitefieldmatrices.jl_b200/` (a directory name that is
not a valid Python identifier) under the importable name `gffm_b200`."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_HERE, "gpufinitefieldmatrices.jl_b200")
_NAME = "gffm_b200"

_spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _mod
_spec.loader.exec_module(_mod)
