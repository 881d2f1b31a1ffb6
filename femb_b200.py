"""Import shim: ``import femb_b200`` loads the package in ``extendablefembase.jl_b200/`` (a directory
name with a dot cannot be imported by name)."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "extendablefembase.jl_b200")
_spec = importlib.util.spec_from_file_location("femb_b200", os.path.join(_d, "__init__.py"), submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["femb_b200"] = _mod
_spec.loader.exec_module(_mod)
