"""Import shim: the package directory is named ``incagg-gnn_b200`` (not a Python identifier).
``import incagg_gnn_b200`` loads that directory as a regular package under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "incagg-gnn_b200")
_spec = importlib.util.spec_from_file_location(
    "incagg_gnn_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["incagg_gnn_b200"] = _mod
_spec.loader.exec_module(_mod)
