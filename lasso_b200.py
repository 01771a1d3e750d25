"""Import alias: ``import lasso_b200`` loads the package in ``pytorch-lasso_b200/``.

The package directory carries the repository's name (with a hyphen), which is
not a valid Python identifier, so this one-file module registers it under the
importable name ``lasso_b200`` and replaces itself in ``sys.modules``.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "pytorch-lasso_b200")
_spec = _ilu.spec_from_file_location(
    "lasso_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["lasso_b200"] = _mod
_spec.loader.exec_module(_mod)
