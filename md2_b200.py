"""Importable alias for the package directory.

The package lives in ``digging-into-self-supervised-monocular-depth-estimation_b200/``
(the name the build contract fixes); hyphens are not legal in a Python module
name, so ``import md2_b200`` loads that directory as the package ``md2_b200``.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(
    os.path.dirname(os.path.abspath(__file__)),
    "digging-into-self-supervised-monocular-depth-estimation_b200",
)

_spec = importlib.util.spec_from_file_location(
    __name__,
    os.path.join(_PKG_DIR, "__init__.py"),
    submodule_search_locations=[_PKG_DIR],
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
