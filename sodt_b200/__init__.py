"""Importable alias of the product package.

The package directory is named ``small-object-detection-transformers_b200`` (not a valid Python
identifier); this shim makes it importable as ``sodt_b200`` by pointing ``__path__`` at it.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "small-object-detection-transformers_b200")
__path__.insert(0, _PKG_DIR)

with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
