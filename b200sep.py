"""Importable alias of the package directory ``audio-classification_b200/`` (the hyphen in the
repository-mandated directory name cannot appear in a Python ``import`` statement)."""
import importlib as _il
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _il.import_module("audio-classification_b200")
_sys.modules[__name__] = _pkg
