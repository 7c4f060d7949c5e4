"""Importable alias of the package directory ``globecom2020-resourceallocationgnn_b200``
(its name carries the reference repository's hyphens, which ``import`` cannot spell)."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _importlib.import_module("globecom2020-resourceallocationgnn_b200")
_sys.modules[__name__] = _pkg
