"""Importable alias for the package directory ``multimodal-long-transformer-2021_b200``.

The directory name required by the repo layout is not a Python identifier, so
``import mlt_b200`` resolves to it through importlib.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
  sys.path.insert(0, _ROOT)
_pkg = importlib.import_module('multimodal-long-transformer-2021_b200')
sys.modules[__name__] = _pkg
