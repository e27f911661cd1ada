"""Importable alias for the package directory ``multimodal-long-transformer-2021_b200``.

The directory name required by the repo layout is not a Python identifier, so
``import mlt_b200`` (and ``mlt_b200.<submodule>``) resolve to the very same module
objects through a meta-path alias -- never a second copy.
"""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_ALIAS = __name__
_REAL = 'multimodal-long-transformer-2021_b200'
_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
  sys.path.insert(0, _ROOT)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):

  def find_spec(self, fullname, path=None, target=None):
    if fullname.startswith(_ALIAS + '.'):
      return importlib.util.spec_from_loader(fullname, self)
    return None

  def create_module(self, spec):
    return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

  def exec_module(self, module):
    pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
  sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
sys.modules[_ALIAS] = _pkg
