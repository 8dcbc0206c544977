"""Importable alias of the product package.

The product lives in ``dynamic-video-depth_b200/`` (the directory name the build contract asks for);
a hyphen is not a legal Python identifier, so ``import dvd_b200`` resolves here and this shim
re-points the package at that directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'dynamic-video-depth_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f
