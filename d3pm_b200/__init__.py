"""Importable alias of the package directory `gif-synthesis-with-discrete-diffusion_b200/`.

The product package lives in a directory whose name (taken from the upstream repository) is not a
valid Python identifier.  This alias points its `__path__` there, so that
`import d3pm_b200.ops`, `from d3pm_b200 import FusedDiffusionTransformer` … resolve to the files
of that directory.  No code lives here.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "gif-synthesis-with-discrete-diffusion_b200")
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _f
