"""Importable alias for the ``f-lite_b200/`` package directory (a hyphen is not a valid module name).

``import flite_b200`` executes ``f-lite_b200/__init__.py`` with ``__path__`` pointing at that directory,
so ``flite_b200.model``, ``flite_b200.ops`` ... resolve to the files under ``f-lite_b200/``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "f-lite_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"), globals())
