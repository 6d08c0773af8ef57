"""Importable alias of the package directory ``nonlinpdes-gpsolver_b200/`` (a hyphen cannot be
imported).  All code lives there; this shim only points the package path at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "nonlinpdes-gpsolver_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
