"""Drop-in package: modules provided here shadow the reference's same-named modules; everything else in
the reference's same-named directory (found on sys.path) stays importable through this package."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
for _p in list(_sys.path):
    _d = _os.path.join(_os.path.abspath(_p or "."), *__name__.split("."))
    if _os.path.isdir(_d) and _os.path.abspath(_d) != _here and _d not in __path__:
        __path__.append(_d)
