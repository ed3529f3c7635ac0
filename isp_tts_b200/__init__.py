"""Importable alias of the product package, whose directory is named
`isp-tts_b200/` (a hyphen is not a legal Python identifier).  This file only
redirects the package search path there and runs the real __init__."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "isp-tts_b200")
__path__[:] = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
