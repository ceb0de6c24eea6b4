"""Import alias: the product package lives in ``asr-ttl-mtl_b200/`` (not an importable name)."""
import os as _os

_pkg = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "asr-ttl-mtl_b200")
__path__.append(_pkg)
with open(_os.path.join(_pkg, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg, "__init__.py"), "exec"))
del _os, _pkg, _f
