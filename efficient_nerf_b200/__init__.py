"""Importable alias of the `efficient-nerf_b200/` package directory (a hyphen is not a valid
Python identifier): `import efficient_nerf_b200` resolves every submodule from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "efficient-nerf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
