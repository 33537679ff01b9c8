"""Importable alias of the `multimodal-dynamics_b200/` package directory (a hyphen cannot appear in
a Python module name).  `import mmdyn_b200` executes multimodal-dynamics_b200/__init__.py with this
module's `__path__` pointing there, so `mmdyn_b200.pytorch.models.vae` etc. resolve into it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "multimodal-dynamics_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
