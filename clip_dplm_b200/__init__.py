"""Import shim: ``import clip_dplm_b200`` resolves to the package that lives in ``clip-dplm_b200/``
(the directory name the project layout prescribes is not a valid Python identifier)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "clip-dplm_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
