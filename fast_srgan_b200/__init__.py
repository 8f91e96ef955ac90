"""Import shim: `import fast_srgan_b200` -> the package directory `fast-srgan_b200/`
(a hyphen is not a valid module name, the task layout fixes the directory name)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "fast-srgan_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
