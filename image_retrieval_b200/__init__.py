"""Import shim: `image_retrieval_b200` is the importable name of the package whose files live in
the contract-mandated directory `image-retrieval-_b200/` (a hyphen cannot appear in a Python
module name).  Everything is loaded from there."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "image-retrieval-_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
