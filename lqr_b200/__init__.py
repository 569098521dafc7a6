"""Import alias for the product package.

The product lives in the directory ``lqr.jl_b200/`` (the name the build contract fixes); a dot is
not legal in a Python package name, so this thin alias package extends its ``__path__`` to that
directory: ``import lqr_b200.problems`` loads ``lqr.jl_b200/problems.py``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "lqr.jl_b200")
__path__.append(_PKG_DIR)

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
