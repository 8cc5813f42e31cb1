"""Import alias: the package directory is `finetools.jl_b200/` (not a legal Python identifier), so this
three-line package splices that directory into its own search path.  All code lives over there."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "finetools.jl_b200"))
from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
