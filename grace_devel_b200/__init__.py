"""Import alias: the product package lives in the directory `grace-devel_b200/`
(not a valid Python identifier), so this stub extends its search path there."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "grace-devel_b200"))

from ._api import *  # noqa: F401,F403,E402
from ._api import __all__  # noqa: F401,E402
