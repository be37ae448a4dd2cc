from ._pkg import *  # noqa: F401,F403
from ._pkg import __all__  # noqa: F401
