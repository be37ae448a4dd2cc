"""Importable alias for the product package, whose directory name (``contextaware-poseformer_b200``)
is not a valid Python identifier.  ``import capf_b200`` == that package."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "contextaware-poseformer_b200"))
from ._pkg import *  # noqa: F401,F403  (re-export of contextaware-poseformer_b200/_pkg.py)
from ._pkg import __all__  # noqa: F401
