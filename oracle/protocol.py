"""TEST INFRASTRUCTURE ONLY -- re-export of the product's data synthesiser (capf_b200/synth.py) under the name
the oracle-side scripts use.  (The dependency runs oracle -> product data synthesis only, never the reverse.)"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from capf_b200.synth import make_inputs, make_weights  # noqa: E402,F401
