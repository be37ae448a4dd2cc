import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by `pytest -m gpu` on the GPU box)")
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))


def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)["cases"]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "state_dict_manifest.json")) as f:
        return json.load(f)


_MODEL_CACHE = {}


def build_case_model(backbone, weight_seed, precision="fp32"):
    """Our CA_PF with the protocol weights of (backbone, seed) loaded (CPU parameters)."""
    import capf_b200
    import protocol
    key = (backbone, weight_seed, precision)
    if key not in _MODEL_CACHE:
        cfg = capf_b200.make_config(backbone)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            m = capf_b200.CA_PF(cfg, precision=precision).eval()
        w = protocol.make_weights([(k, tuple(v.shape)) for k, v in m.state_dict().items()], weight_seed)
        m.load_state_dict(w, strict=True)
        _MODEL_CACHE[key] = (m, w, cfg)
    return _MODEL_CACHE[key]
