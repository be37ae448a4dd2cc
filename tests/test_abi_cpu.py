"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/capf_b200.h declares, and
refuses to compute without a GPU (no CPU fallback anywhere in the product path)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from capf_b200 import lib
    L = lib.load()
    header = open(os.path.join(ROOT, "include", "capf_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(capf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(lib.exported_symbols())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in capf_b200.h but not exported"
    assert L.capf_abi_version() == lib.ABI_VERSION
    m = re.search(r"#define CAPF_ABI_VERSION (\d+)", header)
    assert int(m.group(1)) == lib.ABI_VERSION
    assert ctypes.sizeof(lib.CapfOp) == 16 + 24 * 4 + 16 + 6 * 8 + 4 * 8


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import capf_b200
    from capf_b200 import lib
    L = lib.load()
    op = lib.CapfOp()
    op.kind = lib.OP_CROP_NORMALIZE
    op.i[0] = 4
    buf = (ctypes.c_float * 8)()
    op.out[0] = ctypes.addressof(buf)
    rc = L.capf_op_run(ctypes.byref(op), 0, None)
    assert rc == -3 and b"no CPU path" in L.capf_last_error()          # CAPF_ERR_CUDA
    m = capf_b200.CA_PF(capf_b200.make_config("hrnet_32")).eval()
    with pytest.raises(lib.CapfError, match="no CPU path"):
        with torch.no_grad():
            m(torch.zeros(1, 64, 64, 3), torch.zeros(1, 17, 2), torch.zeros(1, 17, 2))
    with pytest.raises(lib.CapfError):
        m.backbone(torch.zeros(1, 3, 64, 64))
