"""CPU: pin the oracle restatement (oracle/capf_oracle.py) to the reference.

1. against the committed fixtures produced by the unmodified reference (tests/golden, oracle/gen_golden.py);
2. against the live reference when /root/reference is mounted (authoring container only);
3. the element-wise numpy bilinear gather against ATen's F.grid_sample for both padding modes.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import capf_oracle
import protocol
import ref_import
from conftest import build_case_model, golden_cases, load_golden, rel_l2

CASES = golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_golden(case):
    g = load_golden(case["name"])
    _, w, cfg = build_case_model(case["backbone"], case["weight_seed"])
    images, kp2d, crop = protocol.make_inputs(case["B"], case["H"], case["W"], case["input_seed"])
    # RNG guard: same inputs as the generator saw
    chk = g["images_checksum"]
    assert abs(images.double().sum().item() - chk[0]) < 1e-6 * max(1.0, abs(chk[1]))
    assert np.array_equal(kp2d.numpy(), g["kp2d"]) and np.array_equal(crop.numpy(), g["crop_in"])
    tr = {}
    out = capf_oracle.ca_pf_forward(w, case["backbone"], cfg.model.backbone, images, kp2d, crop, trace=tr)
    assert np.array_equal(crop.numpy(), g["crop_after"]), "in-place crop normalisation (conpose.py:34-35)"
    # fp32 reductions may be scheduled differently across thread counts: allow a few ulp
    assert rel_l2(out, g["out"]) < 2e-6
    for l, f in enumerate(tr["features"]):
        assert tuple(f.shape) == tuple(g[f"feat{l}_shape"])
        got = f.reshape(-1)[torch.from_numpy(g[f"feat{l}_idx"])]
        assert rel_l2(got, g[f"feat{l}_val"]) < 2e-6
        assert abs(f.double().abs().sum().item() - g[f"feat{l}_stats"][1]) < 1e-5 * g[f"feat{l}_stats"][1]
    assert rel_l2(tr["tokens_embed"], g["tokens_embed"]) < 2e-6
    assert rel_l2(tr["tokens_context"], g["tokens_context"]) < 2e-6
    assert rel_l2(tr["tokens_res"], g["tokens_res"]) < 2e-6
    assert rel_l2(tr["tokens_joint"], g["tokens_joint"]) < 2e-6


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("backbone,B,H,W", [("hrnet_32", 2, 64, 64), ("hrnet_48", 1, 64, 96), ("cpn", 1, 128, 96)])
def test_oracle_matches_live_reference(backbone, B, H, W):
    ref = ref_import.build_reference_model(backbone)
    w = protocol.make_weights([(k, tuple(v.shape)) for k, v in ref.state_dict().items()], 5)
    ref.load_state_dict(w, strict=True)
    images, kp2d, crop = protocol.make_inputs(B, H, W, 99)
    c1, c2 = crop.clone(), crop.clone()
    with torch.no_grad():
        a = ref(images, kp2d, c1)
    b = capf_oracle.ca_pf_forward(w, backbone, ref_import.reference_config(backbone).model.backbone, images, kp2d, c2)
    assert torch.equal(c1, c2)
    assert rel_l2(b, a) < 2e-6


@pytest.mark.parametrize("border", [False, True])
def test_numpy_gather_matches_aten(border):
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, 8, 9, 7, generator=g)
    grid = torch.rand(2, 50, 2, generator=g) * 2.6 - 1.3
    grid[0, 0] = torch.tensor([-1.0, -1.0]); grid[0, 1] = torch.tensor([1.0, 1.0]); grid[0, 2] = torch.tensor([0.0, 0.0])
    want = F.grid_sample(feat, grid.unsqueeze(-2), padding_mode="border" if border else "zeros",
                         align_corners=True).squeeze(-1).permute(0, 2, 1)
    got = capf_oracle.grid_sample_gather(feat, grid.numpy(), border)
    assert np.abs(got - want.numpy()).max() < 2e-6
    x0, y0, mask, w = capf_oracle.grid_sample_records(grid.numpy(), 9, 7, border)
    if border:
        assert ((x0 >= 0) & (x0 <= 6) & (y0 >= 0) & (y0 <= 8)).all()
        assert (mask & 1).all()                       # clipped: the nw corner is always inside
    assert np.allclose(w.sum(-1), 1.0, atol=1e-5)
