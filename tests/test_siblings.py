"""GPU: CAPF_OP_CONV2D with OUTPUT SEGMENTS (i[20..23], i[14]) -- sibling convolutions of a HighResolutionModule's fuse layers
(pose_hrnet.py:235-277) as one GEMM over the Cout-concatenated weights, one dense tensor per sibling (program.fuse_siblings).
The merged launch must reproduce the separate convolutions BIT FOR BIT (same K order per output column), for every segment
layout the HRNet configs produce, 3x3 / stride 2 and 1x1, ragged tiles, both 16-bit types; and the whole forward with and
without the pass must agree exactly."""
import pytest
import torch

from capf_b200 import lib
from gpu_util import run_op

pytestmark = pytest.mark.gpu

# (name, N, H, W, Cin, k, stride, [(Cout, relu)...])
CASES = [
    ("hrnet32 branch0 x3 (s2)", 4, 64, 64, 32, 3, 2, [(64, 0), (32, 1), (32, 1)]),
    ("hrnet32 branch1 x2 (s2)", 4, 32, 32, 64, 3, 2, [(128, 0), (64, 1)]),
    ("hrnet32 branch2 x2 (1x1)", 4, 16, 16, 128, 1, 1, [(32, 0), (64, 0)]),
    ("hrnet32 branch3 x3 (1x1)", 4, 8, 8, 256, 1, 1, [(32, 0), (64, 0), (128, 0)]),
    ("hrnet48 branch0 x3 (s2)", 2, 96, 72, 48, 3, 2, [(96, 0), (48, 1), (48, 1)]),
    ("hrnet48 branch2 x2 (1x1)", 2, 24, 18, 192, 1, 1, [(48, 0), (96, 0)]),
    ("ragged rows (1x1)", 3, 7, 5, 64, 1, 1, [(16, 1), (48, 0), (16, 1), (32, 0)]),
    ("ragged box (s2)", 3, 30, 22, 32, 3, 2, [(32, 1), (96, 0)]),
    ("all relu", 2, 16, 16, 64, 1, 1, [(64, 1), (64, 1)]),
    ("none", 2, 16, 16, 64, 1, 1, [(128, 0), (128, 0)]),
]


def _conv_ints(N, H, W, Cin, Cout, k, stride, act):
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    return [N, H, W, Cin, Cout, k, k, stride, pad, Ho, Wo, act, lib.IMPL_TCGEN05], Ho, Wo


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_segmented_conv_is_bit_identical_to_separate_convs(case, dt):
    name, N, H, W, Cin, k, stride, sibs = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    K = k * k * Cin
    x = torch.randn(N, H, W, Cin, generator=g).to(dt).cuda()
    ws = [(torch.randn(c, K, generator=g) * K ** -0.5).to(dt).cuda() for c, _ in sibs]
    bs = [torch.randn(c, generator=g).cuda() for c, _ in sibs]
    # separate launches (per-tap kernel forced: the merged op runs on it)
    refs = []
    for (c, relu), w, b in zip(sibs, ws, bs):
        ints, Ho, Wo = _conv_ints(N, H, W, Cin, c, k, stride, lib.ACT_RELU if relu else lib.ACT_NONE)
        y = torch.full((N, Ho, Wo, c), float("nan"), dtype=dt, device="cuda")
        run_op(lib.OP_CONV2D, dt, dt, ints + [1], [], [x, w, b, None], [y])
        refs.append(y)
    # one launch
    ctot = sum(c for c, _ in sibs)
    any_relu = any(r for _, r in sibs)
    ints, Ho, Wo = _conv_ints(N, H, W, Cin, ctot, k, stride, lib.ACT_RELU if any_relu else lib.ACT_NONE)
    ints += [0] * (24 - len(ints))
    ints[14] = sum(1 << n for n, (_, r) in enumerate(sibs) if any_relu and not r)
    ints[20] = len(sibs)
    for n, (c, _) in enumerate(sibs[:-1]):
        ints[21 + n] = c
    outs = [torch.full((N, Ho, Wo, c), float("nan"), dtype=dt, device="cuda") for c, _ in sibs]
    run_op(lib.OP_CONV2D, dt, dt, ints, [], [x, torch.cat(ws, 0).contiguous(), torch.cat(bs).contiguous(), None], outs)
    for n, (y, r) in enumerate(zip(outs, refs)):
        assert not torch.isnan(y.float()).any(), f"segment {n}: unwritten elements"
        assert torch.equal(y, r), f"segment {n}: max diff {(y.float() - r.float()).abs().max().item():.3e}"
    # and against plain fp32 PyTorch
    xf = x.float().permute(0, 3, 1, 2)
    for (c, relu), w, b, y in zip(sibs, ws, bs, outs):
        wf = w.float().view(c, k, k, Cin).permute(0, 3, 1, 2)
        ref = torch.nn.functional.conv2d(xf, wf, b, stride, k // 2)
        if relu:
            ref = ref.relu()
        rel = float((y.float().permute(0, 3, 1, 2) - ref).norm() / ref.norm())
        assert rel < (1.5e-3 if dt == torch.float16 else 8e-3)


@pytest.mark.parametrize("bad", ["residual", "gelu", "width", "count", "null_out"])
def test_segmented_conv_rejects_bad_requests(bad):
    dt = torch.float16
    N, H, W, Cin = 2, 8, 8, 64
    x = torch.zeros(N, H, W, Cin, dtype=dt, device="cuda")
    w = torch.zeros(96, Cin, dtype=dt, device="cuda")
    outs = [torch.zeros(N, H, W, 32, dtype=dt, device="cuda"), torch.zeros(N, H, W, 64, dtype=dt, device="cuda")]
    ints, _, _ = _conv_ints(N, H, W, Cin, 96, 1, 1, lib.ACT_GELU if bad == "gelu" else lib.ACT_NONE)
    ints += [0] * (24 - len(ints))
    ints[20], ints[21] = (5 if bad == "count" else 2), (24 if bad == "width" else 32)
    res = torch.zeros(N, H, W, 96, dtype=dt, device="cuda") if bad == "residual" else None
    with pytest.raises(lib.CapfError):
        run_op(lib.OP_CONV2D, dt, dt, ints, [], [x, w, None, res], [outs[0], None if bad == "null_out" else outs[1]])


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_forward_with_and_without_sibling_fusion_is_identical(precision, monkeypatch):
    import capf_b200
    import protocol
    cfg = capf_b200.make_config("hrnet_32")
    B, H, W = 3, 128, 96
    images, kp2d, crop = protocol.make_inputs(B, H, W, 11)
    outs, launches = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("CAPF_FUSE_SIBLINGS", flag)
        model = capf_b200.CA_PF(cfg, precision=precision).eval()
        w = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 0)
        model.load_state_dict(w, strict=True)
        model = model.cuda()
        with torch.no_grad():
            outs.append(model(images.cuda(), kp2d.cuda(), crop.clone().cuda()).clone())
        plan = next(iter(model._plans.values()))[0]
        launches.append(plan.num_launches)
        if flag == "1":
            assert any("output segments" in plan.op_kernel(k) for k in range(plan.num_launches))
    assert launches[1] - launches[0] == 20
    assert torch.equal(outs[0], outs[1])
