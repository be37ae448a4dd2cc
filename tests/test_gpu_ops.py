"""GPU: per-operator parity through capf_op_run against plain PyTorch fp32 (CPU) references."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from capf_b200 import lib
from conftest import rel_l2
from gpu_util import run_op

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gen(seed=0):
    return torch.Generator().manual_seed(seed)


CONV_SHAPES = [
    # N, H, W, Cin, Cout, k, stride
    (2, 17, 13, 3, 64, 3, 2),      # HRNet stem: Cin=3 (generic path), odd sizes
    (1, 32, 24, 3, 64, 7, 2),      # CPN stem 7x7
    (3, 16, 12, 32, 32, 3, 1),     # HRNet-32 branch conv
    (2, 16, 16, 64, 128, 3, 2),    # fuse down-conv
    (2, 8, 8, 256, 32, 1, 1),      # fuse 1x1
    (1, 9, 7, 48, 96, 3, 2),       # HRNet-48 channel counts (not multiples of 32/64)
    (2, 8, 6, 512, 2048, 1, 1),    # ResNet-50 expansion
    (5, 1, 1, 640, 3, 1, 1),       # head Linear(640 -> 3): Cout=3 (generic path)
    (37, 1, 1, 128, 48, 1, 1),     # attention_weights|sampling_offsets GEMM, ragged M
    (2, 8, 8, 64, 64, 1, 2),       # ResNet 1x1 stride-2 downsample
]


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=[str(s) for s in CONV_SHAPES])
@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
@pytest.mark.parametrize("act,use_res", [(lib.ACT_RELU, True), (lib.ACT_GELU, False), (lib.ACT_NONE, True)])
def test_conv2d_simt(shape, dt, act, use_res):
    N, H, W, Cin, Cout, k, stride = shape
    g = _gen(1)
    pad = k // 2
    x = torch.randn(N, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = torch.randn(N, Ho, Wo, Cout, generator=g) if use_res else None
    xq, wq = x.to(dt).float(), w.to(dt).float()
    resq = res.to(dt).float() if use_res else None
    y = F.conv2d(xq.permute(0, 3, 1, 2), wq, bias, stride, pad).permute(0, 2, 3, 1)
    if act == lib.ACT_GELU:
        y = F.gelu(y)
    if use_res:
        y = y + resq
    if act == lib.ACT_RELU:
        y = F.relu(y)
    wp = w.permute(2, 3, 1, 0).reshape(-1, Cout).contiguous().to(dt).to(DEV)
    out = torch.empty(N, Ho, Wo, Cout, dtype=dt, device=DEV)
    run_op(lib.OP_CONV2D, dt, dt, [N, H, W, Cin, Cout, k, k, stride, pad, Ho, Wo, act, lib.IMPL_SIMT], [],
           [x.to(dt).to(DEV), wp, bias.to(DEV), res.to(dt).to(DEV) if use_res else None], [out])
    assert rel_l2(out.float().cpu(), y) < (2e-6 if dt == torch.float32 else 1.5e-3)


@pytest.mark.parametrize("shape", [(2, 17, 13), (3, 64, 48), (1, 256, 256), (2, 384, 288), (5, 40, 260), (300, 8, 8)], ids=str)
@pytest.mark.parametrize("odt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("ks", [3, 7], ids=["hrnet3x3", "cpn7x7"])
def test_stem_conv_f32_image_to_16bit(shape, odt, ks):
    """Backbone conv1 (HRNet pose_hrnet.py:321-322: 3x3/s2; CPN resnet.py:100: 7x7/s2): fp32 NHWC image in, folded-BN
    3 -> 64 + ReLU, 16-bit out: the tensor-pipe stem of capf_stem.cu (W % 4 == 0; one or several 128-pixel tiles per
    output row, ragged last tile, more tiles than CTAs) and the CUDA-core kernels of capf_simt.cu (odd widths)."""
    N, H, W = shape
    g = _gen(11)
    pad = ks // 2
    x = torch.randn(N, H, W, 3, generator=g)
    w = torch.randn(64, 3, ks, ks, generator=g) / (3 * ks * ks) ** 0.5
    bias = torch.randn(64, generator=g)
    Ho, Wo = (H + 2 * pad - ks) // 2 + 1, (W + 2 * pad - ks) // 2 + 1
    y = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w, bias, 2, pad)).permute(0, 2, 3, 1)
    wp = w.permute(2, 3, 1, 0).reshape(-1, 64).contiguous().to(DEV)
    out = torch.full((N, Ho, Wo, 64), float("nan"), dtype=odt, device=DEV)
    run_op(lib.OP_CONV2D, torch.float32, odt, [N, H, W, 3, 64, ks, ks, 2, pad, Ho, Wo, lib.ACT_RELU, lib.IMPL_SIMT], [],
           [x.to(DEV), wp, bias.to(DEV), None], [out])
    # 3x3: hi/lo split operands -> only the output rounding remains; 7x7: single 16-bit operands (input/weight rounding)
    tol = (6e-4 if odt == torch.float16 else 4e-3) if ks == 3 else (9e-4 if odt == torch.float16 else 6e-3)
    assert rel_l2(out.float().cpu(), y) < tol


def test_conv2d_rejects_bad_arguments():
    x = torch.zeros(1, 4, 4, 16, device=DEV)
    with pytest.raises(lib.CapfError, match="Ho/Wo"):
        run_op(lib.OP_CONV2D, torch.float32, torch.float32, [1, 4, 4, 16, 16, 3, 3, 1, 1, 5, 4, 0, 0], [], [x, x, None, None], [x])
    with pytest.raises(lib.CapfError, match="null"):
        run_op(lib.OP_CONV2D, torch.float32, torch.float32, [1, 4, 4, 16, 16, 3, 3, 1, 1, 4, 4, 0, 0], [], [x, None, None, None], [x])


@pytest.mark.parametrize("dt", [torch.float32, torch.float16])
def test_fuse_sum(dt):
    g = _gen(2)
    N, H, W, C = 2, 16, 8, 48
    terms = [torch.randn(N, H >> s, W >> s, C, generator=g).to(dt) for s in (0, 1, 2, 3)]
    want = terms[0].float()
    for s, t in zip((1, 2, 3), terms[1:]):
        want = want + t.float().repeat_interleave(1 << s, 1).repeat_interleave(1 << s, 2)
    want = F.relu(want)
    out = torch.empty(N, H, W, C, dtype=dt, device=DEV)
    run_op(lib.OP_FUSE_SUM, dt, dt, [N, H, W, C, 4, 0, 1, 2, 3, 1], [], [t.to(DEV) for t in terms], [out])
    assert rel_l2(out.float().cpu(), want.to(dt).float()) < (1e-6 if dt == torch.float32 else 1e-3)


@pytest.mark.parametrize("shape", [(2, 16, 8, 48, (0, 1, 2, 3)), (5, 64, 64, 32, (0, 1, 2, 3)), (3, 32, 32, 64, (0, 0, 1, 2)), (300, 8, 8, 256, (0, 0, 0, 0)),
                                   (2, 96, 72, 48, (0, 1, 2)), (3, 24, 18, 192, (0, 0, 0)), (1, 4, 4, 16, (0, 1)), (7, 16, 16, 128, (0, 0, 0, 1))],
                         ids=lambda s: "x".join(str(v) for v in s[:4]) + f"_t{len(s[4])}")
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16, torch.float32], ids=["f16", "bf16", "f32"])
def test_fuse_sum_rows_kernel_is_exact(shape, dt):
    """HighResolutionModule fuse (pose_hrnet.py:294-301): fp32 sum of the terms in order, nearest-upsampled on the fly, ReLU, one
    rounding -- exactly the torch expression, for rows of 256 vectors (the benchmark shapes), longer and shorter rows, channel
    counts that are not powers of two, more rows than blocks and 2..4 terms."""
    N, H, W, C, shifts = shape
    g = _gen(sum(shape[:4]))
    terms = [torch.randn(N, H >> s, W >> s, C, generator=g).to(dt) for s in shifts]
    want = None
    for s, t in zip(shifts, terms):
        u = t.float().repeat_interleave(1 << s, 1).repeat_interleave(1 << s, 2) if s else t.float()
        want = u if want is None else want + u
    want = F.relu(want).to(dt)
    out = torch.full((N, H, W, C), float("nan"), dtype=dt, device=DEV)
    run_op(lib.OP_FUSE_SUM, dt, dt, [N, H, W, C, len(shifts)] + list(shifts) + [0] * (4 - len(shifts)) + [1], [], [t.to(DEV) for t in terms], [out])
    assert torch.equal(out.cpu(), want)


def test_maxpool_and_bilinear():
    g = _gen(3)
    x = torch.randn(2, 13, 10, 64, generator=g)
    Ho, Wo = (13 - 1) // 2 + 1, (10 - 1) // 2 + 1
    out = torch.empty(2, Ho, Wo, 64, device=DEV)
    run_op(lib.OP_MAXPOOL, torch.float32, torch.float32, [2, 13, 10, 64, Ho, Wo], [], [x.to(DEV)], [out])
    assert torch.equal(out.cpu(), F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1))
    for (Ho, Wo) in ((26, 20), (64, 48), (13, 10), (1, 1)):
        out = torch.empty(2, Ho, Wo, 64, device=DEV)
        run_op(lib.OP_BILINEAR, torch.float32, torch.float32, [2, 13, 10, 64, Ho, Wo], [], [x.to(DEV)], [out])
        want = F.interpolate(x.permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        assert (out.cpu() - want).abs().max() < 2e-5


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_bilinear_16bit_vector_path(dt):
    """nn.Upsample(bilinear, align_corners=True) of the CPN heads (globalNet.py:40, refineNet.py:61) on 16-bit NHWC
    tensors: 8-channel (16-byte) vector kernel vs fp32 F.interpolate on the same rounded input; up- and down-sampling."""
    g = _gen(13)
    x = torch.randn(3, 16, 12, 256, generator=g).to(dt)
    for (Ho, Wo) in ((32, 24), (64, 48), (8, 6), (16, 12), (1, 1)):
        out = torch.full((3, Ho, Wo, 256), float("nan"), dtype=dt, device=DEV)
        run_op(lib.OP_BILINEAR, dt, dt, [3, 16, 12, 256, Ho, Wo], [], [x.to(DEV)], [out])
        want = F.interpolate(x.float().permute(0, 3, 1, 2), size=(Ho, Wo), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
        assert rel_l2(out.float().cpu(), want) < (5e-4 if dt == torch.float16 else 4e-3)


@pytest.mark.parametrize("D,period", [(128, 0), (640, 0), (128, 34), (96, 0), (64, 17), (480, 0), (320, 0)])
@pytest.mark.parametrize("rows", [136, 137, 3])
def test_layernorm(D, period, rows):
    """Wide rows (one warp per row), narrow rows (D <= 128: four rows per warp, ragged row counts), the x + x0 broadcast
    of DeformableBlock (:120), and the MPI-INF-3DHP widths that are not multiples of 128."""
    g = _gen(4)
    x = torch.randn(rows, D, generator=g) * 3 + 1
    gamma, beta = torch.rand(D, generator=g) + 0.5, torch.randn(D, generator=g)
    x0 = torch.randn(period, D, generator=g) if period else None
    xin = x + (x0.repeat((rows + period - 1) // period, 1)[:rows] if period else 0)
    for eps in (1e-5, 1e-6):
        want = F.layer_norm(xin, (D,), gamma, beta, eps)
        out = torch.empty(rows, D, device=DEV)
        run_op(lib.OP_LAYERNORM, torch.float32, torch.float32, [rows, D, period], [eps],
               [x.to(DEV), gamma.to(DEV), beta.to(DEV), x0.to(DEV) if period else None], [out])
        assert (out.cpu() - want).abs().max() < 5e-6


@pytest.mark.parametrize("seq,heads,hd,groups", [(5, 8, 16, 51), (17, 8, 80, 7)])
def test_attention(seq, heads, hd, groups):
    g = _gen(5)
    D = heads * hd
    if seq == 5:      # level-major stream: token t of group g at row g + t*groups
        ts, gs = groups, 1
    else:
        ts, gs = 1, seq
    rows = groups * seq
    qkv = torch.randn(rows, 3 * D, generator=g)
    idx = (torch.arange(groups).view(-1, 1) * gs + torch.arange(seq).view(1, -1) * ts).reshape(-1)
    x = qkv[idx].view(groups, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    a = ((x[0] @ x[1].transpose(-2, -1)) * hd ** -0.5).softmax(-1)
    want = torch.empty(rows, D)
    want[idx] = (a @ x[2]).transpose(1, 2).reshape(groups * seq, D)
    out = torch.empty(rows, D, device=DEV)
    run_op(lib.OP_ATTENTION, torch.float32, torch.float32, [groups, seq, heads, hd, ts, gs], [hd ** -0.5], [qkv.to(DEV)], [out])
    assert (out.cpu() - want).abs().max() < 2e-5


@pytest.mark.parametrize("seq,heads,hd,groups", [(5, 8, 16, 51), (5, 8, 16, 4352), (17, 8, 80, 7), (17, 8, 80, 256), (17, 4, 32, 5)])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_attention_16bit(seq, heads, hd, groups, dt):
    """16-bit qkv in / out (the lifter's fp16 / bf16 modes): the register/broadcast fast paths for (17 x 8 x 80) and
    (5 x 8 x 16) incl. ragged warp tails, and the generic kernel for any other shape, against fp32 PyTorch on the same
    rounded inputs."""
    g = _gen(7)
    D = heads * hd
    ts, gs = (groups, 1) if seq == 5 else (1, seq)
    rows = groups * seq
    qkv = torch.randn(rows, 3 * D, generator=g).to(dt)
    idx = (torch.arange(groups).view(-1, 1) * gs + torch.arange(seq).view(1, -1) * ts).reshape(-1)
    x = qkv.float()[idx].view(groups, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    a = ((x[0] @ x[1].transpose(-2, -1)) * hd ** -0.5).softmax(-1)
    want = torch.empty(rows, D)
    want[idx] = (a @ x[2]).transpose(1, 2).reshape(groups * seq, D)
    out = torch.full((rows, D), float("nan"), dtype=dt, device=DEV)
    run_op(lib.OP_ATTENTION, dt, dt, [groups, seq, heads, hd, ts, gs], [hd ** -0.5], [qkv.to(DEV)], [out])
    assert rel_l2(out.float().cpu(), want) < (1e-3 if dt == torch.float16 else 6e-3)
    if seq == 17 and hd == 80:
        # warp-MMA kernel (mma.sync m16n8k16, P split into hi + lo): fp32-class arithmetic, so what is left is the rounding of the output
        rounding = rel_l2(want.to(dt).float(), want)
        assert rel_l2(out.float().cpu(), want) < 1.15 * rounding + 1e-6


def test_samplers_match_aten_values():
    """Blend values of both gathers vs F.grid_sample (zeros / border), incl. points outside the map and on +-1."""
    import capf_oracle
    g = _gen(6)
    B, J = 3, 17
    geo = [(16, 12, 32), (8, 6, 64), (4, 3, 128), (2, 2, 256)]
    maps = [torch.randn(B, h, w, c, generator=g) for h, w, c in geo]
    ref = torch.rand(B * J, 2, generator=g) * 2.8 - 1.4
    ref[0] = torch.tensor([-1.0, -1.0]); ref[1] = torch.tensor([1.0, 1.0]); ref[2] = torch.tensor([0.0, 0.0])
    offs, goffs = [0], [0]
    for h, w, c in geo:
        offs.append(offs[-1] + B * J * c)
        goffs.append(goffs[-1] + B * J * 4 * c)
    flat_geo = [v for t in geo for v in t]
    out = torch.empty(offs[-1], device=DEV)
    rec = torch.empty(4, B * J, 8, dtype=torch.int32, device=DEV)
    run_op(lib.OP_REF_SAMPLE, torch.float32, torch.float32, [B, J, 4] + flat_geo + offs[:4], [],
           [ref.to(DEV)] + [m.to(DEV) for m in maps], [out, rec])
    for l, (h, w, c) in enumerate(geo):
        want = F.grid_sample(maps[l].permute(0, 3, 1, 2), ref.view(B, J, 1, 2), align_corners=True).squeeze(-1).permute(0, 2, 1)
        got = out[offs[l]:offs[l + 1]].view(B, J, c).cpu()
        assert (got - want).abs().max() < 5e-6
        x0, y0, m, _ = capf_oracle.grid_sample_records(ref.numpy(), h, w, border=False)
        r = rec[l].cpu().numpy()
        assert np.array_equal(r[:, 0], x0) and np.array_equal(r[:, 1], y0) and np.array_equal(r[:, 2], m)
    ow = torch.randn(4 * B * J, 48, generator=g)
    gout = torch.empty(goffs[-1], device=DEV)
    run_op(lib.OP_DEFORM_SAMPLE, torch.float32, torch.float32, [B, J, 4] + flat_geo + goffs[:4], [],
           [ref.to(DEV)] + [m.to(DEV) for m in maps] + [ow.to(DEV)], [gout, None])
    owv = ow.view(4, B, J, 48)
    wts = owv[..., :16].reshape(4, B, J, 4, 4).softmax(-1)
    pos = owv[..., 16:].reshape(4, B, J, 16, 2).tanh() + ref.view(1, B, J, 1, 2)
    for l, (h, w, c) in enumerate(geo):
        s = F.grid_sample(maps[l].permute(0, 3, 1, 2), pos[l], padding_mode="border", align_corners=True).permute(0, 2, 3, 1)
        want = (s.reshape(B, J, 4, 4, c) * wts[l].unsqueeze(-1)).sum(-2)
        got = gout[goffs[l]:goffs[l + 1]].view(B, J, 4, c).cpu()
        assert (got - want).abs().max() < 2e-5


@pytest.mark.parametrize("geo", [[(16, 12, 32), (8, 6, 64), (4, 3, 128), (2, 2, 256)], [(24, 18, 48), (12, 9, 96), (6, 5, 192), (3, 3, 384)],
                                 [(16, 12, 256)] * 4], ids=["hrnet32", "hrnet48", "cpn"])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_deform_sample_16bit_fast_path(geo, dt):
    """The 16-bit deformable gather (one warp per (frame, joint, level), 16-byte loads): values against fp32
    F.grid_sample(border) on the same rounded maps, corner records bit-identical to the ATen restatement and to the fp32
    kernel -- for power-of-two channel counts, HRNet-48's 48/96/192/384 (idle lanes, two passes) and CPN's 4 x 256."""
    import capf_oracle
    g = _gen(21)
    B, J = 5, 17
    maps = [torch.randn(B, h, w, c, generator=g).to(dt) for h, w, c in geo]
    ref = torch.rand(B * J, 2, generator=g) * 2.8 - 1.4
    ref[0] = torch.tensor([-1.0, -1.0]); ref[1] = torch.tensor([1.0, 1.0]); ref[2] = torch.tensor([0.0, 0.0])
    goffs = [0]
    for h, w, c in geo:
        goffs.append(goffs[-1] + B * J * 4 * c)
    flat_geo = [v for t in geo for v in t]
    ow = torch.randn(4 * B * J, 48, generator=g)
    gout = torch.full((goffs[-1],), float("nan"), dtype=dt, device=DEV)
    rec = torch.empty(4, B * J, 16, 8, dtype=torch.int32, device=DEV)
    run_op(lib.OP_DEFORM_SAMPLE, dt, dt, [B, J, 4] + flat_geo + goffs[:4], [],
           [ref.to(DEV)] + [m.to(DEV) for m in maps] + [ow.to(DEV)], [gout, rec])
    rec32 = torch.empty_like(rec)
    g32 = torch.empty(goffs[-1], device=DEV)
    run_op(lib.OP_DEFORM_SAMPLE, torch.float32, torch.float32, [B, J, 4] + flat_geo + goffs[:4], [],
           [ref.to(DEV)] + [m.float().to(DEV) for m in maps] + [ow.to(DEV)], [g32, rec32])
    assert torch.equal(rec, rec32)
    owv = ow.view(4, B, J, 48)
    wts = owv[..., :16].reshape(4, B, J, 4, 4).softmax(-1)
    pos = owv[..., 16:].reshape(4, B, J, 16, 2).tanh() + ref.view(1, B, J, 1, 2)
    for l, (h, w, c) in enumerate(geo):
        s = F.grid_sample(maps[l].float().permute(0, 3, 1, 2), pos[l], padding_mode="border", align_corners=True).permute(0, 2, 3, 1)
        want = (s.reshape(B, J, 4, 4, c) * wts[l].unsqueeze(-1)).sum(-2)
        got = gout[goffs[l]:goffs[l + 1]].view(B, J, 4, c).float().cpu()
        assert rel_l2(got, want) < (6e-4 if dt == torch.float16 else 5e-3)
        # index path: records recomputed by the ATen restatement from the exact fp32 positions the kernel sampled
        r = rec[l].cpu().numpy().reshape(-1, 8)
        pxy = np.ascontiguousarray(r[:, 4:6]).view(np.float32)
        x0r, y0r, mr, _ = capf_oracle.grid_sample_records(pxy, h, w, border=True)
        assert np.array_equal(r[:, 0], x0r) and np.array_equal(r[:, 1], y0r) and np.array_equal(r[:, 2], mr)


def test_layernorm_with_fused_head_projection():
    """head = LayerNorm(640, eps 1e-5) + Linear(640 -> 3) (pose_dformer.py:205-208,240) as one op."""
    g = _gen(9)
    rows, D = 53, 640
    x = torch.randn(rows, D, generator=g) * 3 + 0.5
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    Wp, bp = torch.randn(3, D, generator=g) / D ** 0.5, torch.randn(3, generator=g)
    want = F.linear(F.layer_norm(x, (D,), gamma, beta, 1e-5), Wp, bp)
    out = torch.empty(rows, 3, device=DEV)
    run_op(lib.OP_LAYERNORM, torch.float32, torch.float32, [rows, D, 0, 3], [1e-5],
           [x.to(DEV), gamma.to(DEV), beta.to(DEV), None, Wp.to(DEV), bp.to(DEV)], [out])
    assert (out.cpu() - want).abs().max() < 2e-5


def test_token_glue_ops():
    g = _gen(7)
    B, J, D, S = 3, 17, 128, 5
    kp = torch.randn(B * J, 2, generator=g)
    Wc, bc, pos = torch.randn(D, 2, generator=g), torch.randn(D, generator=g), torch.randn(S, J, D, generator=g)
    X = torch.empty(S, B * J, D, device=DEV)
    run_op(lib.OP_EMBED_COORD, torch.float32, torch.float32, [B, J, D, S], [], [kp.to(DEV), Wc.to(DEV), bc.to(DEV), pos.to(DEV)], [X])
    want = pos.unsqueeze(1).expand(S, B, J, D).reshape(S, B * J, D).clone()
    want[0] += F.linear(kp, Wc, bc)
    assert (X.cpu() - want).abs().max() < 2e-6
    Y = torch.empty(B * J, S * D, device=DEV)
    run_op(lib.OP_LEVELS_TO_JOINT, torch.float32, torch.float32, [B * J, S, D], [], [X], [Y])
    assert torch.equal(Y.cpu(), X.cpu().permute(1, 0, 2).reshape(B * J, S * D))
    c = torch.rand(40, 2, generator=g) * 300 - 20
    cd = c.to(DEV)
    lib.check(lib.load().capf_crop_normalize(cd.data_ptr(), 40, torch.cuda.current_stream().cuda_stream))
    want = c.clone()
    want /= torch.tensor([96, 128])
    want -= torch.tensor([1, 1])
    assert torch.equal(cd.cpu(), want)
