"""Shape list + runner shared by tests/test_gpu_tc.py and tools/tc_probe.py: the tcgen05 implicit-GEMM kernel
(csrc/capf_tc.cu) against a plain PyTorch fp32 reference of the same operator on 16-bit-rounded inputs."""
import zlib

import torch
import torch.nn.functional as F

from capf_b200 import lib
from gpu_util import run_op

DEV = "cuda:0"

# name, (N, H, W, Cin, Cout, k, stride), act, use_res, out_f32
TC_CASES = [
    ("rows_sw128_exact", (384, 1, 1, 64, 64, 1, 1), lib.ACT_NONE, False, False),
    ("rows_ow_ragged_f32", (37, 1, 1, 128, 48, 1, 1), lib.ACT_NONE, False, True),
    ("rows_qkv", (4352, 1, 1, 640, 1920, 1, 1), lib.ACT_NONE, False, False),
    ("rows_sw64_k32", (300, 1, 1, 32, 128, 1, 1), lib.ACT_NONE, True, True),
    ("rows_sw64_k96", (300, 1, 1, 96, 128, 1, 1), lib.ACT_GELU, False, False),
    ("rows_sw32_k48", (260, 1, 1, 48, 32, 1, 1), lib.ACT_NONE, True, True),
    ("rows_sw32_k16", (129, 1, 1, 16, 16, 1, 1), lib.ACT_RELU, False, False),
    ("rows_fc2_res_f32", (1000, 1, 1, 1280, 640, 1, 1), lib.ACT_NONE, True, True),
    ("rows_fc1_gelu", (1000, 1, 1, 640, 1280, 1, 1), lib.ACT_GELU, False, False),
    ("rows_fuse1x1", (2, 8, 8, 256, 32, 1, 1), lib.ACT_NONE, False, False),
    ("rows_resnet_expand", (2, 8, 6, 512, 2048, 1, 1), lib.ACT_RELU, True, False),
    ("conv3_c32_small", (3, 16, 12, 32, 32, 3, 1), lib.ACT_RELU, True, False),
    ("conv3_c64_64x64", (2, 64, 64, 64, 64, 3, 1), lib.ACT_RELU, True, False),
    ("conv3_c32_many_tiles", (8, 64, 64, 32, 32, 3, 1), lib.ACT_RELU, False, False),
    ("conv3_c128_16x16", (4, 16, 16, 128, 128, 3, 1), lib.ACT_RELU, True, False),
    ("conv3_c256_8x8_n5", (5, 8, 8, 256, 256, 3, 1), lib.ACT_RELU, True, False),
    ("conv3_c48_9x7", (2, 9, 7, 48, 48, 3, 1), lib.ACT_RELU, True, False),
    ("conv3s2_64_128", (2, 16, 16, 64, 128, 3, 2), lib.ACT_NONE, False, False),
    ("conv3s2_256_64_64x64", (2, 64, 64, 256, 64, 3, 2), lib.ACT_RELU, False, False),
    ("conv3s2_48_96_odd", (1, 9, 7, 48, 96, 3, 2), lib.ACT_NONE, False, False),
    ("conv1s2_64_64", (2, 8, 8, 64, 64, 1, 2), lib.ACT_NONE, False, False),
    ("conv3_c96_48x36", (2, 48, 36, 96, 96, 3, 1), lib.ACT_RELU, True, False),
    ("conv3_c32_64x48", (2, 64, 48, 32, 32, 3, 1), lib.ACT_RELU, True, False),
]


# 3x3 / stride 1 / pad 1 shapes the halo-band kernel (csrc/capf_tc_halo.cu) must take when asked to (variant 2)
HALO_CASES = [
    ("halo_c32_64x64", (3, 64, 64, 32, 32, 3, 1), lib.ACT_RELU, True, False),
    ("halo_c64_32x32", (5, 32, 32, 64, 64, 3, 1), lib.ACT_RELU, True, False),
    ("halo_c32_many_bands", (40, 64, 64, 32, 32, 3, 1), lib.ACT_RELU, False, False),
    ("halo_c64_many_bands", (150, 32, 32, 64, 64, 3, 1), lib.ACT_NONE, True, False),
    ("halo_c48_9x7", (3, 9, 7, 48, 48, 3, 1), lib.ACT_RELU, True, False),
    ("halo_c32_64x48", (2, 64, 48, 32, 32, 3, 1), lib.ACT_RELU, True, False),
    ("halo_c64_32x24", (2, 32, 24, 64, 64, 3, 1), lib.ACT_NONE, False, False),
    ("halo_c16_tiny", (1, 3, 5, 16, 16, 3, 1), lib.ACT_NONE, False, False),
    ("halo_c64_c32", (2, 16, 16, 64, 32, 3, 1), lib.ACT_NONE, True, False),
    ("halo_c32_c64_96x72", (1, 96, 72, 32, 64, 3, 1), lib.ACT_RELU, False, False),
]


# 3x3 / stride 1 / pad 1, C = Cout = 128: halo band + streamed weights (csrc/capf_tc_halo128.cu) -- taken by default (variant 0)
HALO128_CASES = [
    ("h128_16x16", (4, 16, 16, 128, 128, 3, 1), lib.ACT_RELU, True, False),
    ("h128_16x12_nores", (3, 16, 12, 128, 128, 3, 1), lib.ACT_NONE, False, False),
    ("h128_tiny_8x6", (2, 8, 6, 128, 128, 3, 1), lib.ACT_RELU, True, False),
    ("h128_multiband_32x32", (2, 32, 32, 128, 128, 3, 1), lib.ACT_RELU, True, False),
    ("h128_many_bands", (330, 16, 16, 128, 128, 3, 1), lib.ACT_RELU, True, False),
    ("h128_wide_5x40", (1, 5, 40, 128, 128, 3, 1), lib.ACT_NONE, True, False),
    ("h128_ragged_bands_37x9", (2, 37, 9, 128, 128, 3, 1), lib.ACT_RELU, False, False),
]


# 3x3 / stride 1 / pad 1, C = 256 -> Cout = 32 (HRNet transition1.0): halo tile in four planes + streamed weights
# (csrc/capf_tc_halo256.cu) -- taken by default (variant 0)
HALO256_CASES = [
    ("h256_64x64", (3, 64, 64, 256, 32, 3, 1), lib.ACT_RELU, False, False),
    ("h256_64x48_none", (2, 64, 48, 256, 32, 3, 1), lib.ACT_NONE, False, False),
    ("h256_32x24", (3, 32, 24, 256, 32, 3, 1), lib.ACT_RELU, False, False),
    ("h256_tiny_5x3", (2, 5, 3, 256, 32, 3, 1), lib.ACT_RELU, False, False),
    ("h256_many_tiles", (40, 64, 64, 256, 32, 3, 1), lib.ACT_RELU, False, False),
    ("h256_ragged_37x45", (2, 37, 45, 256, 32, 3, 1), lib.ACT_NONE, False, False),
    ("h256_wide_6x130", (1, 6, 130, 256, 32, 3, 1), lib.ACT_RELU, False, False),
]


# 3x3 / stride 1 / pad 1 over WHOLE small images (H * W divides 128) on CTA pairs (csrc/capf_tc2.cu, conv mode): the 256-channel
# branch of HRNet at 8 x 8 takes it by default, the other shapes are forced (i[17] = 2)
TC2CONV_CASES = [
    ("t2c_c256_8x8_auto", (64, 8, 8, 256, 256, 3, 1), lib.ACT_RELU, True, False),
    ("t2c_c256_8x8_n5_ragged", (5, 8, 8, 256, 256, 3, 1), lib.ACT_RELU, True, False),
    ("t2c_c256_8x8_nores", (9, 8, 8, 256, 256, 3, 1), lib.ACT_RELU, False, False),
    ("t2c_c64_4x4", (19, 4, 4, 64, 64, 3, 1), lib.ACT_NONE, True, False),
    ("t2c_c128_8x4_n48", (7, 8, 4, 128, 48, 3, 1), lib.ACT_RELU, False, False),
    ("t2c_c64_2x2_wide", (300, 2, 2, 64, 512, 3, 1), lib.ACT_NONE, False, False),
    ("t2c_c192_8x8", (11, 8, 8, 192, 192, 3, 1), lib.ACT_RELU, True, False),
]


def run_tc_case(case, dt=torch.float16, impl=None, variant=0, epi=0, msub=0, bn=0, two=0, inplace=False):
    """Returns (rel_l2, max_abs, frac_bad_rows) of the tcgen05 kernel against the fp32 reference.
    inplace: the output buffer IS the residual buffer (how plan_memory runs a residual whose last reader is this op)."""
    name, (N, H, W, Cin, Cout, k, stride), act, use_res, out_f32 = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % (1 << 31))
    pad = k // 2
    x = torch.randn(N, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    odt = torch.float32 if out_f32 else dt
    res = torch.randn(N, Ho, Wo, Cout, generator=g) if use_res else None
    xq, wq = x.to(dt).float().to(DEV), w.to(dt).float().to(DEV)
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y = F.conv2d(xq.permute(0, 3, 1, 2), wq, bias.to(DEV), stride, pad).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    if act == lib.ACT_GELU:
        y = F.gelu(y)
    if use_res:
        y = y + res.to(odt).float().to(DEV)
    if act == lib.ACT_RELU:
        y = F.relu(y)
    impl = lib.IMPL_TCGEN05 if impl is None else impl
    if impl == lib.IMPL_TCGEN05:
        wp = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().to(dt).to(DEV)      # [Cout][(r,s,ci)]
    else:
        wp = w.permute(2, 3, 1, 0).reshape(-1, Cout).contiguous().to(dt).to(DEV)
    out = torch.full((N, Ho, Wo, Cout), float("nan"), dtype=odt, device=DEV)
    res_dev = res.to(odt).to(DEV) if use_res else None
    if inplace:
        assert use_res
        out = res_dev
    run_op(lib.OP_CONV2D, dt, odt, [N, H, W, Cin, Cout, k, k, stride, pad, Ho, Wo, act, impl, variant, epi, msub, bn, two], [],
           [x.to(dt).to(DEV), wp, bias.to(DEV), res_dev], [out])
    o = out.float()
    diff = (o - y)
    bad = ~torch.isfinite(o)
    diff = torch.where(bad, torch.full_like(diff, 1e3), diff)
    rel = float(diff.double().norm() / y.double().norm().clamp_min(1e-30))
    tol_row = 0.05 * float(y.abs().mean()) + 0.02
    bad_rows = float((diff.abs().reshape(-1, Cout).amax(dim=1) > tol_row).float().mean())
    return rel, float(diff.abs().max()), bad_rows


def run_split_case(case, msub=0):
    """bf16x3 split-operand GEMM (CAPF_OP_CAST split planes + CAPF_OP_CONV2D i[18] = 1, fp32 in / out) against the plain fp32
    PyTorch operator on the UN-rounded inputs.  Returns (rel_l2, max_abs)."""
    from capf_b200 import program
    name, (N, H, W, Cin, Cout, k, stride), act, use_res, _ = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % (1 << 31))
    pad = k // 2
    x = torch.randn(N, H, W, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = torch.randn(N, Ho, Wo, Cout, generator=g) if use_res else None
    y = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), stride, pad).permute(0, 2, 3, 1).float()
    if act == lib.ACT_GELU:
        y = F.gelu(y)
    if use_res:
        y = y + res
    if act == lib.ACT_RELU:
        y = F.relu(y)
    wp = program.split_weight_packer(lambda st: w.permute(0, 2, 3, 1))(None).to(DEV)
    assert tuple(wp.shape) == (Cout, 3 * k * k * Cin)
    xs = torch.empty(N, H, W, 2 * Cin, dtype=torch.bfloat16, device=DEV)
    n = x.numel()
    run_op(lib.OP_CAST, torch.float32, torch.bfloat16, [n & 0x7fffffff, n >> 31, Cin], [], [x.to(DEV)], [xs])
    out = torch.full((N, Ho, Wo, Cout), float("nan"), dtype=torch.float32, device=DEV)
    ints = [N, H, W, Cin, Cout, k, k, stride, pad, Ho, Wo, act, lib.IMPL_TCGEN05, 1, 0, msub, 0, 0, 1]
    run_op(lib.OP_CONV2D, torch.bfloat16, torch.float32, ints, [], [xs, wp, bias.to(DEV), res.to(DEV) if use_res else None], [out])
    d = out.cpu() - y
    d = torch.where(torch.isfinite(d), d, torch.full_like(d, 1e3))
    return float(d.double().norm() / y.double().norm()), float(d.abs().max())


def run_dual_case(M, C1, C2, Cout, act, dt=torch.float16, out_f32=False):
    """CAPF_OP_CONV2D with two A operands (i[19] = Cin2, in[5] = x2): out = act([x | x2] . w^T + b) against fp32 PyTorch on the
    16-bit-rounded inputs.  Returns rel_l2."""
    g = torch.Generator().manual_seed(M * 7 + C1 + 3 * C2 + Cout)
    x1, x2 = torch.randn(M, C1, generator=g), torch.randn(M, C2, generator=g)
    w = torch.randn(Cout, C1 + C2, generator=g) / (C1 + C2) ** 0.5
    b = torch.randn(Cout, generator=g)
    y = torch.cat([x1.to(dt).float(), x2.to(dt).float()], 1).double() @ w.to(dt).double().t() + b.double()
    y = y.float()
    if act == lib.ACT_RELU:
        y = F.relu(y)
    odt = torch.float32 if out_f32 else dt
    out = torch.full((M, Cout), float("nan"), dtype=odt, device=DEV)
    ints = [M, 1, 1, C1, Cout, 1, 1, 1, 0, 1, 1, act, lib.IMPL_TCGEN05, 0, 0, 0, 0, 0, 0, C2]
    run_op(lib.OP_CONV2D, dt, odt, ints, [], [x1.to(dt).to(DEV), w.to(dt).to(DEV), b.to(DEV), None, None, x2.to(dt).to(DEV)], [out])
    d = out.float().cpu() - y
    d = torch.where(torch.isfinite(d), d, torch.full_like(d, 1e3))
    return float(d.double().norm() / y.double().norm())
