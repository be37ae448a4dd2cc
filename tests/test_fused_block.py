"""Fused HRNet BasicBlocks (CAPF_OP_BASICBLOCK; csrc/capf_tc_block.cu: 32 channels, one CTA per band; csrc/capf_tc_block64.cu:
64 channels on CTA pairs): the program peephole, its CPU interpreter semantics, and on the GPU bit-identity with the two
halo-band convolutions a block replaces."""
import contextlib
import ctypes
import io
import os

import pytest
import torch
import torch.nn.functional as F

import capf_b200
import interp
import protocol
from capf_b200 import lib, program
from conftest import rel_l2


def _programs(B, H, W, var="CAPF_FUSE_BLOCKS"):
    cfg = capf_b200.make_config("hrnet_32")
    with contextlib.redirect_stdout(io.StringIO()):
        m = capf_b200.CA_PF(cfg, precision="fp16").eval()
    w = protocol.make_weights([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 0)
    shapes = {k: tuple(v.shape) for k, v in w.items()}
    progs = {}
    old = os.environ.get(var)
    try:
        for flag in ("0", "1"):
            os.environ[var] = flag
            progs[flag] = program.build_forward_program("hrnet_32", m.backbone.cfg, m._pf_cfg, shapes, B, H, W, "fp16", use_tc=True)
    finally:
        if old is None:
            os.environ.pop(var, None)
        else:
            os.environ[var] = old
    return progs, w


def test_peephole_fuses_exactly_the_32_and_64_channel_basic_blocks():
    """CPU: the BasicBlocks of branches 0 and 1 (stage2: 4 + 4, stage3: 16 + 16, stage4: 12 + 12) collapse into
    CAPF_OP_BASICBLOCK, nothing else changes, and the interpreter gives the same network output for both programs."""
    B, H, W = 1, 64, 64
    progs, w = _programs(B, H, W)
    plain, fused = progs["0"], progs["1"]
    blocks = [op for op in fused.ops if op.kind == lib.OP_BASICBLOCK]
    assert len(blocks) == 64 and len(plain.ops) - len(fused.ops) == 64
    assert sum(op.i[3] == 32 and "branches.0." in op.tag for op in blocks) == 32
    assert sum(op.i[3] == 64 and "branches.1." in op.tag for op in blocks) == 32
    only32 = _programs(B, H, W, "CAPF_FUSE_BLOCKS64")[0]["0"]
    assert sum(op.kind == lib.OP_BASICBLOCK for op in only32.ops) == 32
    assert abs(plain.flops() - fused.flops()) == 0
    images, kp2d, crop = protocol.make_inputs(B, H, W, 3)
    crop /= torch.tensor([96.0, 128.0])
    crop -= 1.0
    outs = []
    for prog in (plain, fused):
        it = interp.Interp(prog, w)
        it.t(prog.inputs["images"]).copy_(images)
        it.t(prog.inputs["kp2d"]).copy_(kp2d.reshape(-1, 2))
        it.t(prog.inputs["ref"]).copy_(crop.reshape(-1, 2))
        it.run()
        outs.append(it.t(prog.outputs["out"]).clone())
    assert rel_l2(outs[1], outs[0]) < 1e-6


def _run_convs(x, w1, b1, w2, b2, dt):
    """The two-kernel form through the halo-band kernel: u = relu(conv1(x) + b1); y = relu(conv2(u) + b2 + x)."""
    N, H, W, C = x.shape
    L = lib.load()
    st = torch.cuda.current_stream().cuda_stream
    u = torch.empty_like(x)
    y = torch.empty_like(x)
    for (src, w, b, res, dst) in ((x, w1, b1, None, u), (u, w2, b2, x, y)):
        op = lib.CapfOp()
        op.kind = lib.OP_CONV2D
        op.dtype_in = op.dtype_out = lib.F16 if dt == torch.float16 else lib.BF16
        for n, v in enumerate([N, H, W, C, C, 3, 3, 1, 1, H, W, lib.ACT_RELU, lib.IMPL_TCGEN05, 2]):
            op.i[n] = v
        op.inp[0], op.inp[1], op.inp[2] = src.data_ptr(), w.data_ptr(), b.data_ptr()
        op.inp[3] = res.data_ptr() if res is not None else None
        op.out[0] = dst.data_ptr()
        lib.check(L.capf_op_run(ctypes.byref(op), 0, st), "conv")
    torch.cuda.synchronize()
    return u, y


def _run_block(x, w1, b1, w2, b2, dt):
    N, H, W, C = x.shape
    y = torch.full_like(x, float("nan"))
    op = lib.CapfOp()
    op.kind = lib.OP_BASICBLOCK
    op.dtype_in = op.dtype_out = lib.F16 if dt == torch.float16 else lib.BF16
    for n, v in enumerate([N, H, W, C]):
        op.i[n] = v
    op.inp[0], op.inp[1], op.inp[2], op.inp[3], op.inp[4] = x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
    op.out[0] = y.data_ptr()
    lib.check(lib.load().capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "basicblock")
    torch.cuda.synchronize()
    return y


_SHAPES = [(32, s) for s in [(3, 64, 64), (2, 64, 48), (5, 13, 9), (1, 5, 127), (2, 96, 72), (40, 64, 64)]] + \
          [(64, s) for s in [(3, 32, 32), (2, 32, 24), (5, 13, 9), (1, 5, 61), (1, 3, 5), (3, 48, 36), (7, 31, 32), (80, 32, 32)]]


@pytest.mark.gpu
@pytest.mark.parametrize("C,shape", _SHAPES, ids=str)
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_fused_block_equals_two_halo_convs(C, shape, dt):
    """Same MMAs in the same order, same 16-bit rounding of the intermediate: the fused kernel must reproduce the
    two-kernel result bit for bit, and both match an fp32 reference (band tails, ragged last band, tiny and wide images,
    more bands than SMs; for the pair kernel also an odd number of bands, i.e. a pair whose second CTA has no band)."""
    N, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(H * 1000 + W)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).to(dt)
    w1 = (torch.randn(C, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).to(dt)
    w2 = (torch.randn(C, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).to(dt)
    b1 = torch.randn(C, device="cuda", generator=g)
    b2 = torch.randn(C, device="cuda", generator=g)
    u, want = _run_convs(x, w1, b1, w2, b2, dt)
    got = _run_block(x, w1, b1, w2, b2, dt)
    assert torch.isfinite(got.float()).all()
    assert torch.equal(got, want), float((got.float() - want.float()).abs().max())
    xf = x.float().permute(0, 3, 1, 2)
    k1 = w1.float().reshape(C, 3, 3, C).permute(0, 3, 1, 2)
    k2 = w2.float().reshape(C, 3, 3, C).permute(0, 3, 1, 2)
    uu = F.relu(F.conv2d(xf, k1, b1, 1, 1)).to(dt).float()
    ref = F.relu(F.conv2d(uu, k2, b2, 1, 1) + xf).permute(0, 2, 3, 1)
    assert rel_l2(got.float().cpu(), ref.cpu()) < (1.5e-3 if dt == torch.float16 else 8e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,C", [(64, 64, 32), (32, 32, 64)], ids=str)
def test_fused_block_at_benchmark_size_is_deterministic(H, W, C):
    N = 256
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).half()
    w1 = (torch.randn(C, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).half()
    w2 = (torch.randn(C, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).half()
    b1 = torch.randn(C, device="cuda", generator=g)
    b2 = torch.randn(C, device="cuda", generator=g)
    _, want = _run_convs(x, w1, b1, w2, b2, torch.float16)
    outs = [_run_block(x, w1, b1, w2, b2, torch.float16) for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]) and torch.equal(outs[0], want)


@pytest.mark.gpu
def test_fused_block_rejects_images_too_wide_for_shared_memory():
    """A 255-pixel-wide band (input + intermediate + double buffer) does not fit one SM: the op must fail loudly -- the
    host peephole only emits it for W <= 128 (program.fuse_basic_blocks) and keeps the two-kernel form otherwise."""
    x = torch.zeros(1, 5, 255, 32, device="cuda", dtype=torch.float16)
    w = torch.zeros(32, 288, device="cuda", dtype=torch.float16)
    b = torch.zeros(32, device="cuda")
    with pytest.raises(lib.CapfError, match="not supported"):
        _run_block(x, w, b, w, b, torch.float16)
