"""CAPF_OP_EXPAND_REDUCE (csrc/capf_tc_chain.cu): Bottleneck conv3 + residual + ReLU chained with conv1 + ReLU of the next block
(pose_hrnet.py:116-136).  The program peephole, its CPU interpreter semantics, and on the GPU bit-identity with the two
CAPF_OP_CONV2D launches it replaces."""
import contextlib
import ctypes
import io
import os

import pytest
import torch

import capf_b200
import interp
import protocol
from capf_b200 import lib, program
from conftest import rel_l2


def _programs(B, H, W):
    cfg = capf_b200.make_config("hrnet_32")
    with contextlib.redirect_stdout(io.StringIO()):
        m = capf_b200.CA_PF(cfg, precision="fp16").eval()
    w = protocol.make_weights([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 0)
    shapes = {k: tuple(v.shape) for k, v in w.items()}
    progs = {}
    old = os.environ.get("CAPF_FUSE_CHAIN")
    try:
        for flag in ("0", "1"):
            os.environ["CAPF_FUSE_CHAIN"] = flag
            progs[flag] = program.build_forward_program("hrnet_32", m.backbone.cfg, m._pf_cfg, shapes, B, H, W, "fp16", use_tc=True)
    finally:
        if old is None:
            os.environ.pop("CAPF_FUSE_CHAIN", None)
        else:
            os.environ["CAPF_FUSE_CHAIN"] = old
    return progs, w


def test_peephole_chains_the_plain_bottleneck_pairs_of_layer1():
    """CPU: layer1.1.conv3 + layer1.2.conv1 and layer1.2.conv3 + layer1.3.conv1 become one op each (layer1.0's conv3 carries the fused
    downsample operand and layer1.3's conv3 is followed by the transition convolutions: both stay); the interpreter gives the same
    network output for both programs."""
    B, H, W = 1, 64, 64
    progs, w = _programs(B, H, W)
    plain, fused = progs["0"], progs["1"]
    chained = [op for op in fused.ops if op.kind == lib.OP_EXPAND_REDUCE]
    assert [op.tag for op in chained] == ["backbone.layer1.1.conv3+2.conv1", "backbone.layer1.2.conv3+3.conv1"]
    assert len(plain.ops) - len(fused.ops) == 2 and plain.flops() == fused.flops()
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


def _conv1x1(x, w, b, res, dt):
    rows, cin = x.shape
    cout = w.shape[0]
    y = torch.empty(rows, cout, device="cuda", dtype=dt)
    op = lib.CapfOp()
    op.kind = lib.OP_CONV2D
    op.dtype_in = op.dtype_out = lib.F16 if dt == torch.float16 else lib.BF16
    for n, v in enumerate([rows, 1, 1, cin, cout, 1, 1, 1, 0, 1, 1, lib.ACT_RELU, lib.IMPL_TCGEN05]):
        op.i[n] = v
    op.inp[0], op.inp[1], op.inp[2] = x.data_ptr(), w.data_ptr(), b.data_ptr()
    op.inp[3] = res.data_ptr() if res is not None else None
    op.out[0] = y.data_ptr()
    lib.check(lib.load().capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "conv1x1")
    torch.cuda.synchronize()
    return y


def _chain(t, w3, b3, x, w1, b1, dt, in_place=False):
    rows = t.shape[0]
    y = x if in_place else torch.full((rows, 256), float("nan"), device="cuda", dtype=dt)
    u = torch.full((rows, 64), float("nan"), device="cuda", dtype=dt)
    op = lib.CapfOp()
    op.kind = lib.OP_EXPAND_REDUCE
    op.dtype_in = op.dtype_out = lib.F16 if dt == torch.float16 else lib.BF16
    for n, v in enumerate([rows, 64, 256, 64]):
        op.i[n] = v
    op.inp[0], op.inp[1], op.inp[2], op.inp[3], op.inp[4], op.inp[5] = (t.data_ptr(), w3.data_ptr(), b3.data_ptr(), x.data_ptr(), w1.data_ptr(),
                                                                         b1.data_ptr())
    op.out[0], op.out[1] = y.data_ptr(), u.data_ptr()
    lib.check(lib.load().capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "expand_reduce")
    torch.cuda.synchronize()
    return y, u


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [128, 117, 1173, 2 * 64 * 64, 19 * 1000 + 5, 60 * 64 * 64], ids=str)
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_expand_reduce_equals_two_convs(rows, dt):
    """Same MMAs in the same order and the same 16-bit rounding of y: bit for bit the two-launch result, for a single tile, a
    partial tile, a ragged last tile, and more tiles than SMs; also with y written over the residual (the memory planner's
    in-place update) and against an fp32 reference."""
    g = torch.Generator(device="cuda").manual_seed(rows)
    t = torch.randn(rows, 64, device="cuda", generator=g).to(dt)
    x = torch.randn(rows, 256, device="cuda", generator=g).to(dt)
    w3 = (torch.randn(256, 64, device="cuda", generator=g) / 8).to(dt)
    w1 = (torch.randn(64, 256, device="cuda", generator=g) / 16).to(dt)
    b3 = torch.randn(256, device="cuda", generator=g)
    b1 = torch.randn(64, device="cuda", generator=g)
    y_want = _conv1x1(t, w3, b3, x, dt)
    u_want = _conv1x1(y_want, w1, b1, None, dt)
    y, u = _chain(t, w3, b3, x, w1, b1, dt)
    assert torch.isfinite(y.float()).all() and torch.isfinite(u.float()).all()
    assert torch.equal(y, y_want), float((y.float() - y_want.float()).abs().max())
    assert torch.equal(u, u_want), float((u.float() - u_want.float()).abs().max())
    x2 = x.clone()
    y2, u2 = _chain(t, w3, b3, x2, w1, b1, dt, in_place=True)
    assert y2.data_ptr() == x2.data_ptr() and torch.equal(y2, y_want) and torch.equal(u2, u_want)
    yr = torch.relu(t.float() @ w3.float().t() + b3 + x.float())
    ur = torch.relu(yr.to(dt).float() @ w1.float().t() + b1)
    tol = 2e-3 if dt == torch.float16 else 1e-2
    assert rel_l2(y.float().cpu(), yr.cpu()) < tol and rel_l2(u.float().cpu(), ur.cpu()) < tol


@pytest.mark.gpu
def test_expand_reduce_rejects_other_shapes():
    t = torch.zeros(128, 32, device="cuda", dtype=torch.float16)
    op = lib.CapfOp()
    op.kind, op.dtype_in, op.dtype_out = lib.OP_EXPAND_REDUCE, lib.F16, lib.F16
    for n, v in enumerate([128, 32, 256, 64]):
        op.i[n] = v
    for n in range(6):
        op.inp[n] = t.data_ptr()
    op.out[0] = op.out[1] = t.data_ptr()
    with pytest.raises(lib.CapfError, match="not supported"):
        lib.check(lib.load().capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "expand_reduce")
