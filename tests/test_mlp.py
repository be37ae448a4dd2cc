"""CAPF_OP_MLP (csrc/capf_tc_mlp.cu): the Mlp of a 128-wide transformer block (pose_dformer.py:25-31; DeformableBlock :138-141, Block :78),
fc1 + GELU + fc2 + residual in one kernel.  The program peephole and its CPU interpreter semantics; on the GPU bit-identity with the
two CAPF_OP_CONV2D launches it replaces (same K order, same epilogue arithmetic), ragged row counts, both 16-bit types, and the whole
forward with and without the pass."""
import pytest
import torch

import capf_b200
import interp
import protocol
from capf_b200 import lib, program
from conftest import build_case_model, golden_cases, load_golden, rel_l2


def test_peephole_fuses_the_eight_128_wide_mlps(monkeypatch):
    case = next(c for c in golden_cases() if c["name"] == "hrnet32_b2_128x96")
    g = load_golden(case["name"])
    m, w, cfg = build_case_model(case["backbone"], case["weight_seed"])
    B, H, W = case["B"], case["H"], case["W"]
    images, kp2d, crop = protocol.make_inputs(B, H, W, case["input_seed"])
    shapes = {k: tuple(v.shape) for k, v in w.items()}
    outs, progs = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("CAPF_FUSE_MLP", flag)
        prog = program.build_forward_program("hrnet_32", m.backbone.cfg, m._pf_cfg, shapes, B, H, W, "fp16", use_tc=True)
        it = interp.Interp(prog, w)
        it.t(prog.inputs["images"]).copy_(images)
        it.t(prog.inputs["kp2d"]).copy_(kp2d.reshape(-1, 2))
        it.t(prog.inputs["ref"]).copy_(torch.from_numpy(g["crop_after"]).reshape(-1, 2))
        it.run()
        outs.append(it.t(prog.outputs["out"]).clone())
        progs.append(prog)
    mlps = [op for op in progs[0].ops if op.kind == lib.OP_MLP]
    assert [op.tag for op in mlps] == [f"volume_net.context_blocks.{i}.mlp" for i in range(4)] + [f"volume_net.res_blocks.{i}.mlp" for i in range(4)]
    assert len(progs[1].ops) - len(progs[0].ops) == 8 and progs[0].flops() == progs[1].flops()
    assert all(op.ins[3] is op.outs[0] or program._same_buf(op.ins[3], op.outs[0]) for op in mlps)      # the token stream is updated in place
    assert not any(op.kind == lib.OP_MLP for op in progs[1].ops)
    assert rel_l2(outs[0], outs[1]) < 1e-4      # torch-CPU sums the two forms in different orders (a hidden value may round differently); the GPU tests demand equality
    assert rel_l2(outs[0].view(B, 1, 17, 3), g["out"]) < 2.5e-3
    # joint blocks (640 -> 1280 -> 640) and the fp32 / bf16x3 programs keep their separate Linears
    for prec in ("fp32", "bf16x3"):
        prog = program.build_forward_program("hrnet_32", m.backbone.cfg, m._pf_cfg, shapes, B, H, W, prec, use_tc=True)
        assert not any(op.kind == lib.OP_MLP for op in prog.ops)


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [128, 1, 300, 4352 * 4, 21760, 148 * 128 + 77])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_fused_mlp_is_bit_identical_to_the_two_linears(rows, dt):
    from gpu_util import run_op
    g = torch.Generator().manual_seed(rows % 1000)
    t = torch.randn(rows, 128, generator=g).to(dt).cuda()
    x = torch.randn(rows, 128, generator=g).cuda()
    w1 = (torch.randn(256, 128, generator=g) * 128 ** -0.5).to(dt).cuda()
    b1 = torch.randn(256, generator=g).cuda()
    w2 = (torch.randn(128, 256, generator=g) * 256 ** -0.5).to(dt).cuda()
    b2 = torch.randn(128, generator=g).cuda()
    # two launches: hidden in dtype dt, residual stream fp32 updated in place
    h = torch.full((rows, 256), float("nan"), dtype=dt, device="cuda")
    ref = x.clone()
    run_op(lib.OP_CONV2D, dt, dt, [rows, 1, 1, 128, 256, 1, 1, 1, 0, 1, 1, lib.ACT_GELU, lib.IMPL_TCGEN05], [], [t, w1, b1, None], [h])
    run_op(lib.OP_CONV2D, dt, torch.float32, [rows, 1, 1, 256, 128, 1, 1, 1, 0, 1, 1, lib.ACT_NONE, lib.IMPL_TCGEN05], [], [h, w2, b2, ref], [ref])
    out = x.clone()
    run_op(lib.OP_MLP, dt, torch.float32, [rows, 128, 256, 128], [], [t, w1, b1, out, w2, b2], [out])
    assert torch.equal(out, ref), f"max diff {(out - ref).abs().max().item():.3e}"
    # not in place, and against plain fp32 PyTorch
    out2 = torch.full((rows, 128), float("nan"), device="cuda")
    run_op(lib.OP_MLP, dt, torch.float32, [rows, 128, 256, 128], [], [t, w1, b1, x, w2, b2], [out2])
    assert torch.equal(out2, ref)
    hh = torch.nn.functional.gelu(t.float() @ w1.float().t() + b1).to(dt).float()
    want = x + hh @ w2.float().t() + b2
    assert rel_l2(out.cpu(), want.cpu()) < 1e-4      # erff vs torch's GELU may round a hidden value of the 16-bit tile differently


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_forward_with_and_without_fused_mlp_is_identical(precision, monkeypatch):
    cfg = capf_b200.make_config("hrnet_32")
    B, H, W = 9, 128, 96
    images, kp2d, crop = protocol.make_inputs(B, H, W, 12)
    outs, launches = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("CAPF_FUSE_MLP", flag)
        model = capf_b200.CA_PF(cfg, precision=precision).eval()
        w = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 0)
        model.load_state_dict(w, strict=True)
        model = model.cuda()
        with torch.no_grad():
            outs.append(model(images.cuda(), kp2d.cuda(), crop.clone().cuda()).clone())
        plan = next(iter(model._plans.values()))[0]
        launches.append(plan.num_launches)
        if flag == "1":
            assert sum("tc_mlp128_kernel" in plan.op_kernel(k) for k in range(plan.num_launches)) == 8
    assert launches[1] - launches[0] == 8
    assert torch.equal(outs[0], outs[1])
