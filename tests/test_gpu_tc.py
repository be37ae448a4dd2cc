"""GPU: the tcgen05/TMA implicit-GEMM kernel (csrc/capf_tc.cu) through capf_op_run against plain PyTorch fp32."""
import pytest
import torch

from tc_cases import HALO_CASES, TC_CASES, run_tc_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_tc_conv_matches_fp32_reference(case, dt):
    rel, max_abs, bad_rows = run_tc_case(case, dt)
    print(f"{case[0]} {dt}: rel-L2 {rel:.3e} max-abs {max_abs:.3e} bad-rows {bad_rows:.4f}")
    out_f32 = case[4]
    # fp32 accumulation of exactly-representable 16-bit products: only the output rounding remains
    tol = 2e-5 if out_f32 else (1.5e-3 if dt == torch.float16 else 8e-3)
    assert bad_rows == 0.0 and rel < tol


@pytest.mark.parametrize("case", HALO_CASES, ids=[c[0] for c in HALO_CASES])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_halo_conv_matches_fp32_reference(case, dt):
    rel, max_abs, bad_rows = run_tc_case(case, dt, variant=2)
    print(f"{case[0]} {dt}: rel-L2 {rel:.3e} max-abs {max_abs:.3e} bad-rows {bad_rows:.4f}")
    tol = 2e-5 if case[4] else (1.5e-3 if dt == torch.float16 else 8e-3)
    assert bad_rows == 0.0 and rel < tol


def test_halo_and_per_tap_variants_agree_bitwise():
    """Both tcgen05 variants accumulate the same fp32 products; only the summation order inside the tensor pipe
    may differ, so their 16-bit outputs agree to one rounding step."""
    from tc_cases import DEV  # noqa: F401
    case = ("ab_c32", (4, 64, 64, 32, 32, 3, 1), 1, True, False)
    a = run_tc_case(case, torch.float16, variant=1)
    b = run_tc_case(case, torch.float16, variant=2)
    assert abs(a[0] - b[0]) < 5e-5 and a[2] == 0.0 and b[2] == 0.0


@pytest.mark.parametrize("shape", [(256, 64, 64, 32, 32), (256, 32, 32, 64, 64)], ids=["c32_bs256", "c64_bs256"])
def test_halo_conv_at_benchmark_size_is_stable(shape):
    """Benchmark-size run (every SM busy for many bands, both MMA issuer warps and all epilogue groups racing):
    repeated launches must agree bit for bit with each other and, to output rounding, with the per-tap variant."""
    import ctypes
    from capf_b200 import lib
    N, H, W, C, Co = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).half()
    w = (torch.randn(Co, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).half()
    bias = torch.randn(Co, device="cuda", generator=g)
    res = torch.randn(N, H, W, Co, device="cuda", generator=g).half()
    outs = []
    for variant in (2, 0, 0, 0, 1):
        out = torch.empty(N, H, W, Co, device="cuda", dtype=torch.float16)
        op = lib.CapfOp()
        op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F16
        for n, v in enumerate([N, H, W, C, Co, 3, 3, 1, 1, H, W, lib.ACT_RELU, lib.IMPL_TCGEN05, variant]):
            op.i[n] = v
        op.inp[0], op.inp[1], op.inp[2], op.inp[3] = x.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr()
        op.out[0] = out.data_ptr()
        lib.check(lib.load().capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "conv")
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[1], outs[2]) and torch.equal(outs[1], outs[3])
    for o in (outs[0], outs[4]):
        d = (o.float() - outs[1].float()).abs().max().item()
        assert d <= 4e-3, d      # one fp16 rounding step at |y| <= 8
