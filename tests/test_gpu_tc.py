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
