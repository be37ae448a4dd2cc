"""GPU: the tcgen05/TMA implicit-GEMM kernel (csrc/capf_tc.cu) through capf_op_run against plain PyTorch fp32."""
import pytest
import torch

from tc_cases import TC_CASES, run_tc_case

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
