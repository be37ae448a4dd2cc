"""GPU: the tcgen05/TMA implicit-GEMM kernel (csrc/capf_tc.cu) through capf_op_run against plain PyTorch fp32."""
import pytest
import torch

from tc_cases import HALO128_CASES, HALO256_CASES, HALO_CASES, TC2CONV_CASES, TC_CASES, run_tc_case

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


@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
@pytest.mark.parametrize("msub", [1, 2], ids=["m128", "m256"])
def test_tc_gemm_tile_heights(case, msub):
    """Per-tap kernel forced (variant 1) with both tile heights (one / two 128-row sub-tiles sharing each B stage): same
    operator, same tolerance."""
    rel, max_abs, bad_rows = run_tc_case(case, torch.float16, variant=1, msub=msub)
    print(f"{case[0]} msub={msub}: rel-L2 {rel:.3e} max-abs {max_abs:.3e} bad-rows {bad_rows:.4f}")
    tol = 2e-5 if case[4] else 1.5e-3
    assert bad_rows == 0.0 and rel < tol


@pytest.mark.parametrize("case", HALO128_CASES, ids=[c[0] for c in HALO128_CASES])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_halo128_conv_matches_fp32_reference(case, dt):
    """C = Cout = 128 3x3 convolutions (HRNet branch 2; pose_hrnet.py:79-95): halo band in shared memory + weights streamed
    through a ring, every band shape (one / several sub-tiles, several bands per image, more bands than SMs, ragged last band),
    against plain fp32 PyTorch and against the per-tap kernel (same accumulation order -> at most one output rounding step)."""
    import ctypes
    from capf_b200 import lib
    rel, max_abs, bad_rows = run_tc_case(case, dt)
    print(f"{case[0]} {dt}: rel-L2 {rel:.3e} max-abs {max_abs:.3e} bad-rows {bad_rows:.4f}")
    assert bad_rows == 0.0 and rel < (1.5e-3 if dt == torch.float16 else 8e-3)
    rel1, _, _ = run_tc_case(case, dt, variant=1)
    assert abs(rel - rel1) < 0.05 * rel1 + 1e-6
    # the default dispatch really is the new kernel
    name, (N, H, W, Cin, Cout, k, stride), act, use_res, _ = case
    op = lib.CapfOp()
    op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F16
    x = torch.zeros(N, H, W, Cin, dtype=torch.float16, device="cuda")
    w = torch.zeros(Cout, 9 * Cin, dtype=torch.float16, device="cuda")
    y = torch.zeros(N, H, W, Cout, dtype=torch.float16, device="cuda")
    for n, v in enumerate([N, H, W, Cin, Cout, 3, 3, 1, 1, H, W, act, lib.IMPL_TCGEN05]):
        op.i[n] = v
    op.inp[0], op.inp[1], op.out[0] = x.data_ptr(), w.data_ptr(), y.data_ptr()
    h = ctypes.c_void_p()
    arr = (lib.CapfOp * 1)(op)
    lib.check(lib.load().capf_plan_create(arr, 1, 0, ctypes.byref(h)), "plan")
    buf = ctypes.create_string_buffer(160)
    lib.load().capf_plan_op_kernel(h, 0, buf, 160)
    lib.load().capf_plan_destroy(h)
    assert buf.value.decode().startswith("tc_conv3_halo128_kernel"), buf.value


def _dispatched_kernel(case, dt=None):
    import ctypes
    from capf_b200 import lib
    name, (N, H, W, Cin, Cout, k, stride), act, use_res, _ = case
    op = lib.CapfOp()
    op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F16
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    x = torch.zeros(N, H, W, Cin, dtype=torch.float16, device="cuda")
    w = torch.zeros(Cout, k * k * Cin, dtype=torch.float16, device="cuda")
    y = torch.zeros(N, Ho, Wo, Cout, dtype=torch.float16, device="cuda")
    for n, v in enumerate([N, H, W, Cin, Cout, k, k, stride, pad, Ho, Wo, act, lib.IMPL_TCGEN05]):
        op.i[n] = v
    op.inp[0], op.inp[1], op.out[0] = x.data_ptr(), w.data_ptr(), y.data_ptr()
    h = ctypes.c_void_p()
    arr = (lib.CapfOp * 1)(op)
    lib.check(lib.load().capf_plan_create(arr, 1, 0, ctypes.byref(h)), "plan")
    buf = ctypes.create_string_buffer(160)
    lib.load().capf_plan_op_kernel(h, 0, buf, 160)
    lib.load().capf_plan_destroy(h)
    return buf.value.decode()


@pytest.mark.parametrize("case", HALO256_CASES, ids=[c[0] for c in HALO256_CASES])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_halo256_conv_matches_fp32_reference(case, dt):
    """C = 256 -> 32 3x3 convolutions (HRNet transition1.0; pose_hrnet.py:372-411): halo tile in four 64-channel planes (plane-major
    K loop, per-plane barriers) + weights streamed through a ring -- full-width and split tiles, ragged last tile row / column, more
    tiles than SMs, both accumulator stages -- against plain fp32 PyTorch and against the per-tap kernel (a different summation
    order of the same fp32 accumulation: the errors must agree)."""
    rel, max_abs, bad_rows = run_tc_case(case, dt)
    print(f"{case[0]} {dt}: rel-L2 {rel:.3e} max-abs {max_abs:.3e} bad-rows {bad_rows:.4f}")
    assert bad_rows == 0.0 and rel < (1.5e-3 if dt == torch.float16 else 8e-3)
    rel1, _, _ = run_tc_case(case, dt, variant=1)
    assert abs(rel - rel1) < 0.05 * rel1 + 1e-6
    assert _dispatched_kernel(case).startswith("tc_conv3_halo256_kernel")


@pytest.mark.parametrize("case", TC2CONV_CASES, ids=[c[0] for c in TC2CONV_CASES])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_pair_gemm_conv_mode_matches_fp32_reference(case, dt):
    """3x3 convolutions over whole small images on CTA pairs (cta_group::2; HRNet's 256-channel branch at 8 x 8,
    pose_hrnet.py:79-95): per K step one 4-D TMA box per CTA (zero fill = padding), half of the weight tile per CTA.  Against plain
    fp32 PyTorch, and bit for bit against the per-tap kernel (same K order, same fp32 accumulation); the residual may be updated
    in place."""
    rel, max_abs, bad_rows = run_tc_case(case, dt, two=2)
    print(f"{case[0]} {dt}: rel-L2 {rel:.3e} max-abs {max_abs:.3e} bad-rows {bad_rows:.4f}")
    assert bad_rows == 0.0 and rel < (1.5e-3 if dt == torch.float16 else 8e-3)
    rel1, max1, _ = run_tc_case(case, dt, variant=1, two=1)
    assert rel == rel1 and max_abs == max1
    if case[3]:
        assert run_tc_case(case, dt, two=2, inplace=True)[0] == rel
    if "auto" in case[0]:
        assert _dispatched_kernel(case).startswith("tc_gemm2_kernel")


@pytest.mark.parametrize("M,C1,C2,Cout", [(1000, 64, 64, 256), (4096, 128, 256, 256), (300, 64, 256, 256), (129, 32, 16, 48), (5000, 512, 256, 1024)])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_two_operand_linear_matches_fp32_reference(M, C1, C2, Cout, dt):
    """A Bottleneck's conv3 and its downsample conv as ONE GEMM over the K-concatenated inputs (program.fuse_downsample;
    pose_hrnet.py:116-136, networks/refineNet.py:17-45): every chunk-width combination, ragged M, both 16-bit types."""
    from capf_b200 import lib
    from tc_cases import run_dual_case
    rel = run_dual_case(M, C1, C2, Cout, lib.ACT_RELU, dt)
    print(f"dual {M}x({C1}+{C2})->{Cout} {dt}: rel-L2 {rel:.3e}")
    assert rel < (1.5e-3 if dt == torch.float16 else 8e-3)
    assert run_dual_case(M, C1, C2, Cout, lib.ACT_NONE, dt, out_f32=True) < 2e-5


SPLIT_CASES = [c for c in TC_CASES if c[1][3] % 16 == 0]


@pytest.mark.parametrize("case", SPLIT_CASES, ids=[c[0] for c in SPLIT_CASES])
@pytest.mark.parametrize("msub", [0, 1, 2], ids=["auto", "m128", "m256"])
def test_split_operand_gemm_is_fp32_class(case, msub):
    """precision "bf16x3": x = hi + lo, w = Wh + Wl in bfloat16, hi*Wh + lo*Wh + hi*Wl on the tensor cores (two passes over
    the taps inside one launch) -- against the fp32 operator on un-rounded inputs: operand precision 2^-16, so ~1e-5."""
    from tc_cases import run_split_case
    rel, max_abs = run_split_case(case, msub)
    print(f"{case[0]} split msub={msub}: rel-L2 {rel:.3e} max-abs {max_abs:.3e}")
    assert rel < 3e-5


TWO_CTA_CASES = [
    ("rows2_qkv", (4352, 1, 1, 640, 1920, 1, 1), 0, False, False),
    ("rows2_fc2_res_f32", (1000, 1, 1, 1280, 640, 1, 1), 0, True, True),
    ("rows2_fc1_gelu", (1000, 1, 1, 640, 1280, 1, 1), 2, False, False),
    ("rows2_proj_res_f32_inplace_sized", (4352, 1, 1, 640, 640, 1, 1), 0, True, True),
    ("rows2_ragged_small", (37, 1, 1, 128, 48, 1, 1), 0, False, True),
    ("rows2_relu_res_f16", (300, 1, 1, 64, 256, 1, 1), 1, True, False),
    ("rows2_many_tiles", (40000, 1, 1, 64, 64, 1, 1), 0, False, False),
]


@pytest.mark.parametrize("case", TWO_CTA_CASES, ids=[c[0] for c in TWO_CTA_CASES])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_two_cta_gemm_matches_fp32_reference(case, dt):
    """The cta_group::2 GEMM (csrc/capf_tc2.cu: a CTA pair per 256 x BN tile, each SM holding half of B) forced on
    Linear shapes of the lifter and on edge shapes (ragged M, one pair, more tiles than pairs, every epilogue mode)."""
    rel, max_abs, bad_rows = run_tc_case(case, dt, two=2)
    print(f"{case[0]} {dt}: rel-L2 {rel:.3e} max-abs {max_abs:.3e} bad-rows {bad_rows:.4f}")
    tol = 2e-5 if case[4] else (1.5e-3 if dt == torch.float16 else 8e-3)
    assert bad_rows == 0.0 and rel < tol


def test_two_cta_gemm_at_benchmark_size_is_deterministic():
    """QKV-sized GEMM on CTA pairs (every pair busy for two tiles, both accumulator stages, remote accumulator release):
    repeated launches agree bit for bit and match the 1-CTA kernel to output rounding."""
    import ctypes
    from capf_b200 import lib
    M, K, N = 4352, 640, 1920
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    outs = []
    for two in (2, 2, 2, 1):
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        op = lib.CapfOp()
        op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F16
        for n, v in enumerate([M, 1, 1, K, N, 1, 1, 1, 0, 1, 1, lib.ACT_NONE, lib.IMPL_TCGEN05, 0, 0, 0, 0, two]):
            op.i[n] = v
        op.inp[0], op.inp[1], op.inp[2] = a.data_ptr(), w.data_ptr(), bias.data_ptr()
        op.out[0] = out.data_ptr()
        lib.check(lib.load().capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "linear")
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert (outs[0].float() - outs[3].float()).abs().max().item() <= 8e-3      # one fp16 rounding step at |y| <= 8


@pytest.mark.parametrize("bn", [16, 48, 80, 240], ids=lambda v: f"bn{v}")
def test_two_cta_gemm_column_tiles(bn):
    case = ("rows2_bn_sweep", (1100, 1, 1, 128, 240, 1, 1), 0, True, False)
    rel, _, bad = run_tc_case(case, torch.float16, two=2, bn=bn)
    assert bad == 0.0 and rel < 1.5e-3, (bn, rel)


@pytest.mark.parametrize("bn", [16, 48, 80, 240], ids=lambda v: f"bn{v}")
def test_tc_gemm_column_tiles(bn):
    """Forced column-tile widths (incl. BN = 16, where one epilogue warp of each quadrant has no columns, and the
    single-accumulator-stage 256 x 240 tile of the QKV GEMM)."""
    case = ("rows_bn_sweep", (1100, 1, 1, 128, 240, 1, 1), 0, True, False)
    for msub in (1, 2):
        rel, _, bad = run_tc_case(case, torch.float16, variant=1, msub=msub, bn=bn)
        assert bad == 0.0 and rel < 1.5e-3, (msub, bn, rel)


INPLACE_CASES = [   # (case, variant, two): every kernel a residual conv / Linear of the program can be dispatched to
    (("inplace_layer1_conv3", (6, 64, 64, 64, 256, 1, 1), 1, True, False), 0, 0),
    (("inplace_conv3_c128", (9, 16, 16, 128, 128, 3, 1), 1, True, False), 1, 0),
    (("inplace_conv3_c256", (7, 8, 8, 256, 256, 3, 1), 1, True, False), 1, 0),
    (("inplace_halo_c64", (150, 32, 32, 64, 64, 3, 1), 1, True, False), 2, 0),
    (("inplace_halo_c32", (40, 64, 64, 32, 32, 3, 1), 1, True, False), 2, 0),
    (("inplace_halo_c48_odd", (3, 9, 7, 48, 48, 3, 1), 1, True, False), 2, 0),
    (("inplace_two_cta_f32", (4352, 1, 1, 640, 640, 1, 1), 0, True, True), 0, 2),
]


@pytest.mark.parametrize("spec", INPLACE_CASES, ids=[c[0][0] for c in INPLACE_CASES])
def test_residual_updated_in_place(spec):
    """out aliases the residual (plan_memory hands a dying residual's slot to the op's output): every epilogue reads a
    residual element before the same thread group overwrites it, on all three kernels."""
    case, variant, two = spec
    for dt in (torch.float16, torch.bfloat16):
        rel, max_abs, bad_rows = run_tc_case(case, dt, variant=variant, two=two, inplace=True)
        tol = 2e-5 if case[4] else (1.5e-3 if dt == torch.float16 else 8e-3)
        assert bad_rows == 0.0 and rel < tol, (case[0], dt, rel, max_abs)


def test_tc_gemm_in_place_residual_at_benchmark_size():
    """x += Linear(t) in place on the fp32 token stream (pose_dformer.py:77-78 as the lifter runs it), every SM busy;
    repeated launches from the same input agree bit for bit and match the fp32 reference."""
    import ctypes
    from capf_b200 import lib
    M, K, N = 4352, 1280, 640
    g = torch.Generator(device="cuda").manual_seed(11)
    t = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    x0 = torch.randn(M, N, device="cuda", generator=g)
    want = x0 + t.float() @ w.float().t() + bias
    outs = []
    for _ in range(3):
        x = x0.clone()
        op = lib.CapfOp()
        op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F32
        for n, v in enumerate([M, 1, 1, K, N, 1, 1, 1, 0, 1, 1, lib.ACT_NONE, lib.IMPL_TCGEN05, 0]):
            op.i[n] = v
        op.inp[0], op.inp[1], op.inp[2], op.inp[3] = t.data_ptr(), w.data_ptr(), bias.data_ptr(), x.data_ptr()
        op.out[0] = x.data_ptr()
        lib.check(lib.load().capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "linear")
        torch.cuda.synchronize()
        outs.append(x)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    rel = float((outs[0] - want).norm() / want.norm())
    assert rel < 2e-5, rel


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
