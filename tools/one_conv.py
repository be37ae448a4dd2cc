"""Run ONE conv/linear op of the benchmark geometry through capf_op_run a few times (for ncu captures and quick
CUDA-event timings):  python tools/one_conv.py N H W Cin Cout k stride variant [reps] [res] [f32out]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ctypes  # noqa: E402

import torch  # noqa: E402

from capf_b200 import lib  # noqa: E402


def main():
    a = [int(v) for v in sys.argv[1:]]
    N, H, W, Cin, Cout, k, stride, variant = a[:8]
    reps = a[8] if len(a) > 8 else 5
    use_res = bool(a[9]) if len(a) > 9 else True
    f32out = bool(a[10]) if len(a) > 10 else False
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    dev = "cuda:0"
    dt = torch.float16
    odt = torch.float32 if f32out else dt
    x = torch.randn(N, H, W, Cin, device=dev).to(dt)
    w = (torch.randn(Cout, k * k * Cin, device=dev) / (Cin * k * k) ** 0.5).to(dt)
    bias = torch.randn(Cout, device=dev)
    res = torch.randn(N, Ho, Wo, Cout, device=dev).to(odt) if use_res else None
    out = torch.empty(N, Ho, Wo, Cout, device=dev, dtype=odt)
    op = lib.CapfOp()
    op.kind = lib.OP_CONV2D
    op.dtype_in = lib.F16
    op.dtype_out = lib.F32 if f32out else lib.F16
    for n, v in enumerate([N, H, W, Cin, Cout, k, k, stride, pad, Ho, Wo, lib.ACT_RELU, lib.IMPL_TCGEN05, variant]):
        op.i[n] = v
    op.i[15] = int(os.environ.get("ONE_MSUB", "0"))      # tile hints (0 = chooser)
    op.i[16] = int(os.environ.get("ONE_BN", "0"))
    op.inp[0], op.inp[1], op.inp[2] = x.data_ptr(), w.data_ptr(), bias.data_ptr()
    op.inp[3] = res.data_ptr() if use_res else None
    op.out[0] = out.data_ptr()
    trace = None
    if os.environ.get("ONE_TRACE"):         # kernels with a debug timeline (in[4]): print the raw counters of CTA 0
        trace = torch.zeros(256, dtype=torch.int64, device=dev)
        op.inp[4] = trace.data_ptr()
    L = lib.load()
    arr = (lib.CapfOp * 1)(op)
    h = ctypes.c_void_p()
    lib.check(L.capf_plan_create(arr, 1, 0, ctypes.byref(h)), "plan")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.check(L.capf_plan_run(h, 0, 1, st), "run")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    if trace is not None:
        print("trace:", trace[:16].tolist())
    fl = 2.0 * N * Ho * Wo * Cout * k * k * Cin
    best = min(ts)
    print(f"conv {a[:8]} res={use_res}: best {best:.1f} us  median {sorted(ts)[len(ts) // 2]:.1f} us  {fl / best / 1e6:.1f} TFLOP/s (L2 flushed)")


if __name__ == "__main__":
    main()
