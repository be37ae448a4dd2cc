"""Do independent tcgen05 convs overlap usefully when issued on different streams?  Times a set of branch-like
convs (a) back to back on one stream and (b) round-robin on separate streams (CUDA-graph captured both ways)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from capf_b200 import lib  # noqa: E402

dev = "cuda:0"
L = lib.load()
SHAPES = [(256, 64, 64, 32, 32), (256, 32, 32, 64, 64), (256, 16, 16, 128, 128), (256, 8, 8, 256, 256)]
REPS = 8      # convs per branch (4 BasicBlocks)


def make(N, H, W, C, Co):
    x = torch.randn(N, H, W, C, device=dev).half()
    y = torch.empty(N, H, W, Co, device=dev, dtype=torch.float16)
    w = (torch.randn(Co, 9 * C, device=dev) / (9 * C) ** 0.5).half()
    b = torch.randn(Co, device=dev)
    ops = (lib.CapfOp * 2)()
    for k, (src, dst) in enumerate(((x, y), (y, x))):
        op = ops[k]
        op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F16
        for n, v in enumerate([N, H, W, C, Co, 3, 3, 1, 1, H, W, lib.ACT_RELU, lib.IMPL_TCGEN05, 0]):
            op.i[n] = v
        op.inp[0], op.inp[1], op.inp[2] = src.data_ptr(), w.data_ptr(), b.data_ptr()
        op.inp[3] = dst.data_ptr() if k == 1 else None          # second conv of the pair has a residual
        op.out[0] = dst.data_ptr()
    h = ctypes.c_void_p()
    lib.check(L.capf_plan_create(ops, 2, 0, ctypes.byref(h)), "plan")
    return h, (x, y, w, b, ops)


plans = [make(*s) for s in SHAPES]


def run_serial(stream):
    for h, _ in plans:
        for r in range(REPS // 2):
            lib.check(L.capf_plan_run(h, 0, 2, stream.cuda_stream), "run")


def run_parallel(main, side):
    for s in side:
        s.wait_stream(main)
    for (h, _), s in zip(plans, side):
        for r in range(REPS // 2):
            lib.check(L.capf_plan_run(h, 0, 2, s.cuda_stream), "run")
    for s in side:
        main.wait_stream(s)


def timed(fn):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn_w = lambda: fn(st)
        fn_w()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            fn_w()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)


side = [torch.cuda.Stream() for _ in plans]
a = timed(lambda st: run_serial(st))
b = timed(lambda st: run_parallel(st, side))
print(f"4 branches x {REPS} convs: one stream {a:.1f} us, four streams {b:.1f} us  ({a / b:.2f}x)")
