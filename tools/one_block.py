"""Run the fused BasicBlock op a few times (ncu captures / event timings): python tools/one_block.py N H W [reps] [C]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import ctypes  # noqa: E402

import torch  # noqa: E402

from capf_b200 import lib  # noqa: E402

N, H, W = [int(v) for v in sys.argv[1:4]]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
C = int(sys.argv[5]) if len(sys.argv) > 5 else 32
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(N, H, W, C, device="cuda", generator=g).half()
w1 = (torch.randn(C, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).half()
w2 = (torch.randn(C, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).half()
b1 = torch.randn(C, device="cuda", generator=g)
b2 = torch.randn(C, device="cuda", generator=g)
y = torch.empty_like(x)
op = lib.CapfOp()
op.kind, op.dtype_in, op.dtype_out = lib.OP_BASICBLOCK, lib.F16, lib.F16
for n, v in enumerate([N, H, W, C]):
    op.i[n] = v
op.inp[0], op.inp[1], op.inp[2], op.inp[3], op.inp[4] = x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
op.out[0] = y.data_ptr()
L = lib.load()
arr = (lib.CapfOp * 1)(op)
h = ctypes.c_void_p()
lib.check(L.capf_plan_create(arr, 1, 0, ctypes.byref(h)), "plan")
st = torch.cuda.current_stream().cuda_stream
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.check(L.capf_plan_run(h, 0, 1, st), "run")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
fl = 2 * 2.0 * N * H * W * C * 9 * C
print(f"fused block {N}x{H}x{W}x{C}: best {min(ts):.1f} us, {fl / min(ts) / 1e6:.1f} TFLOP/s")
