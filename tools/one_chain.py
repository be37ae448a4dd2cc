"""Time CAPF_OP_EXPAND_REDUCE against the two 1x1 convolutions it replaces (CUDA events, plans): python tools/one_chain.py ROWS [reps]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from capf_b200 import lib  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 64 * 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = torch.Generator(device="cuda").manual_seed(1)
t = torch.randn(rows, 64, device="cuda", generator=g).half()
x = torch.randn(rows, 256, device="cuda", generator=g).half()
w3 = (torch.randn(256, 64, device="cuda", generator=g) / 8).half()
w1 = (torch.randn(64, 256, device="cuda", generator=g) / 16).half()
b3 = torch.randn(256, device="cuda", generator=g)
b1 = torch.randn(64, device="cuda", generator=g)
y = torch.empty_like(x)
u = torch.empty(rows, 64, device="cuda", dtype=torch.float16)
L = lib.load()
st = torch.cuda.current_stream().cuda_stream


def plan(ops):
    arr = (lib.CapfOp * len(ops))(*ops)
    h = ctypes.c_void_p()
    lib.check(L.capf_plan_create(arr, len(ops), 0, ctypes.byref(h)), "plan")
    return h, arr


def conv(src, w, b, res, dst, cin, cout):
    op = lib.CapfOp()
    op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F16
    for n, v in enumerate([rows, 1, 1, cin, cout, 1, 1, 1, 0, 1, 1, lib.ACT_RELU, lib.IMPL_TCGEN05]):
        op.i[n] = v
    op.inp[0], op.inp[1], op.inp[2] = src.data_ptr(), w.data_ptr(), b.data_ptr()
    op.inp[3] = res.data_ptr() if res is not None else None
    op.out[0] = dst.data_ptr()
    return op


ch = lib.CapfOp()
ch.kind, ch.dtype_in, ch.dtype_out = lib.OP_EXPAND_REDUCE, lib.F16, lib.F16
for n, v in enumerate([rows, 64, 256, 64]):
    ch.i[n] = v
ch.inp[0], ch.inp[1], ch.inp[2], ch.inp[3], ch.inp[4], ch.inp[5] = t.data_ptr(), w3.data_ptr(), b3.data_ptr(), x.data_ptr(), w1.data_ptr(), b1.data_ptr()
ch.out[0], ch.out[1] = y.data_ptr(), u.data_ptr()

for name, ops in (("two convolutions", [conv(t, w3, b3, x, y, 64, 256), conv(y, w1, b1, None, u, 256, 64)]), ("expand-reduce", [ch])):
    h, keep = plan(ops)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.check(L.capf_plan_run(h, 0, len(ops), st), "run")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    moved = rows * (64 + 256 + 256 + 64) * 2
    print(f"{name:18s} rows {rows}: best {min(ts):7.1f} us   ({moved / min(ts) / 1e6:.2f} TB/s of the fused kernel's algorithmic bytes)")
