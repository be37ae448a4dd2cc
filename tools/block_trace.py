"""Per-band wait cycles of CTA 0's two MMA issuers in the fused BasicBlock kernel: python tools/block_trace.py N H W [C]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from capf_b200 import lib  # noqa: E402

N, H, W = [int(v) for v in sys.argv[1:4]]
C = int(sys.argv[4]) if len(sys.argv) > 4 else 32
x = torch.randn(N, H, W, C, device="cuda").half()
w1 = (torch.randn(C, 9 * C, device="cuda") / (9 * C) ** 0.5).half()
w2 = (torch.randn(C, 9 * C, device="cuda") / (9 * C) ** 0.5).half()
b1 = torch.randn(C, device="cuda")
b2 = torch.randn(C, device="cuda")
y = torch.empty_like(x)
trace = torch.zeros(32 * 32, dtype=torch.int64, device="cuda")
op = lib.CapfOp()
op.kind, op.dtype_in, op.dtype_out = lib.OP_BASICBLOCK, lib.F16, lib.F16
for n, v in enumerate([N, H, W, C]):
    op.i[n] = v
op.inp[0], op.inp[1], op.inp[2], op.inp[3], op.inp[4] = x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr()
op.inp[5] = trace.data_ptr()
op.out[0] = y.data_ptr()
for _ in range(3):
    trace.zero_()
    lib.check(lib.load().capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "block")
    torch.cuda.synchronize()
t = trace.cpu().view(32, 32)
for me in (0, 1):
    start = [int(v) for v in t[me * 8 + 0] if int(v)]
    end = [int(v) for v in t[me * 8 + 5] if int(v)]
    n = len(end)
    t0 = start[0]
    print(f"issuer {me}: bands {n}")
    print("  band start (rel):", [s - t0 for s in start[:n]])
    print("  band issue span :", [e - s for s, e in zip(start, end)])
    for idx, name in ((1, "wait X full"), (2, "wait tempty (phase A)"), (3, "wait MID ready"), (4, "wait tempty (phase B)"), (6, "issue (64-channel kernel)")):
        print(f"  {name:22s}:", [int(v) for v in t[me * 8 + idx][:n]])
if C == 64:      # epilogue timelines of CTA 0 (warp 4: epilogue 1, warp 12: epilogue 2), relative to the first conv1 band start
    t0 = int(t[0][0])
    rel = lambda row: [int(v) - t0 for v in t[row] if int(v)]
    print("E1 enter / tfull / tmem read / MID written / arrived:")
    for a, b, c, d, e in zip(rel(16), rel(17), rel(18), rel(19), rel(20)):
        print(f"   {a:7d} {b:7d} {c:7d} {d:7d} {e:7d}   (wait {b - a}, ld {c - b}, math+store {d - c}, fence+arrive {e - d})")
    print("E2 (group 2) enter / tfull / done:")
    for a, b, c in zip(rel(21), rel(22), rel(23)):
        print(f"   {a:7d} {b:7d} {c:7d}   (wait {b - a}, work {c - b})")
