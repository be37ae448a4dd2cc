"""Timeline of CTA 0 of the halo-band conv kernel (debug trace through op.in[4]): clock64 stamps, in cycles since
kernel start.  python tools/halo_trace.py N H W C Cout [res]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from capf_b200 import lib  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
N, H, W, C, Co = a[:5]
use_res = bool(a[5]) if len(a) > 5 else True
dev = "cuda:0"
x = torch.randn(N, H, W, C, device=dev).half()
w = (torch.randn(Co, 9 * C, device=dev) / (9 * C) ** 0.5).half()
bias = torch.randn(Co, device=dev)
res = torch.randn(N, H, W, Co, device=dev).half()
out = torch.empty(N, H, W, Co, device=dev, dtype=torch.float16)
trace = torch.zeros(16 * 256, dtype=torch.int64, device=dev)
op = lib.CapfOp()
op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F16
for n, v in enumerate([N, H, W, C, Co, 3, 3, 1, 1, H, W, lib.ACT_RELU, lib.IMPL_TCGEN05, 0]):
    op.i[n] = v
op.inp[0], op.inp[1], op.inp[2] = x.data_ptr(), w.data_ptr(), bias.data_ptr()
op.inp[3] = res.data_ptr() if use_res else None
op.inp[4] = trace.data_ptr()
op.out[0] = out.data_ptr()
L = lib.load()
for _ in range(3):
    trace.zero_()
    lib.check(L.capf_op_run(ctypes.byref(op), 0, torch.cuda.current_stream().cuda_stream), "conv")
    torch.cuda.synchronize()
t = trace.cpu().view(16, 256)
t0 = int(t[11, 0])
rel = lambda v: int(v) - t0 if int(v) else None
names = ["producer: TMA issue (band)", "issuer A: halo seen (band)", "issuer B: halo seen (band)", "issuer A: tile issued", "issuer B: tile issued",
         "epi g0: acc ready", "epi g1: acc ready", "epi g2: acc ready", "epi g0: tile done", "epi g1: tile done", "epi g2: tile done"]
for r, nm in enumerate(names):
    if r in (1, 2):
        vals = [int(v) for v in t[r][:8]]
        print(f"{nm.replace('halo seen', 'halo WAIT cycles'):32s}", vals)
        continue
    vals = [rel(v) for v in t[r] if int(v)]
    print(f"{nm:32s} n={len(vals):3d}", vals[:40])
for r, nm in ((12, "issuer A: tempty WAIT cycles"), (13, "issuer B: tempty WAIT cycles")):
    print(f"{nm:32s}", [int(v) for v in t[r][:32]])
import numpy as np
ta = np.array(sorted([rel(v) for r in (3, 4) for v in t[r] if int(v)]))
print("tile issue intervals (A+B merged):", np.diff(ta)[:60].tolist())
g = lambda r: [rel(v) for v in t[r] if int(v)]
top, rdy, ld, done = g(15), g(5), g(14), g(8)
print("epi g0 per tile: [loop top -> acc ready (wait)] [ready -> tmem loaded] [loaded -> done]")
print([(b - a, c - b, d - c) for a, b, c, d in zip(top, rdy, ld, done)][:24])
print("kernel end (last tile done):", max(rel(v) for r in (8, 9, 10) for v in t[r] if int(v)))
