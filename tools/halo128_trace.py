"""Timeline of CTA 0 of tc_conv3_halo128_kernel (debug trace through op.in[4]).  python tools/halo128_trace.py [N H W]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from capf_b200 import lib
a = [int(v) for v in sys.argv[1:]] or [296, 16, 16]
N, H, W = a[:3]
C = 128
dev = "cuda:0"
x = torch.randn(N, H, W, C, device=dev).half()
w = (torch.randn(C, 9 * C, device=dev) / (9 * C) ** 0.5).half()
bias = torch.randn(C, device=dev)
res = torch.randn(N, H, W, C, device=dev).half()
out = torch.empty(N, H, W, C, device=dev, dtype=torch.float16)
trace = torch.zeros(256, dtype=torch.int64, device=dev)
op = lib.CapfOp()
op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F16
for n, v in enumerate([N, H, W, C, C, 3, 3, 1, 1, H, W, lib.ACT_RELU, lib.IMPL_TCGEN05, 0]):
    op.i[n] = v
op.inp[0], op.inp[1], op.inp[2], op.inp[3], op.inp[4] = x.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr(), trace.data_ptr()
op.out[0] = out.data_ptr()
L = lib.load()
arr = (lib.CapfOp * 1)(op)
h = ctypes.c_void_p()
lib.check(L.capf_plan_create(arr, 1, 0, ctypes.byref(h)), "plan")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    trace.zero_()
    lib.check(L.capf_plan_run(h, 0, 1, st), "run")
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    lib.check(L.capf_plan_run(h, 0, 1, st), "run")
e1.record()
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[0]
r = lambda v: v - t0 if v else None
buf = ctypes.create_string_buffer(160)
L.capf_plan_op_kernel(h, 0, buf, 160)
print(buf.value.decode(), f"  {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/launch back to back; CTA0 bands {t[2]}, exit +{r(t[1])}")
for b in range(min(8, t[2])):
    print(f"band {b}: issuer tempty ok +{r(t[16 + 4 * b])}  A landed +{r(t[17 + 4 * b])}  first chunk issued +{r(t[18 + 4 * b])}  last chunk issued +{r(t[19 + 4 * b])}"
          f" | epi warp4: acc ready +{r(t[48 + 4 * b])}  band done +{r(t[49 + 4 * b])}")
print("band 0 chunk issue times, issuer 0:", [r(v) for v in t[96:114]])
print("band 0 chunk issue times, issuer 1:", [r(v) for v in t[114:132]])
print("band 1 epilogue warp 4 per half [residual landed, epi done -> stores, stores done]:", [[r(t[80 + 4 * hh + k]) for k in range(3)] for hh in range(2)])
