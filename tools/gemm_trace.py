"""Timeline of CTA 0 of tc_gemm_kernel (debug trace through op.in[4]): clock64 stamps relative to kernel entry.
python tools/gemm_trace.py N H W Cin Cout k stride [res] [f32out] [msub] [bn]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from capf_b200 import lib  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
N, H, W, C, Co, k, stride = a[:7]
use_res = bool(a[7]) if len(a) > 7 else False
f32out = bool(a[8]) if len(a) > 8 else False
msub = a[9] if len(a) > 9 else 0
bn = a[10] if len(a) > 10 else 0
variant = a[11] if len(a) > 11 else 1     # i[13]: 1 = per-tap tc_gemm_kernel, 0 = automatic (2-CTA kernel for the wide Linears)
pad = k // 2
Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
dev = "cuda:0"
odt = torch.float32 if f32out else torch.float16
x = torch.randn(N, H, W, C, device=dev).half()
w = (torch.randn(Co, k * k * C, device=dev) / (k * k * C) ** 0.5).half()
bias = torch.randn(Co, device=dev)
res = torch.randn(N, Ho, Wo, Co, device=dev).to(odt)
out = torch.empty(N, Ho, Wo, Co, device=dev, dtype=odt)
trace = torch.zeros(256, dtype=torch.int64, device=dev)
op = lib.CapfOp()
op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.F16, lib.F32 if f32out else lib.F16
for n, v in enumerate([N, H, W, C, Co, k, k, stride, pad, Ho, Wo, lib.ACT_NONE, lib.IMPL_TCGEN05, variant, 0, msub, bn]):
    op.i[n] = v
op.inp[0], op.inp[1], op.inp[2] = x.data_ptr(), w.data_ptr(), bias.data_ptr()
op.inp[3] = res.data_ptr() if use_res else None
op.inp[4] = trace.data_ptr()
op.out[0] = out.data_ptr()
L = lib.load()
arr = (lib.CapfOp * 1)(op)
h = ctypes.c_void_p()
lib.check(L.capf_plan_create(arr, 1, 0, ctypes.byref(h)), "plan")
st = torch.cuda.current_stream().cuda_stream
ev = []
for _ in range(6):
    trace.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.check(L.capf_plan_run(h, 0, 1, st), "run")
    e1.record()
    torch.cuda.synchronize()
    ev.append(e0.elapsed_time(e1) * 1e3)
# back-to-back launches: amortised per-launch time without the event/launch gap
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    lib.check(L.capf_plan_run(h, 0, 1, st), "run")
e1.record()
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[0]
print(f"shape {a[:7]} res={use_res} f32out={f32out}: msub={t[6]} BN={t[7]} stages={t[8]} tiles={t[9]} grid={t[10]} nacc={t[11]}")
print(f"event-timed single launches (us): {[round(v, 1) for v in ev]};  50 back-to-back: {e0.elapsed_time(e1) * 1e3 / 50:.2f} us/launch")
print(f"CTA0: setup done +{t[2] - t0}  pdl_wait done +{t[3] - t0}  exit +{t[4] - t0} cycles;  globaltimer entry->exit {t[5] - t[1]} ns")
prod = [v - t0 for v in t[64:128] if v]
iss = [v - t0 for v in t[128:192] if v]
print("producer stage issue:", prod[:40])
print("issuer stage issued :", iss[:40])
epi = [(t[192 + 2 * i] - t0, t[193 + 2 * i] - t0) for i in range(16) if t[192 + 2 * i]]
print("epilogue warp 4 (acc ready, tile done):", epi)
fine = [v - t0 for v in t[224:256] if v]
print("epilogue warp 4, first tile, per 32-col group: [tmem_ld issued, ld waited, residual waited, finished->smem] ... + [slab stored]:")
print(fine)
