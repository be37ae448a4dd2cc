"""Where does the fused BasicBlock kernel differ from the two-conv form?  python tools/block_diag.py N H W"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from test_fused_block import _run_block, _run_convs  # noqa: E402

N, H, W = [int(v) for v in sys.argv[1:4]]
C = 32
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(N, H, W, C, device="cuda", generator=g).half()
w1 = (torch.randn(C, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).half()
w2 = (torch.randn(C, 9 * C, device="cuda", generator=g) / (9 * C) ** 0.5).half()
b1 = torch.randn(C, device="cuda", generator=g)
b2 = torch.randn(C, device="cuda", generator=g)
u, want = _run_convs(x, w1, b1, w2, b2, torch.float16)
got = _run_block(x, w1, b1, w2, b2, torch.float16)
bad = (got.float() - want.float()).abs().amax(dim=-1) > 0          # [N,H,W]
print("mismatching pixels:", int(bad.sum()), "of", bad.numel())
for n in range(min(N, 2)):
    rows = bad[n].sum(dim=1).tolist()
    print(f"image {n}: bad pixels per row:", [int(v) for v in rows])
    cols = bad[n].sum(dim=0).tolist()
    print(f"image {n}: bad pixels per col:", [int(v) for v in cols])
print("per image:", bad.reshape(N, -1).sum(dim=1).tolist())
