"""Time CAPF_OP_WARP_AFFINE_U8 at the Human3.6M geometry (1000x1000 frames -> 192x256 crops) against cv2.warpAffine on
the host (what the reference's DataLoader workers run per frame):  python tools/crop_bench.py [B]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from capf_b200.mvn.utils import img as host  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rng = np.random.default_rng(0)
    frames = torch.randint(0, 256, (B, 1000, 1000, 3), dtype=torch.uint8, device="cuda")
    centers = rng.uniform(350, 650, (B, 2))
    boxes = rng.uniform(400, 750, B)                                  # box height in pixels; scale = box / 200 (w = 3/4 h)
    trans = np.stack([host.get_affine_transform(c.astype(np.float32), np.array([b * 0.75 / 200, b / 200], np.float32), 0, (192, 256))
                      for c, b in zip(centers, boxes)])
    out_u8 = torch.empty(B, 256, 192, 3, dtype=torch.uint8, device="cuda")
    out_f = torch.empty(B, 256, 192, 3, dtype=torch.float32, device="cuda")
    src_bytes = float(np.sum(boxes * boxes * 0.75 * 3))               # bytes of the source boxes (each read once, ideally)
    minv = torch.from_numpy(np.stack([host.invert_affine(t) for t in trans])).cuda()
    for name, kw, out in (("uint8 crop", {}, out_u8), ("crop + normalise fp32", {"normalise": "hrnet_32"}, out_f),
                          ("uint8 crop, maps on the device (kernel only)", {"minv": minv}, out_u8),
                          ("crop + normalise fp32, maps on the device (kernel only)", {"minv": minv, "normalise": "hrnet_32"}, out_f)):
        for _ in range(3):
            host.crop_images(frames, trans, (192, 256), out=out, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            host.crop_images(frames, trans, (192, 256), out=out, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        alg = src_bytes + out.numel() * out.element_size()
        print(f"{name}: {ms * 1e3:.1f} us for {B} frames = {B / ms * 1e3:.0f} frames/s, {alg / ms / 1e6:.0f} GB/s algorithmic")
    try:
        import cv2
    except ImportError:
        print("cv2 not importable: no host timing")
        return
    cv2.setNumThreads(1)
    f = frames[:8].cpu().numpy()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 3.0:
        for k in range(8):
            ref = cv2.warpAffine(f[k], trans[k], (192, 256), flags=cv2.INTER_LINEAR)
            n += 1
    dt = time.perf_counter() - t0
    same = np.array_equal(ref, out_u8[7].cpu().numpy())
    print(f"cv2.warpAffine, 1 thread: {n / dt:.0f} frames/s; last crop identical to the GPU's: {same}")


if __name__ == "__main__":
    main()
