"""nvJPEG vs cv2.imread on the mini dataset: per-frame difference statistics (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from capf_b200.mvn.datasets.human36m import Human36MSingleViewDataset  # noqa: E402

MINI = os.path.join(ROOT, "tests", "golden", "h36m_mini")
out = open(os.path.join(ROOT, "gpurun_out", "nvjpeg_diag.txt"), "w")


def say(*a):
    print(*a)
    print(*a, file=out, flush=True)


ds = Human36MSingleViewDataset(os.path.join(MINI, "processed"), os.path.join(MINI, "labels.pkl"), image_shape=(48, 64))
idx = list(range(len(ds)))
frames, sizes = ds.decode_frames(idx, "cuda")
f = frames.cpu().numpy().astype(np.int32)
s = sizes.cpu().numpy()
for k in idx:
    want = ds.read_frame(k).astype(np.int32)
    got = f[k, :want.shape[0], :want.shape[1]]
    d = np.abs(got - want)
    pad_clean = (not f[k, want.shape[0]:].any()) and (not f[k, :, want.shape[1]:].any())
    say(k, "size", tuple(s[k]), want.shape[:2], "mean", round(float(d.mean()), 3), "per-channel", [round(float(d[..., c].mean()), 3) for c in range(3)],
        "max", int(d.max()), "frac<=6", round(float((d <= 6).mean()), 4), "frac<=16", round(float((d <= 16).mean()), 4), "pad_clean", pad_clean)
a, b = ds.batch(idx, "cuda", decode="nvjpeg"), ds.batch(idx, "cuda", decode="cv2")
dc = (a["images"].int() - b["images"].int()).abs().float()
say("crops: mean", round(float(dc.mean()), 3), "max", int(dc.max()), "shape", tuple(a["images"].shape))
