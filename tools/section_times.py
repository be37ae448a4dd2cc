"""In-graph device time of contiguous op ranges of the forward program (single stream, CUDA-graph replay of the range,
CUDA events, average of `reps` replays).  Shows what each section of the step costs WITHOUT the per-op host launch gaps
that Plan.time_ops() includes.

  python tools/section_times.py [--backbone hrnet_32] [--precision fp16] [--batch 256] [--hw 256 256]
"""
import argparse
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import capf_b200
from capf_b200 import lib


def section_of(tag):
    if tag.startswith("backbone."):
        p = tag.split(".")
        if p[1].startswith("stage"):
            return f"{p[1]}.{p[2]}"
        return p[1] if not p[1].startswith("layer1") else "layer1"
    if tag.startswith("volume_net."):
        return "lifter." + tag.split(".")[1]
    return "glue"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="hrnet_32")
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--hw", type=int, nargs=2, default=[256, 256])
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = capf_b200.make_config(a.backbone)
    model = capf_b200.CA_PF(cfg, precision=a.precision, use_cuda_graph=True).eval()
    w = capf_b200.synth.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 0)
    model.load_state_dict(w)
    model = model.to(dev)
    B, (H, W) = a.batch, a.hw
    images, kp2d, crop = capf_b200.synth.make_inputs(B, H, W, 1234)
    with torch.no_grad():
        model(images.to(dev), kp2d.to(dev), crop.to(dev))
    plan = model.plan_for(B, H, W, dev)
    ops = plan.prog.ops
    nb = plan.prog.n_backbone_ops

    def timed(fn):
        side = torch.cuda.Stream(dev)
        with torch.cuda.stream(side):
            fn(side)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn(side)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.reps

    def rng(first, count):
        return lambda st: plan.run(first, count, stream=st)

    full = timed(lambda st: plan.run(stream=st))
    full_1s = timed(lambda st: lib.check(plan._L.capf_plan_run(plan._h, 0, -1, st.cuda_stream)))
    print(f"full step (lanes): {full:.3f} ms   single stream: {full_1s:.3f} ms   ops {len(ops)} (backbone {nb})")
    print(f"backbone (single stream): {timed(rng(0, nb)):.3f} ms   lifter: {timed(rng(nb, len(ops) - nb)):.3f} ms")
    # contiguous runs of one section
    runs = []
    for k, op in enumerate(ops):
        s = section_of(op.tag) if op.tag else "glue"
        if op.kind == lib.OP_FUSE_SUM:
            s = runs[-1][0] if runs else s
        if runs and runs[-1][0] == s:
            runs[-1][2] += 1
        else:
            runs.append([s, k, 1])
    for s, k0, n in runs:
        ms = timed(rng(k0, n))
        fl = sum(o.flops for o in ops[k0:k0 + n])
        print(f"  {s:28s} ops [{k0:3d},{k0 + n:3d}) n={n:3d}  {ms * 1e3:8.1f} us   {fl / ms / 1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
