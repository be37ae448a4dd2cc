"""GPU diagnosis (test tooling): per-op deviation of the CUDA kernels from the TEST interpreter on the same 16-bit program.

For every op k of the program: copy the interpreter's input buffers into the plan's device buffers, launch op k alone,
run op k in the interpreter, compare every output buffer.  Both sides execute the same program with the same packed
weights and memory plan, so any op whose deviation is far above one output rounding step is a kernel that adds error
of its own.  Then: end-to-end error of the GPU path and of the interpreter against the fp32 oracle on the same frames.

  python tools/gpu_op_diag.py [--backbone hrnet_32] [--precision fp16] [--frames 4] [--hw 256 256] [--seed 1234]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch

import capf_b200
from capf_b200 import program
from capf_b200.program import Buf
import capf_oracle
import interp
import protocol


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="hrnet_32")
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--hw", type=int, nargs=2, default=[256, 256])
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--bench-batch", type=int, default=256, help="inputs are the first --frames of this many (bench.py's batch)")
    ap.add_argument("--thresh", type=float, default=2e-4)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    H, W = args.hw
    n = args.frames
    cfg = capf_b200.make_config(args.backbone)
    model = capf_b200.CA_PF(cfg, precision=args.precision).eval()
    w = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 0)
    model.load_state_dict(w, strict=True)
    model = model.to(dev)
    images, kp2d, crop = protocol.make_inputs(args.bench_batch, H, W, args.seed)
    images, kp2d, crop = images[:n].contiguous(), kp2d[:n].contiguous(), crop[:n].contiguous()

    c = crop.clone()
    tr = {}
    want = capf_oracle.ca_pf_forward(w, args.backbone, cfg.model.backbone, images, kp2d, c, trace=tr)
    with torch.no_grad():
        got = model(images.to(dev), kp2d.to(dev), crop.clone().to(dev)).cpu()
    plan = model.plan_for(n, H, W, dev)
    prog = plan.prog
    print(f"GPU {args.precision} vs fp32 oracle: rel-L2 {rel(got, want):.3e}; per frame " +
          " ".join(f"{rel(got[i], want[i]):.2e}" for i in range(n)))
    for l, f in enumerate(prog.feature_maps):
        print(f"  map{l}: GPU vs oracle {rel(plan.tensor(f).float().permute(0, 3, 1, 2).cpu(), tr['features'][l]):.3e}")

    it = interp.Interp(prog, w)
    it.t(prog.inputs["images"]).copy_(images)
    it.t(prog.inputs["kp2d"]).copy_(kp2d.reshape(-1, 2))
    it.t(prog.inputs["ref"]).copy_(c.reshape(-1, 2))
    worst = []
    for k, op in enumerate(prog.ops):
        for b in op.ins:
            if isinstance(b, Buf):
                plan.tensor(b).copy_(it.t(b))
        for b in op.outs:           # in-place / partially written outputs start from the same state
            if isinstance(b, Buf):
                plan.tensor(b).copy_(it.t(b))
        plan.run(k, 1)
        torch.cuda.synchronize()
        getattr(it, "_op%d" % op.kind)(op)
        for b in op.outs:
            if isinstance(b, Buf) and b.dtype != "i32":
                e = rel(plan.tensor(b).float().cpu(), it.t(b).float())
                worst.append((e, k, op.tag, plan.op_kernel(k), b.dtype))
    out_i = it.t(prog.outputs["out"]).reshape(n, 1, 17, 3)
    print(f"interpreter ({args.precision} program on the CPU) vs fp32 oracle: rel-L2 {rel(out_i, want):.3e}; per frame " +
          " ".join(f"{rel(out_i[i], want[i]):.2e}" for i in range(n)))
    print(f"ops with per-op deviation above {args.thresh} (inputs re-synchronised before every op):")
    for e, k, tag, kern, dt in worst:
        if e > args.thresh:
            print(f"  op {k:4d} {e:.3e} {dt} {tag} :: {kern}")
    worst.sort(reverse=True)
    print("largest 12:")
    for e, k, tag, kern, dt in worst[:12]:
        print(f"  op {k:4d} {e:.3e} {dt} {tag} :: {kern}")


if __name__ == "__main__":
    main()
