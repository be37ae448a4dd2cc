"""CPU study: where does the 16-bit error of the forward come from?  (test tooling; runs the TEST interpreter)

Runs the fp32 op program through tests/interp.py with selectable rounding points and prints the rel-L2 of the
[B,1,17,3] output against the committed reference fixture.

  python tools/precision_study.py [case] [dtype] [only: comma-separated substrings of the variant labels]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.nn.functional as F

import capf_b200
from capf_b200 import lib, program
import interp
import protocol
from conftest import build_case_model, golden_cases, load_golden, rel_l2


class Study(interp.Interp):
    """flags: set of strings
       bb_w / bb_a / bb_store   : backbone weights / conv operand A / stored activations rounded
       lf_w / lf_a / lf_store   : lifter Linear weights / operand A / 16-bit stored intermediates (qkv, attn, hidden, samples)
    """

    def __init__(self, prog, state, flags, dt):
        super().__init__(prog, state)
        self.flags = flags
        self.dt = dt
        self.nbb = prog.n_backbone_ops
        self.k = 0

    def r(self, x, flag):
        return x.to(self.dt).float() if flag in self.flags else x

    def run(self):
        for k, op in enumerate(self.prog.ops):
            self.k = k
            getattr(self, "_op%d" % op.kind)(op)
            pre = "bb" if k < self.nbb else "lf"
            # storage rounding of 16-bit tensors in the 16-bit program
            if pre == "bb" and "bb_store" in self.flags:
                for o in op.outs:
                    if o is not None:
                        t = self.t(o)
                        t.copy_(t.to(self.dt).float())
            if pre == "lf" and "lf_store" in self.flags:
                for o in op.outs:
                    if o is None or o.dtype != "f32":
                        continue
                    name = o.root.name
                    # X / Y token streams and ow stay fp32 in the 16-bit program
                    if name.startswith("X#") or name.startswith("Y#") or ".ow#" in name or name.startswith("out"):
                        continue
                    t = self.t(o)
                    t.copy_(t.to(self.dt).float())

    def _op1(self, op):
        N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, act, impl = op.i[:13]
        pre = "bb" if self.k < self.nbb else "lf"
        x = self.t(op.ins[0]).reshape(N, H, W, Cin).float().permute(0, 3, 1, 2)
        w = self.t(op.ins[1]).float()
        if Cin >= 16 or pre == "lf":            # the 3-channel stem runs hi/lo split (fp32 class)
            x = self.r(x, pre + "_a")
            w = self.r(w, pre + "_w")
        if impl == lib.IMPL_TCGEN05:
            w = w.reshape(Cout, KH, KW, Cin).permute(0, 3, 1, 2)
        else:
            w = w.reshape(KH, KW, Cin, Cout).permute(3, 2, 0, 1)
        y = F.conv2d(x.double(), w.contiguous().double(), None, stride, pad).float().permute(0, 2, 3, 1)
        if op.ins[2] is not None:
            y = y + self.t(op.ins[2])
        if act == lib.ACT_GELU:
            y = F.gelu(y)
        if op.ins[3] is not None:
            y = y + self.t(op.ins[3]).reshape(N, Ho, Wo, Cout).float()
        if act == lib.ACT_RELU:
            y = F.relu(y)
        out = self.t(op.outs[0])
        out.copy_(y.reshape(out.shape).to(out.dtype))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "hrnet32_b2_128x96"
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[sys.argv[2] if len(sys.argv) > 2 else "fp16"]
    if name == "bench":           # the frames bench.py's parity leg uses: first 4 of the 256-frame batch, seed 1234
        import capf_oracle
        case = {"backbone": "hrnet_32", "weight_seed": 0, "B": 4, "H": 256, "W": 256}
        m, w, cfg = build_case_model("hrnet_32", 0, "fp32")
        B, H, W = 4, 256, 256
        images, kp2d, crop = protocol.make_inputs(256, H, W, 1234)
        images, kp2d, crop = images[:4].contiguous(), kp2d[:4].contiguous(), crop[:4].contiguous()
        tr = {}
        want = capf_oracle.ca_pf_forward(w, "hrnet_32", cfg.model.backbone, images, kp2d, crop.clone(), trace=tr)
        g = {"out": want.numpy()}
        for l in range(4):
            g[f"feat{l}_idx"] = np.arange(tr["features"][l].numel())
            g[f"feat{l}_val"] = tr["features"][l].reshape(-1).numpy()
    else:
        case = next(c for c in golden_cases() if c["name"] == name)
        g = load_golden(name)
        m, w, cfg = build_case_model(case["backbone"], case["weight_seed"], "fp32")
        B, H, W = case["B"], case["H"], case["W"]
        images, kp2d, crop = protocol.make_inputs(B, H, W, case["input_seed"])
    shapes = {k: tuple(v.shape) for k, v in w.items()}
    variants = [
        ("none", set()),
        ("all (the 16-bit program)", {"bb_w", "bb_a", "bb_store", "lf_w", "lf_a", "lf_store"}),
        ("backbone only", {"bb_w", "bb_a", "bb_store"}),
        ("lifter only", {"lf_w", "lf_a", "lf_store"}),
        ("backbone weights only", {"bb_w"}),
        ("backbone activations only (store)", {"bb_a", "bb_store"}),
        ("backbone operands only, fp32 storage (hi+lo trunk)", {"bb_w", "bb_a"}),
        ("backbone operand A only, fp32 storage", {"bb_a"}),
        ("lifter weights only", {"lf_w"}),
        ("lifter operand A only", {"lf_a"}),
        ("lifter storage only", {"lf_store", "lf_a"}),
        ("all, but backbone hi+lo trunk", {"bb_w", "bb_a", "lf_w", "lf_a", "lf_store"}),
        ("all, but backbone weights exact", {"bb_a", "bb_store", "lf_w", "lf_a", "lf_store"}),
        ("all, but lifter exact weights", {"bb_w", "bb_a", "bb_store", "lf_a", "lf_store"}),
    ]
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    for label, flags in variants:
        if only is not None and not any(o in label for o in only):
            continue
        prog = program.build_forward_program(case["backbone"], cfg.model.backbone, m._pf_cfg, shapes, B, H, W, "fp32")
        it = Study(prog, w, flags, dt)
        it.t(prog.inputs["images"]).copy_(images)
        it.t(prog.inputs["kp2d"]).copy_(kp2d.reshape(-1, 2))
        c = crop.clone().reshape(-1, 2)
        c /= torch.tensor([96.0, 128.0])
        c -= 1.0
        it.t(prog.inputs["ref"]).copy_(c)
        it.run()
        out = it.t(prog.outputs["out"]).reshape(B, 1, 17, 3)
        fe = []
        for l, f in enumerate(prog.feature_maps):
            nchw = it.t(f).float().permute(0, 3, 1, 2).reshape(-1)
            fe.append(rel_l2(nchw[torch.from_numpy(g[f"feat{l}_idx"])], g[f"feat{l}_val"]))
        gt = torch.from_numpy(g["out"])
        pf = " ".join(f"{rel_l2(out[i], gt[i]):.1e}" for i in range(B))
        print(f"{label:55s} out rel-L2 {rel_l2(out, g['out']):.3e}   maps " + " ".join(f"{e:.1e}" for e in fe) + "  frames " + pf, flush=True)


if __name__ == "__main__":
    main()
