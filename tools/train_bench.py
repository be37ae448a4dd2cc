"""Training-step throughput: CA_PF.forward under autograd + MPJPE + backward + AdamW on one B200,
next to the same step in PyTorch-eager CUDA (the oracle's functions on the GPU = the reference's torch calls + autograd +
torch.optim.AdamW).  python tools/train_bench.py [--batch 256] [--precision fp16] [--steps 10]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch

import capf_b200
import capf_oracle
from capf_b200 import train


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backbone", default="hrnet_32")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--hw", type=int, nargs=2, default=[256, 256])
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-yardstick", action="store_true")
    ap.add_argument("--profile", action="store_true", help="per-kernel-kind device time of one step + host time of a step")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, (H, W) = a.batch, a.hw
    cfg = capf_b200.make_config(a.backbone)
    model = capf_b200.CA_PF(cfg, precision=a.precision)
    sd = capf_b200.synth.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 0)
    model.load_state_dict(sd)
    model = model.to(dev)
    model.train(); model.backbone.eval(); model.volume_net.train()
    images, kp2d, crop = capf_b200.synth.make_inputs(B, H, W, 1234)
    images, kp2d, crop = images.to(dev), kp2d.to(dev), crop.to(dev)
    gt = (torch.randn(B, 1, 17, 3, generator=torch.Generator().manual_seed(9)) * 0.3).to(dev)
    opt = train.FusedAdamW(model.volume_net.parameters(), lr=6.4e-4, weight_decay=0.1)

    def step():
        pred = model(images, kp2d, crop.clone())
        loss = torch.mean(torch.norm(pred - gt, dim=3))
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def timed(fn, n):
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            l = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, float(l)

    ms, loss = timed(step, a.steps)
    print(f"tensor-core GEMM plans cached: {len(train._TC_PLANS)} (USE_TC={train.USE_TC})", file=sys.stderr)
    if a.profile:
        import collections
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        train.PROFILE = []
        step()
        torch.cuda.synchronize()
        agg = collections.defaultdict(lambda: [0, 0.0])
        for kind, ints, e0, e1 in train.PROFILE:
            key = (kind,) + (ints[:3] if kind == capf_b200.lib.OP_GEMM_F32 else ())
            agg[key][0] += 1
            agg[key][1] += e0.elapsed_time(e1)
        train.PROFILE = None
        names = {v: k for k, v in vars(capf_b200.lib).items() if k.startswith("OP_")}
        print(f"host time to enqueue one step {t_host * 1e3:.1f} ms; step wall {t_all * 1e3:.1f} ms; launches {sum(v[0] for v in agg.values())}", file=sys.stderr)
        for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"  {names.get(key[0], key[0]):22s} {str(key[1:]):28s} n={n:4d} {t:8.3f} ms", file=sys.stderr)
    with torch.no_grad():
        def fwd_only():
            model.eval()
            o = model(images, kp2d, crop.clone())
            model.train(); model.backbone.eval(); model.volume_net.train()
            return o.sum()
        ms_f, _ = timed(fwd_only, a.steps)
    out = {"workload": f"{a.backbone}, bs={B}, {H}x{W}, backbone {a.precision} (frozen), lifter fp32 training step (forward + backward + fused AdamW)",
           "train_step_ms": ms, "train_frames_per_s": B / ms * 1e3, "loss_after": loss, "inference_forward_ms_same_precision": ms_f}
    if not a.no_yardstick:
        sdd = {k: v.to(dev) for k, v in sd.items()}
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sdd.items() if k.startswith("volume_net.") and v.is_floating_point()}
        sd2 = dict(sdd); sd2.update(leaves)
        topt = torch.optim.AdamW([{"params": list(leaves.values()), "lr": 6.4e-4}], weight_decay=0.1)
        torch.backends.cudnn.benchmark = True
        for name, ctx in (("fp32_tf32_convs_default", None), ("autocast_f16", torch.float16)):
            def ystep():
                with torch.no_grad():
                    x = images.permute(0, 3, 1, 2)
                    ref = capf_oracle.normalize_crop_(crop.clone())
                    if ctx is None:
                        feats = capf_oracle.hrnet_forward(sdd, x.contiguous(), cfg.model.backbone)
                    else:
                        with torch.autocast("cuda", dtype=ctx):
                            feats = [f.float() for f in capf_oracle.hrnet_forward(sdd, x, cfg.model.backbone)]
                pred = capf_oracle.lifter_forward(sd2, kp2d, ref, feats)
                loss = torch.mean(torch.norm(pred - gt, dim=3))
                topt.zero_grad()
                loss.backward()
                topt.step()
                return loss
            yms, _ = timed(ystep, max(3, a.steps // 2))
            out[f"pytorch_eager_{name}_ms"] = yms
            out[f"pytorch_eager_{name}_frames_per_s"] = B / yms * 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
