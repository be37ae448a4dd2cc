"""Throughput with 1 vs 2 forward pipelines in flight (two CA_PF instances / plans / CUDA graphs on two streams).
python tools/overlap_probe.py [steps]"""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import capf_b200  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
B, H, W = 256, 256, 256
cfg = capf_b200.make_config("hrnet_32")
models = []
for _ in range(2):
    with contextlib.redirect_stdout(io.StringIO()):
        m = capf_b200.CA_PF(cfg, precision="fp16", use_cuda_graph=True).eval()
    w = capf_b200.synth.make_weights([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 0)
    m.load_state_dict(w, strict=True)
    models.append(m.to(dev))
images, kp2d, crop = capf_b200.synth.make_inputs(B, H, W, 1234)
statics = [m.static_inputs(B, H, W, dev) for m in models]
for s in statics:
    s["images"].copy_(images.to(dev))
kp, cr = kp2d.to(dev), crop.to(dev)
crops = [cr.clone(), cr.clone()]
streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]


def run(n_pipes, n):
    main = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(main)
    for i in range(n):
        p = i % n_pipes
        with torch.cuda.stream(streams[p]):
            crops[p].copy_(cr)
            models[p](statics[p]["images"], kp, crops[p])
    for s in streams:
        main.wait_stream(s)


with torch.no_grad():
    for n_pipes in (1, 2, 1, 2):
        run(n_pipes, 6)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(n_pipes, steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"pipelines in flight {n_pipes}: {B * steps / ms * 1e3:.0f} frames/s ({ms / steps:.3f} ms/step)")
