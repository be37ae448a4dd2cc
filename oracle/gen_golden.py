"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):   python oracle/gen_golden.py
For every case the real ``mvn.models.conpose.CA_PF`` is built, loaded with the seeded protocol weights
(oracle/protocol.py) and run on the seeded protocol inputs; outputs and hooked intermediates are stored.
The fixtures pin oracle/capf_oracle.py (CPU tests) and the CUDA path (GPU tests) to the reference itself.
Weights and images are *not* stored (they are regenerated from their seeds; a checksum guards the RNG).
"""
import json
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import protocol  # noqa: E402
import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, backbone, B, H, W, weight seed, input seed
CASES = [
    ("hrnet32_b2_128x96", "hrnet_32", 2, 128, 96, 0, 11),
    ("hrnet32_b4_256x256", "hrnet_32", 4, 256, 256, 0, 1234),     # BASELINE.json configs[0]
    ("hrnet32_b3_256x192", "hrnet_32", 3, 256, 192, 1, 12),       # H36M native crop, odd batch
    ("hrnet48_b1_384x288", "hrnet_48", 1, 384, 288, 0, 13),       # configs[2] geometry
    ("cpn_b2_256x256", "cpn", 2, 256, 256, 0, 14),                # configs[3] geometry
]
N_SAMPLES = 512


def sample_positions(numel, name):
    g = np.random.Generator(np.random.PCG64([0x5A3F, zlib.crc32(name.encode()), numel % 9973]))
    return g.integers(0, numel, size=min(N_SAMPLES, numel))


def run_case(name, backbone, B, H, W, wseed, iseed):
    torch.manual_seed(0)
    model = ref_import.build_reference_model(backbone)
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    weights = protocol.make_weights(spec, wseed)
    model.load_state_dict(weights, strict=True)
    images, kp2d, crop = protocol.make_inputs(B, H, W, iseed)
    crop_in = crop.clone()

    grabbed = {}
    vn = model.volume_net
    hooks = [
        model.backbone.register_forward_hook(lambda m, i, o: grabbed.__setitem__("features", [t.detach().clone() for t in o])),
        vn.context_blocks[-1].register_forward_hook(lambda m, i, o: grabbed.__setitem__("tokens_context", o.detach().clone())),
        vn.res_blocks[-1].register_forward_hook(lambda m, i, o: grabbed.__setitem__("tokens_res", o.detach().clone())),
        vn.joint_blocks[-1].register_forward_hook(lambda m, i, o: grabbed.__setitem__("tokens_joint", o.detach().clone())),
        vn.context_blocks[0].register_forward_pre_hook(lambda m, i: grabbed.__setitem__("tokens_embed", i[0].detach().clone())),
    ]
    with torch.no_grad():
        out = model(images, kp2d, crop)
    for h in hooks:
        h.remove()

    rec = {
        "kp2d": kp2d.numpy(), "crop_in": crop_in.numpy(), "crop_after": crop.numpy(), "out": out.numpy(),
        "images_checksum": np.array([images.double().sum().item(), images.double().abs().sum().item(),
                                     float(images.reshape(-1)[12345 % images.numel()])]),
        "tokens_embed": grabbed["tokens_embed"].numpy(), "tokens_context": grabbed["tokens_context"].numpy(),
        "tokens_res": grabbed["tokens_res"].numpy(), "tokens_joint": grabbed["tokens_joint"].numpy(),
    }
    for l, f in enumerate(grabbed["features"]):
        flat = f.reshape(-1)                                  # NCHW order
        idx = sample_positions(flat.numel(), f"{name}/{l}")
        rec[f"feat{l}_shape"] = np.array(f.shape)
        rec[f"feat{l}_idx"] = idx.astype(np.int64)
        rec[f"feat{l}_val"] = flat[idx].numpy()
        rec[f"feat{l}_stats"] = np.array([f.double().sum().item(), f.double().abs().sum().item(), f.abs().max().item()])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    meta = dict(name=name, backbone=backbone, B=B, H=H, W=W, weight_seed=wseed, input_seed=iseed,
                torch=torch.__version__, out_abs_max=float(out.abs().max()))
    print(meta)
    return meta, spec


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    metas, manifests = [], {}
    for c in CASES:
        meta, spec = run_case(*c)
        metas.append(meta)
        manifests[c[1]] = [[k, list(s)] for k, s in spec]
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump({"reference_commit": "31875c9", "cases": metas}, f, indent=1)
    with open(os.path.join(OUT, "state_dict_manifest.json"), "w") as f:
        json.dump(manifests, f)


if __name__ == "__main__":
    main()
