"""TEST INFRASTRUCTURE ONLY -- fixtures for the MPI-INF-3DHP variant from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):   python oracle/gen_golden_mpi.py
Builds the real ``ContextPose_mpi/model/conpose.py:VolumetricTriangulationNet`` with the config of
``ContextPose_mpi/common/cfg.py`` + the ``run_3dhp.py:219-232`` backbone overrides, loads the seeded protocol weights and
runs the seeded protocol inputs on CPU; stores the output and the state_dict manifest (keys + shapes).
"""
import contextlib
import copy
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import protocol  # noqa: E402
import ref_import  # noqa: E402

REF_MPI = "/root/reference/ContextPose_mpi"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# name, backbone, B, H, W, weight seed, input seed
CASES = [
    ("mpi_hrnet32_b2_128x96", "hrnet_32", 2, 128, 96, 0, 21),
    ("mpi_hrnet48_b2_128x96", "hrnet_48", 2, 128, 96, 1, 22),   # B >= 2: the reference squeezes the batch dim away at B = 1 (pose_dformer.py:240)
]


def reference_model(backbone):
    ref_import._install_shims()
    layers = sys.modules["timm.models.layers"]          # the MPI tree also imports (and never uses) these two names
    if not hasattr(layers, "to_2tuple"):
        layers.to_2tuple = lambda v: (v, v)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
    import types
    for modname, attrs in (("timm.models.registry", {"register_model": lambda f: f}), ("timm.models.helpers", {"load_pretrained": None}),
                           ("timm.data", {"IMAGENET_DEFAULT_MEAN": None, "IMAGENET_DEFAULT_STD": None})):
        if modname not in sys.modules:
            mod = types.ModuleType(modname)
            for k, v in attrs.items():
                setattr(mod, k, v)
            sys.modules[modname] = mod
    sys.modules["timm.models"].__path__ = []          # let `import timm.models.x` resolve through sys.modules
    sys.modules["timm"].__path__ = []
    if REF_MPI not in sys.path:
        sys.path.insert(0, REF_MPI)
    from common.cfg import config as refcfg          # ContextPose_mpi/common/cfg.py
    c = copy.deepcopy(refcfg)
    if backbone == "hrnet_32":                       # run_3dhp.py:223-232
        c.model.backbone.STAGE2.NUM_CHANNELS = [32, 64]
        c.model.backbone.STAGE3.NUM_CHANNELS = [32, 64, 128]
        c.model.backbone.STAGE4.NUM_CHANNELS = [32, 64, 128, 256]
        c.model.poseformer.base_dim = 32
        c.model.poseformer.embed_dim_ratio = 64
    from model.conpose import VolumetricTriangulationNet
    with contextlib.redirect_stdout(io.StringIO()):
        m = VolumetricTriangulationNet(c, "cpu")
    return m.eval()


def main():
    manifest = {}
    for name, backbone, B, H, W, wseed, iseed in CASES:
        torch.manual_seed(0)
        model = reference_model(backbone)
        spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
        manifest[backbone] = [[k, list(s)] for k, s in spec]
        weights = protocol.make_weights(spec, wseed)
        model.load_state_dict(weights, strict=True)
        images, kp2d, crop = protocol.make_inputs(B, H, W, iseed)
        with torch.no_grad():
            out, second = model(images, kp2d, crop)
        assert second is None and tuple(out.shape) == (B, 3, 1, 17, 1)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), out=out.numpy(), crop_after=crop.numpy(),
                            meta=np.array(json.dumps({"backbone": backbone, "B": B, "H": H, "W": W, "wseed": wseed, "iseed": iseed})))
        print(name, tuple(out.shape), float(out.abs().mean()))
    with open(os.path.join(OUT, "state_dict_manifest_mpi.json"), "w") as f:
        json.dump(manifest, f)


if __name__ == "__main__":
    main()
