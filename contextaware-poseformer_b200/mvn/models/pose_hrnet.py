"""HRNet-W32/W48 feature extractor -- API mirror of the reference's mvn/models/pose_hrnet.py.

Same constructor contract (``get_pose_net(config.model.backbone)``, reading STAGE2..4 / PRETRAINED_LAYERS,
pose_hrnet.py:314-370, :536-539) and the same 1752-key ``state_dict``; the network itself is described once in
``arch.walk_hrnet`` and executed by libcapf_b200.  Calling the module directly runs the backbone-only program and
returns the four NCHW maps of pose_hrnet.py:501.
"""
import torch
import torch.nn as nn

from ... import arch
from ._tree import ModuleVisitor
from ._runtime import BackboneRuntimeMixin


class PoseHighResolutionNet(BackboneRuntimeMixin, nn.Module):
    kind = "hrnet"

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.cfg = cfg
        self.pretrained_layers = cfg["PRETRAINED_LAYERS"]
        arch.walk_hrnet(ModuleVisitor(self), arch.T(256, 256, 3), cfg)

    def walk(self, visitor, x):
        return arch.walk_hrnet(visitor, x, self.cfg)

    def init_weights(self, pretrained=""):
        """pose_hrnet.py:503-533: N(0, 1e-3) convs, unit BN, then an optional filtered checkpoint load."""
        import os
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.001)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if os.path.isfile(pretrained):
            sd = torch.load(pretrained, map_location="cpu")
            keep = {k: v for k, v in sd.items()
                    if self.pretrained_layers[0] == "*" or k.split(".")[0] in self.pretrained_layers}
            self.load_state_dict(keep, strict=False)
        elif pretrained:
            raise ValueError("{} is not exist!".format(pretrained))


def get_pose_net(config, is_train=False, **kwargs):
    return PoseHighResolutionNet(config, **kwargs)
