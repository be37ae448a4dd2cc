"""CPN-50 feature extractor -- API mirror of the reference's mvn/models/networks/network.py (CPN50 factory,
582-key ``state_dict`` with resnet / global_net / refine_net sub-trees, dead prediction heads included).
Topology: ``arch.walk_cpn``; arithmetic: libcapf_b200."""
import torch.nn as nn

from .... import arch
from .._tree import ModuleVisitor
from .._runtime import BackboneRuntimeMixin

__all__ = ["CPN50", "CPN"]


class CPN(BackboneRuntimeMixin, nn.Module):
    kind = "cpn"

    def __init__(self, output_shape=(64, 48), num_class=17):
        super().__init__()
        self.output_shape = tuple(output_shape)
        self.num_class = num_class
        arch.walk_cpn(ModuleVisitor(self), arch.T(256, 256, 3), self.output_shape, num_class)

    def walk(self, visitor, x):
        return arch.walk_cpn(visitor, x, self.output_shape, self.num_class)


def CPN50(out_size, num_class, pretrained=True):
    if pretrained:
        raise NotImplementedError("ImageNet download is not available; load a checkpoint with load_state_dict")
    return CPN(out_size, num_class)
