"""Parameter-holder trees with the reference's ``state_dict`` layout.

The modules created here own parameters/buffers only; they are never *called* -- all arithmetic runs in
libcapf_b200.  Real ``nn.Conv2d`` / ``nn.BatchNorm2d`` instances are used as holders so that
``SyncBatchNorm.convert_sync_batchnorm`` (train.py:317-318), ``.to()``, ``requires_grad`` handling and
checkpoint (de)serialisation behave exactly as with the reference classes.
"""
import torch.nn as nn

from ... import arch


def put(root: nn.Module, dotted: str, module: nn.Module):
    parts = dotted.split(".")
    cur = root
    for p in parts[:-1]:
        nxt = cur._modules.get(p)
        if nxt is None:
            nxt = nn.Module()
            cur.add_module(p, nxt)
        cur = nxt
    cur.add_module(parts[-1], module)


class ModuleVisitor:
    """arch.walk_* visitor that materialises the conv/bn holders."""

    def __init__(self, root: nn.Module):
        self.root = root

    def conv(self, cname, bname, x, cout, k=1, stride=1, act=arch.NONE, residual=None):
        pad = k // 2
        put(self.root, cname, nn.Conv2d(x.C, cout, k, stride, pad, bias=False))
        put(self.root, bname, nn.BatchNorm2d(cout))
        Ho = (x.H + 2 * pad - k) // stride + 1
        Wo = (x.W + 2 * pad - k) // stride + 1
        return arch.T(Ho, Wo, cout)

    def fuse(self, terms, relu=True):
        t = next(t for t, s in terms if s == 0)
        return arch.T(t.H, t.W, t.C)

    def maxpool(self, x):
        return arch.T((x.H - 1) // 2 + 1, (x.W - 1) // 2 + 1, x.C)

    def bilinear(self, x, Ho, Wo):
        return arch.T(Ho, Wo, x.C)

    def dead_conv(self, name, cin, cout, k):
        put(self.root, name, nn.Conv2d(cin, cout, k, 1, k // 2, bias=False))

    def dead_bn(self, name, c):
        put(self.root, name, nn.BatchNorm2d(c))
