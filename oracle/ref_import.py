"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference implementation.

Imports ``/root/reference/ContextPose/mvn`` in-process (it is pure Python/PyTorch) so that
  * ``oracle/gen_golden.py`` can produce the committed fixtures under ``tests/golden/`` and
  * the ``-m "not gpu"`` tests can pin ``oracle/capf_oracle.py`` (our restatement) to the reference.

The reference tree does not exist on the GPU box; everything here degrades to
``available() == False`` there.  Nothing in the product package may import this module.

Two third-party packages the reference needs are absent from this image, so tiny stand-ins are
registered in ``sys.modules`` first (behaviour per the pinned versions in
ContextPose/requirements.txt: easydict==1.10, timm==0.6.7):
  * ``easydict.EasyDict``  (mvn/utils/cfg.py:2)  -- recursive attribute dict
  * ``timm.models.layers.DropPath`` (mvn/models/pose_dformer.py:12) -- identity in eval mode
"""
import copy
import os
import sys
import types

REF_ROOT = os.environ.get("CAPF_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "ContextPose")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "mvn", "models", "conpose.py"))


def _install_shims():
    import torch
    import torch.nn as nn

    if "easydict" not in sys.modules:
        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                d = dict(d or {}, **kw)
                for k, v in d.items():
                    setattr(self, k, v)

            @classmethod
            def _wrap(cls, v):
                if isinstance(v, dict) and not isinstance(v, cls):
                    return cls(v)
                if isinstance(v, (list, tuple)):
                    return type(v)(cls._wrap(x) for x in v)
                return v

            def __setattr__(self, k, v):
                v = self._wrap(v)
                super().__setattr__(k, v)
                super().__setitem__(k, v)

            __setitem__ = __setattr__

        m = types.ModuleType("easydict")
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m

    if "timm" not in sys.modules:
        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1.0 - self.drop_prob
                mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
                return x * mask / keep

        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = DropPath
        timm.models = models
        models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


_CFG = None


def reference_config(backbone: str):
    """EasyDict config as ContextPose/train.py:254-277 would hand to CA_PF for ``--backbone``."""
    global _CFG
    _install_shims()
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)
    from mvn.utils import cfg as refcfg  # noqa
    if _CFG is None:
        refcfg.update_config(os.path.join(REF_PKG, "experiments", "human36m", "human36m.yaml"))
        _CFG = copy.deepcopy(refcfg.config)
    c = copy.deepcopy(_CFG)
    c.model.backbone.type = backbone
    if backbone == "hrnet_32":
        c.model.poseformer.base_dim = 32
    elif backbone == "hrnet_48":
        c.model.backbone.STAGE2.NUM_CHANNELS = [48, 96]
        c.model.backbone.STAGE3.NUM_CHANNELS = [48, 96, 192]
        c.model.backbone.STAGE4.NUM_CHANNELS = [48, 96, 192, 384]
        c.model.poseformer.base_dim = 48
    elif backbone == "cpn":
        c.model.poseformer.base_dim = 256
    else:
        raise ValueError(backbone)
    return c


def build_reference_model(backbone: str):
    """Unmodified reference ``CA_PF`` (mvn/models/conpose.py:10) on CPU, eval mode."""
    import contextlib
    import io
    c = reference_config(backbone)
    from mvn.models.conpose import CA_PF
    with contextlib.redirect_stdout(io.StringIO()):
        model = CA_PF(c, "cpu")
    return model.eval()
