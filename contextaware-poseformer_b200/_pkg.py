"""Public surface of the package (imported by both ``contextaware-poseformer_b200`` and the ``capf_b200`` alias)."""
import copy as _copy

from . import lib, program, arch, synth, dist, frontend, mpi, train  # noqa: F401
from .mvn.models.conpose import CA_PF
from .mvn.utils import cfg as _cfg

__all__ = ["CA_PF", "make_config", "lib", "program", "arch", "synth", "dist", "frontend", "mpi", "train"]


def make_config(backbone: str = "hrnet_32", yaml_path: str = None):
    """Fresh config (defaults [+ YAML overlay] + the train.py:265-277 backbone overrides)."""
    c = _copy.deepcopy(_cfg.AttrDict(_cfg.DEFAULTS))
    if yaml_path:
        import yaml
        with open(yaml_path) as f:
            _cfg.update_dict(_cfg.AttrDict(yaml.safe_load(f)), c)
    else:   # the values experiments/human36m/human36m.yaml sets for the model section
        c.model.backbone.fix_weights = True
        c.model.init_weights = False
        c.model.checkpoint = ""
    return _cfg.backbone_overrides(c, backbone)
