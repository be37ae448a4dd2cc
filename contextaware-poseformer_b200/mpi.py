"""MPI-INF-3DHP variant of the lifting path ("next" row f3).

API mirror of the reference's second tree, ``ContextPose_mpi/model/conpose.py:VolumetricTriangulationNet`` (:15-42) with its
``PoseTransformer`` (``ContextPose_mpi/model/pose_dformer.py:174-262``): HRNet backbone only, NO DeformableBlocks, block depth
from ``config.model.poseformer.depth``, ``embed_dim_ratio`` 96 (HRNet-48) / 64 (HRNet-32) (common/cfg.py:81-84,
run_3dhp.py:219-232), and the output layout ``(x.view(b,1,p,3,1).permute(0,3,1,2,4), None)`` (:260-261).  Same kernels and
C ABI as ``CA_PF``; the state_dict has the reference's keys (``backbone.*`` as HRNet, ``volume_net.*`` without
``context_blocks``), so ``run_3dhp.py``'s checkpoints (bare state_dict, ``module.`` prefix stripped, :252-255) load strictly.
"""
import copy

from .mvn.models.conpose import CA_PF
from .mvn.utils import cfg as _cfg


def make_mpi_config(backbone: str = "hrnet_48"):
    """Defaults of ContextPose_mpi/common/cfg.py + the run_3dhp.py:219-232 backbone overrides."""
    if backbone not in ("hrnet_32", "hrnet_48"):
        raise NotImplementedError("This backbone is not implemented yet.")      # run_3dhp.py:235
    c = copy.deepcopy(_cfg.AttrDict(_cfg.DEFAULTS))
    c.model.backbone.fix_weights = True
    c = _cfg.backbone_overrides(c, backbone)
    c.model.poseformer.embed_dim_ratio = 96 if backbone == "hrnet_48" else 64
    c.model.poseformer.depth = 4
    c.model.poseformer.levels = 4
    return c


class VolumetricTriangulationNet(CA_PF):
    _variant = "mpi"

    def forward(self, images, keypoints_2d_cpn, keypoints_2d_cpn_crop):
        y = super().forward(images, keypoints_2d_cpn, keypoints_2d_cpn_crop)          # [b,1,p,3]
        b, _, p, _ = y.shape
        return y.view(b, 1, p, 3, 1).permute(0, 3, 1, 2, 4).contiguous(), None
