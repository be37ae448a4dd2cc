"""PoseFormer lifter parameters -- API mirror of the reference's mvn/models/pose_dformer.py:PoseTransformer.

Holds the 191 ``volume_net.*`` tensors with the reference's names, shapes and default initialisation
(pose_dformer.py:145-208 incl. DeformableBlock._reset_parameters :103-113).  The forward pass
(sampler + 4 context + 4 res + 4 joint blocks + head, :210-241) is emitted by program.build_forward_program.
"""
import math

import torch
import torch.nn as nn


def _mlp(dim, hidden):
    m = nn.Module()
    m.fc1 = nn.Linear(dim, hidden)
    m.fc2 = nn.Linear(hidden, dim)
    return m


def _block(dim, mlp_ratio=2.0, eps=1e-6):
    b = nn.Module()
    b.norm1 = nn.LayerNorm(dim, eps=eps)
    b.attn = nn.Module()
    b.attn.qkv = nn.Linear(dim, dim * 3, bias=True)
    b.attn.proj = nn.Linear(dim, dim)
    b.norm2 = nn.LayerNorm(dim, eps=eps)
    b.mlp = _mlp(dim, int(dim * mlp_ratio))
    return b


def _context_block(feature_dims, dim, num_heads=4, num_samples=4, mlp_ratio=2):
    b = nn.Module()
    b.norm1 = nn.LayerNorm(dim)                       # eps 1e-5: norm_layer is not forwarded (:202)
    b.attention_weights = nn.Linear(dim, num_heads * num_samples)
    b.sampling_offsets = nn.Linear(dim, 2 * num_heads * num_samples)
    b.embed_proj = nn.ModuleList([nn.Linear(c, dim // num_heads) for c in feature_dims])
    b.norm2 = nn.LayerNorm(dim)
    b.mlp = _mlp(dim, int(dim * mlp_ratio))
    with torch.no_grad():                             # _reset_parameters (:103-113)
        b.sampling_offsets.weight.zero_()
        th = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
        d = torch.stack([th.cos(), th.sin()], -1)
        d = 0.01 * (d / d.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 2).repeat(1, num_samples, 1)
        d = d * torch.arange(1, num_samples + 1, dtype=torch.float32).view(1, num_samples, 1)
        b.sampling_offsets.bias.copy_(d.reshape(-1))
        b.attention_weights.weight.zero_()
        b.attention_weights.bias.zero_()
    return b


class PoseTransformer(nn.Module):
    def __init__(self, config=None, backbone="hrnet_32", num_joints=17, in_chans=2, num_heads=8, mlp_ratio=2.0,
                 qkv_bias=True, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=None,
                 variant="h36m"):
        """variant="mpi": the MPI-INF-3DHP tree's PoseTransformer (ContextPose_mpi/model/pose_dformer.py:174-262): block
        depth from config.depth, no context_blocks in the module tree / state_dict."""
        super().__init__()
        from ... import arch
        if not qkv_bias or qk_scale is not None or num_heads != 8 or mlp_ratio != 2.0 or num_joints != 17:
            raise NotImplementedError("libcapf_b200 implements the reference's shipped hyper-parameters only")
        base_dim, D, levels = int(config["base_dim"]), int(config["embed_dim_ratio"]), int(config["levels"])
        depth = levels if variant == "h36m" else int(config["depth"])
        self.levels = levels
        self.embed_dim_ratio = D
        self.drop_path_rate = drop_path_rate          # identity in eval; stochastic depth is a training feature
        E = D * (levels + 1)
        dims = arch.feature_dims(backbone, base_dim)
        self.feature_dim_list = dims
        self.coord_embed = nn.Linear(in_chans, D)
        self.feat_embed = nn.ModuleList([nn.Linear(c, D) for c in dims])
        self.Spatial_pos_embed = nn.Parameter(torch.zeros(1, 1 + levels, num_joints, D))
        self.joint_blocks = nn.ModuleList([_block(E) for _ in range(depth)])
        self.res_blocks = nn.ModuleList([_block(D) for _ in range(depth)])
        if variant == "h36m":
            self.context_blocks = nn.ModuleList([_context_block(dims, D) for _ in range(levels)])
        self.head = nn.Sequential(nn.LayerNorm(E), nn.Linear(E, 3))

    def forward(self, keypoints_2d, ref, features_list):
        raise RuntimeError("PoseTransformer is executed as part of CA_PF.forward by libcapf_b200; "
                           "call the CA_PF module (there is no standalone PyTorch path)")
