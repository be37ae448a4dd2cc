"""Experiment configuration -- API mirror of the reference's mvn/utils/cfg.py.

Same surface: a module-level ``config`` attribute-dict pre-filled with defaults, ``update_config(path)`` that
overlays a YAML file and *rejects unknown keys* (cfg.py:166-181), and ``update_dir``.  The reference's
``experiments/human36m/human36m.yaml`` loads unchanged.  No dependency on ``easydict``.
"""
import os

import yaml


class AttrDict(dict):
    """dict with attribute access, nested dicts converted recursively (the subset of EasyDict the path uses:
    ``cfg.model.backbone.type``, ``cfg['STAGE2']``, ``key in cfg``, ``cfg[k] = v``)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    @classmethod
    def _conv(cls, v):
        if isinstance(v, dict) and not isinstance(v, AttrDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._conv(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._conv(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __deepcopy__(self, memo):
        import copy
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _stage(modules, branches, channels):
    return dict(NUM_MODULES=modules, NUM_BRANCHES=branches, NUM_BLOCKS=[4] * branches, NUM_CHANNELS=channels,
                BLOCK="BASIC", FUSE_METHOD="SUM")


_H36M_EXTRA = "data/human36m/extra"
_PRETRAINED = "data/pretrained/human36m"

DEFAULTS = {
    "title": "human36m_vol_softmax_single", "kind": "human36m", "azureroot": "", "logdir": "logs",
    "batch_output": False, "vis_freq": 1000, "vis_n_elements": 10, "id": 600, "frame": 1,
    "model": {
        "image_shape": [192, 256], "init_weights": True, "checkpoint": None,
        "backbone": {
            "type": "hrnet_32", "num_final_layer_channel": 17, "num_joints": 17, "num_layers": 152,
            "init_weights": True, "fix_weights": False,
            "checkpoint": _PRETRAINED + "/pose_hrnet_w32_256x192.pth",
            # HRNet (pose_hrnet.py:330-370)
            "NUM_JOINTS": 17, "PRETRAINED_LAYERS": ["*"], "STEM_INPLANES": 64, "FINAL_CONV_KERNEL": 1,
            "STAGE2": _stage(1, 2, [32, 64]),
            "STAGE3": _stage(4, 3, [32, 64, 128]),
            "STAGE4": _stage(3, 4, [32, 64, 128, 256]),
            # legacy pose_resnet keys (unused by the lifting path, accepted for YAML compatibility)
            "NUM_LAYERS": 50, "DECONV_WITH_BIAS": False, "NUM_DECONV_LAYERS": 3,
            "NUM_DECONV_FILTERS": [256, 256, 256], "NUM_DECONV_KERNELS": [4, 4, 4],
        },
        "volume_net": {
            "volume_aggregation_method": "softmax", "use_gt_pelvis": False, "cuboid_size": 2500.0, "volume_size": 64,
            "volume_multiplier": 1.0, "volume_softmax": True, "use_feature_v2v": True, "att_channels": 51,
            "temperature": 1500,
        },
        "poseformer": {"base_dim": 32, "embed_dim_ratio": 128, "depth": 4, "levels": 4},
    },
    "loss": {
        "criterion": "MAE", "mse_smooth_threshold": 0, "grad_clip": 0, "scale_keypoints_3d": 0.1,
        "use_volumetric_ce_loss": True, "volumetric_ce_loss_weight": 0.01,
        "use_global_attention_loss": True, "global_attention_loss_weight": 1000000,
    },
    "dataset": {
        "kind": "human36m", "data_format": "", "transfer_cmu_to_human36m": False, "root": "../H36M-Toolbox/images/",
        "extra_root": _H36M_EXTRA,
        "train_labels_path": _H36M_EXTRA + "/human36m-multiview-labels-GTbboxes.npy",
        "val_labels_path": _H36M_EXTRA + "/human36m-multiview-labels-GTbboxes.npy",
        "train_dataset": "multiview_human36m", "val_dataset": "human36m",
    },
    "train": {
        "n_objects_per_epoch": 15000, "n_epochs": 9999, "n_iters_per_epoch": 5000, "batch_size": 3, "optimizer": "Adam",
        "backbone_lr": 0.0001, "backbone_lr_step": [1000], "backbone_lr_factor": 0.1, "process_features_lr": 0.001,
        "volume_net_lr": 0.001, "volume_net_lr_decay": 0.99, "volume_net_lr_step": [1000], "volume_net_lr_factor": 0.5,
        "with_damaged_actions": True, "undistort_images": True, "scale_bbox": 1.0, "ignore_cameras": [], "crop": True,
        "erase": False, "shuffle": True, "randomize_n_views": True, "min_n_views": 1, "max_n_views": 1, "num_workers": 8,
        "limb_length_path": _H36M_EXTRA + "/mean_and_std_limb_length.h5",
        "pred_results_path": _PRETRAINED + "/human36m_alg_10-04-2019/checkpoints/0060/results/train.pkl",
    },
    "val": {
        "flip_test": True, "batch_size": 6, "with_damaged_actions": True, "undistort_images": True, "scale_bbox": 1.0,
        "ignore_cameras": [], "crop": True, "erase": False, "shuffle": False, "randomize_n_views": True,
        "min_n_views": 1, "max_n_views": 1, "num_workers": 10, "retain_every_n_frames_in_test": 1,
        "limb_length_path": _H36M_EXTRA + "/mean_and_std_limb_length.h5",
        "pred_results_path": _PRETRAINED + "/human36m_alg_10-04-2019/checkpoints/0060/results/val.pkl",
    },
}

config = AttrDict(DEFAULTS)


def update_dict(overlay, target):
    for key, val in overlay.items():
        if key not in target:
            raise ValueError("{} not exist in cfg.py".format(key))
        if isinstance(val, dict):
            update_dict(val, target[key])
        else:
            target[key] = val


def update_config(path):
    with open(path) as fin:
        update_dict(AttrDict(yaml.safe_load(fin)), config)


def _reroot(node, root):
    for key, val in node.items():
        if isinstance(val, str) and val.startswith("data/"):
            node[key] = os.path.join(root, val)
        elif isinstance(val, dict):
            _reroot(val, root)


def update_dir(azureroot, logdir):
    config.azureroot = azureroot
    config.logdir = os.path.join(config.azureroot, logdir)
    ckpt = config.model.checkpoint
    if ckpt is not None and not ckpt.startswith("data/"):
        config.model.checkpoint = os.path.join(config.azureroot, ckpt)
    _reroot(config, config.azureroot)


def backbone_overrides(cfg, backbone):
    """The per-backbone edits train.py:265-277 applies after parsing ``--backbone``."""
    cfg.model.backbone.type = backbone
    if backbone == "hrnet_32":
        cfg.model.poseformer.base_dim = 32
    elif backbone == "hrnet_48":
        cfg.model.backbone.checkpoint = "data/pretrained/coco/pose_hrnet_w48_256x192.pth"
        cfg.model.backbone.STAGE2.NUM_CHANNELS = [48, 96]
        cfg.model.backbone.STAGE3.NUM_CHANNELS = [48, 96, 192]
        cfg.model.backbone.STAGE4.NUM_CHANNELS = [48, 96, 192, 384]
        cfg.model.poseformer.base_dim = 48
    elif backbone == "cpn":
        cfg.train.batch_size = 256
        cfg.model.backbone.checkpoint = "data/pretrained/coco/CPN50_256x192.pth.tar"
        cfg.model.poseformer.base_dim = 256
    else:
        raise ValueError(backbone)
    return cfg
