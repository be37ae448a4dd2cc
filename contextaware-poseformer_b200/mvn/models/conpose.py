"""``CA_PF`` -- drop-in for the reference's mvn/models/conpose.py:CA_PF backed by libcapf_b200.

Same constructor (``CA_PF(config, device)``, conpose.py:11-27), attributes (``backbone``, ``volume_net``),
``state_dict`` layout and ``forward(images[B,H,W,3], keypoints_2d_cpn[B,17,2], keypoints_2d_cpn_crop[B,17,2])
-> [B,1,17,3]`` including the in-place normalisation of the caller's crop tensor (conpose.py:34-35).
In ``eval()`` / ``no_grad`` the output is detached.  With ``volume_net.train()`` under autograd the call is a training step
(SURVEY.md section 8 f2): frozen backbone through its plan, lifter forward/backward through ``capf_b200.train``.
"""
import torch
from torch import nn

from ... import lib, program
from . import pose_hrnet
from ._runtime import PlanCacheMixin, default_precision, default_use_tc, state_version
from .networks import network
from .pose_dformer import PoseTransformer

CPN_OUTPUT_SHAPE = (64, 48)     # cpn/test_config.py:24
CPN_NUM_CLASS = 17              # cpn/test_config.py:16


class CA_PF(PlanCacheMixin, nn.Module):
    _variant = "h36m"          # program variant; the MPI-INF-3DHP subclass (capf_b200.mpi) overrides it

    def __init__(self, config, device="cuda:0", precision=None, use_cuda_graph=False):
        super().__init__()
        bb = config.model.backbone
        self.num_joints = bb.num_joints
        self.backbone_type = bb.type
        if bb.type in ("hrnet_32", "hrnet_48"):
            self.backbone = pose_hrnet.get_pose_net(bb)
        elif bb.type == "cpn":
            self.backbone = network.CPN50(CPN_OUTPUT_SHAPE, CPN_NUM_CLASS, pretrained=False)
        else:
            raise ValueError(f"unknown backbone type {bb.type!r}")
        if bb.fix_weights:
            print("model backbone weights are fixed")
            for p in self.backbone.parameters():
                p.requires_grad = False
        self.volume_net = PoseTransformer(config.model.poseformer, backbone=bb.type, variant=self._variant)
        self._pf_cfg = {k: config.model.poseformer[k] for k in ("base_dim", "embed_dim_ratio", "levels", "depth")}
        self.precision = precision or default_precision()
        self.use_cuda_graph = use_cuda_graph
        self._plans = {}
        self._warned_train = False

    # -- plans -------------------------------------------------------------------------------------------
    def plan_for(self, B, H, W, device, debug_records=False):
        """Build (or fetch) the device plan for a batch geometry; repacks weights if parameters changed."""
        key = (B, H, W, self.precision, torch.device(device).index or 0, debug_records)
        ver = state_version(self)
        ent = self._plans.get(key)
        if ent is None:
            state = self.state_dict()
            shapes = {k: tuple(v.shape) for k, v in state.items()}
            prog = program.build_forward_program(self.backbone_type, getattr(self.backbone, "cfg", None), self._pf_cfg,
                                                 shapes, B, H, W, self.precision, use_tc=default_use_tc(),
                                                 debug_records=debug_records, variant=self._variant)
            ent = [program.Plan(prog, state, device), ver]
            self._plans[key] = ent
        elif ent[1] != ver:
            ent[0].repack(self.state_dict())
            ent[1] = ver
        return ent[0]

    def static_inputs(self, B, H, W, device):
        """The plan's own input buffers (images [B,H,W,3] f32, kp2d [B*17,2], ref [B*17,2]).  A loader may write
        the next batch's images straight into ``images`` (e.g. its H2D copy) and pass that tensor to forward(),
        which then skips its device-to-device staging copy."""
        plan = self.plan_for(B, H, W, device)
        return {k: plan.tensor(b) for k, b in plan.prog.inputs.items()}

    # -- forward -----------------------------------------------------------------------------------------
    def forward(self, images, keypoints_2d_cpn, keypoints_2d_cpn_crop):
        if not (images.is_cuda and keypoints_2d_cpn.is_cuda and keypoints_2d_cpn_crop.is_cuda):
            raise lib.CapfError("CA_PF runs on a B200 through libcapf_b200; got CPU tensors (there is no CPU path)")
        # train.py:145-148,186-201: volume_net.train() under autograd = a training step (DropPath live, differentiable output,
        # backbone frozen).  volume_net.train() under no_grad has no meaning here: eval semantics, warn once.
        train_step = self.volume_net.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.volume_net.parameters())
        if self.volume_net.training and not train_step and not self._warned_train:
            import warnings
            warnings.warn("CA_PF.forward: volume_net is in train() mode but autograd is off; running with eval semantics "
                          "(no DropPath, non-differentiable output)", stacklevel=2)
            self._warned_train = True
        B, H, W, C = images.shape
        J = self.num_joints
        if C != 3 or tuple(keypoints_2d_cpn.shape) != (B, J, 2):
            raise ValueError("expected images [B,H,W,3] and keypoints [B,17,2]")
        if tuple(keypoints_2d_cpn_crop.shape) != (B, J, 2):
            # conpose.py:34-35 normalises [..., :2]; anything but a [B,17,2] tensor would be mis-strided by the float2 kernel
            raise ValueError(f"keypoints_2d_cpn_crop must be [B,{J},2], got {tuple(keypoints_2d_cpn_crop.shape)}")
        dev = images.device
        if B == 0:          # an empty shard (ragged last batch of a rank): the reference's torch ops return an empty [0,1,17,3] as well
            return torch.zeros((0, 1, J, 3), dtype=torch.float32, device=dev)
        plan = None if train_step else self.plan_for(B, H, W, dev)
        stream = torch.cuda.current_stream(dev).cuda_stream

        # conpose.py:34-35 mutates the caller's tensor in place; keep that contract.
        crop = keypoints_2d_cpn_crop
        if crop.dtype == torch.float32 and crop.is_contiguous() and crop.data_ptr() % 8 == 0:
            lib.check(lib.load().capf_crop_normalize(crop.data_ptr(), B * self.num_joints, stream), "capf_crop_normalize")
            ref_src = crop
        else:
            tmp = crop.detach().to(torch.float32).contiguous()
            lib.check(lib.load().capf_crop_normalize(tmp.data_ptr(), B * self.num_joints, stream), "capf_crop_normalize")
            crop.copy_(tmp)
            ref_src = tmp
        if train_step:
            if self._variant != "h36m":
                raise NotImplementedError("the training step is implemented for the Human3.6M model (ContextPose/train.py)")
            if any(p.requires_grad for p in self.backbone.parameters()):
                raise NotImplementedError("only volume_net trains (config.model.backbone.fix_weights, conpose.py:22-25)")
            from ... import train
            return train.forward_train(self, images, keypoints_2d_cpn, ref_src)
        p = plan.prog
        static_images = plan.tensor(p.inputs["images"])
        if images.data_ptr() != static_images.data_ptr():      # callers may fill static_inputs() directly
            static_images.copy_(images)
        plan.tensor(p.inputs["kp2d"]).copy_(keypoints_2d_cpn.reshape(-1, 2))
        plan.tensor(p.inputs["ref"]).copy_(ref_src.reshape(-1, 2))
        if self.use_cuda_graph:
            plan.run_graph()
        else:
            plan.run()
        return plan.tensor(p.outputs["out"]).view(B, 1, self.num_joints, 3).clone()
