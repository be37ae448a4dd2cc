"""Host mirror of the evaluation reducer of the reference's dataset class
(``Human36MSingleViewDataset.evaluate_using_pred``, mvn/datasets/human36m.py:358-422).

The reference loops over the 30 action trials and calls MPJPE (torch), P_MPJPE (numpy: batched SVD on the host) and MPJVE
(numpy) on the masked frames of each.  Here one kernel (``CAPF_OP_POSE_ERRORS``) writes the three per-frame terms for all
frames with the predictions where ``CA_PF.forward`` left them (GPU), the rows are summed per action on the device, and only
a [30,3] table comes back to the host, which merges the two trials of an action exactly like the reference does.
There is no CPU path: CPU tensors are rejected.
"""
import ctypes

import numpy as np
import torch

from ... import lib

ACTIONS = ["Directions", "Discussion", "Eating", "Greeting", "Phoning", "Posing", "Purchases", "Sitting", "SittingDown", "Smoking",
           "TakingPhoto", "Waiting", "Walking", "WalkingDog", "WalkingTogether"]
retval = {"action_names": [f"{a}-{t}" for a in ACTIONS for t in (1, 2)]}      # human36m.py:18-33 (trial = subaction)


def previous_in_action(labels_action_idx: np.ndarray) -> np.ndarray:
    """prev[n] = the frame that precedes n among the frames of n's action (-1 for the first): the pairs np.diff forms on the
    action-masked sequences (loss.py:96-97 called from human36m.py:375)."""
    labels = np.asarray(labels_action_idx)
    prev = np.full(labels.shape[0], -1, dtype=np.int32)
    order = np.argsort(labels, kind="stable")
    same = labels[order][1:] == labels[order][:-1]
    prev[order[1:][same]] = order[:-1][same]
    return prev


def pose_errors(keypoints_3d_predicted: torch.Tensor, keypoints_gt: torch.Tensor, prev: torch.Tensor = None) -> torch.Tensor:
    """[N,(1,)J,3] fp32 on the GPU -> fp64 [N,3]: per-frame MPJPE, P-MPJPE and velocity error against frame prev[n]."""
    if not keypoints_3d_predicted.is_cuda or not keypoints_gt.is_cuda:
        raise lib.CapfError("pose_errors runs on a B200 through libcapf_b200; got a CPU tensor (no CPU path)")
    if keypoints_3d_predicted.shape != keypoints_gt.shape or keypoints_gt.shape[-1] != 3:
        raise ValueError("pose_errors: prediction and ground truth must both be [N,(1,)J,3]")
    n, j = keypoints_gt.shape[0], keypoints_gt.shape[-2]
    pred = keypoints_3d_predicted.reshape(n, j, 3).float().contiguous()
    gt = keypoints_gt.reshape(n, j, 3).float().contiguous()
    out = torch.empty(n, 3, dtype=torch.float64, device=gt.device)
    op = lib.CapfOp()
    op.kind, op.dtype_in, op.dtype_out = lib.OP_POSE_ERRORS, lib.F32, lib.F32
    op.i[0], op.i[1] = n, j
    op.inp[0], op.inp[1] = pred.data_ptr(), gt.data_ptr()
    if prev is not None:
        if prev.dtype != torch.int32 or tuple(prev.shape) != (n,) or not prev.is_cuda:
            raise ValueError("pose_errors: `prev` must be a CUDA int32 [N] tensor")
        prev = prev.contiguous()
        op.inp[2] = prev.data_ptr()
    op.out[0] = out.data_ptr()
    lib.check(lib.load().capf_op_run(ctypes.byref(op), gt.device.index or 0, torch.cuda.current_stream(gt.device).cuda_stream), "pose_errors")
    return out


def evaluate_using_pred(keypoints_gt: torch.Tensor, keypoints_3d_predicted: torch.Tensor, labels_action_idx) -> dict:
    """human36m.py:358-422.  ``labels_action_idx`` is the dataset attribute of that name (:529-530):
    (action - 2) * 2 + (subaction - 1) per frame.  Returns {action: {'MPJPE', 'P_MPJPE', 'MPJVE'}} in the inputs' unit."""
    labels = np.asarray(labels_action_idx).astype(np.int64)
    names = retval["action_names"]
    dev = keypoints_gt.device
    prev = torch.from_numpy(previous_in_action(labels)).to(dev)
    rows = pose_errors(keypoints_3d_predicted, keypoints_gt, prev)
    sums = torch.zeros(len(names), 3, dtype=torch.float64, device=dev).index_add_(0, torch.from_numpy(labels).to(dev), rows).cpu().numpy()
    count = np.bincount(labels, minlength=len(names)).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        # frame_count * mean(...) of the reference: MPJPE / P_MPJPE average F frames, MPJVE averages the F - 1 differences
        scores = {name: {"MPJPE": sums[k, 0], "P_MPJPE": sums[k, 1], "MPJVE": count[k] * (sums[k, 2] / (count[k] - 1) if count[k] != 1 else np.nan),
                         "frame_count": count[k]} for k, name in enumerate(names)}
        for base in [x[:-2] for x in names if x.endswith("-1")]:
            both = [scores.pop(f"{base}-{t}") for t in (1, 2)]
            scores[base] = {k: both[0][k] + both[1][k] for k in both[0]}
        return {k: {m: float(v[m] / v["frame_count"]) for m in ("MPJPE", "P_MPJPE", "MPJVE")} for k, v in scores.items()}
