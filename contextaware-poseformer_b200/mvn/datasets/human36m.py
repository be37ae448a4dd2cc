"""Host mirror of the evaluation reducer of the reference's dataset class
(``Human36MSingleViewDataset.evaluate_using_pred``, mvn/datasets/human36m.py:358-422).

The reference loops over the 30 action trials and calls MPJPE (torch), P_MPJPE (numpy: batched SVD on the host) and MPJVE
(numpy) on the masked frames of each.  Here one kernel (``CAPF_OP_POSE_ERRORS``) writes the three per-frame terms for all
frames with the predictions where ``CA_PF.forward`` left them (GPU), the rows are summed per action on the device, and only
a [30,3] table comes back to the host, which merges the two trials of an action exactly like the reference does.
There is no CPU path: CPU tensors are rejected.
"""
import ctypes
import os

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
        # (a trial without frames scores nan like the reference's 0 * mean(empty); with ONE frame the reference's .squeeze()
        # breaks, here MPJVE is nan and the two position errors are that frame's)
        scores = {name: {"MPJPE": sums[k, 0] if count[k] else np.nan, "P_MPJPE": sums[k, 1] if count[k] else np.nan,
                         "MPJVE": count[k] * (sums[k, 2] / (count[k] - 1)) if count[k] > 1 else np.nan,
                         "frame_count": count[k]} for k, name in enumerate(names)}
        for base in [x[:-2] for x in names if x.endswith("-1")]:
            both = [scores.pop(f"{base}-{t}") for t in (1, 2)]
            scores[base] = {k: both[0][k] + both[1][k] for k in both[0]}
        return {k: {m: float(v[m] / v["frame_count"]) for m in ("MPJPE", "P_MPJPE", "MPJVE")} for k, v in scores.items()}


class Human36MSingleViewDataset:
    """Batched, GPU-side counterpart of the reference class of the same name (human36m.py:482-584) over the SAME on-disk
    formats: the pickled list of label dicts (keys ``joints_3d, joints_2d_cpn, joints_2d_cpn_crop, center, scale, subject,
    action, subaction, camera_id, image_id, video_id``) and JPEG frames under
    ``root/s_SS_act_AA_subact_BB_ca_CC/s_SS_act_AA_subact_BB_ca_CC_NNNNNN.jpg``.

    Where the reference decodes AND crops one frame per ``__getitem__`` on a DataLoader worker, ``batch(indices, device)``
    decodes on the host (``cv2.imread`` -- JPEG decode is a library call in both), uploads the raw frames once and leaves
    the crop (``mvn.utils.img.crop_images`` = CAPF_OP_WARP_AFFINE_U8) and everything after it to the GPU.  ``rank`` /
    ``world_size`` slice the labels like ``prepare_labels`` (:536-552: ``n // world`` per rank, the remainder on the last);
    as in the reference ``labels_action_idx`` keeps covering the WHOLE table, because evaluation runs on the all-gathered
    predictions (train.py:216-235)."""

    def __init__(self, root, labels_path, image_shape=(192, 256), rank=None, world_size=None):
        import pickle
        self.root = root
        self.image_shape = tuple(image_shape)
        with open(labels_path, "rb") as f:
            self.labels = pickle.loads(f.read())
        self.labels_action_idx = (np.array([s["action"] for s in self.labels]) - 2) * 2 + (np.array([s["subaction"] for s in self.labels]) - 1)
        self.dist_size = self.prepare_labels(rank, world_size)
        self.video_idx = np.array([s["video_id"] for s in self.labels])

    def prepare_labels(self, rank, world_size):
        if rank is None or world_size is None:
            return None
        total = len(self.labels)
        n = total // world_size
        start = n * rank
        self.labels = self.labels[start:(total if rank == world_size - 1 else start + n)]
        return [n if i < world_size - 1 else total - n * (world_size - 1) for i in range(world_size)]

    def __len__(self):
        return len(self.labels)

    def image_path(self, idx) -> str:
        s = self.labels[idx]
        sub = "s_{:02d}_act_{:02d}_subact_{:02d}_ca_{:02d}".format(s["subject"], s["action"], s["subaction"], s["camera_id"] + 1)
        return os.path.join(self.root, sub, "{}_{:06d}.jpg".format(sub, s["image_id"]))

    def read_frame(self, idx) -> np.ndarray:
        import cv2
        frame = cv2.imread(self.image_path(idx), cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        if frame is None:
            raise FileNotFoundError(self.image_path(idx))
        return frame

    def read_frames(self, indices, workers: int = 8) -> list:
        """read_frame for many indices on a thread pool (cv2.imread releases the GIL), in the order asked."""
        if workers <= 1 or len(indices) <= 1:
            return [self.read_frame(i) for i in indices]
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(workers, len(indices))) as pool:
            return list(pool.map(self.read_frame, indices))

    def decode_frames(self, indices, device):
        """JPEG files -> (uint8 [B,Hs,Ws,3] BGR frames on `device`, int32 [B,2] live (h, w)) through nvJPEG: the compressed
        bytes cross PCIe and each frame is decoded straight into the padded storage the crop reads (capf_jpeg_decode_batch).
        nvJPEG and libjpeg (cv2.imread) are different decoders of the same streams: pixels agree to a few grey levels."""
        L = lib.load()
        if not L.capf_jpeg_available():
            raise lib.CapfError("decode_frames: libnvjpeg could not be loaded on this machine (use decode='cv2')")
        dev = torch.device(device)
        blobs = []
        for i in indices:
            with open(self.image_path(i), "rb") as f:
                blobs.append(f.read())
        n = len(blobs)
        sizes = np.zeros((n, 2), dtype=np.int32)
        h, w = ctypes.c_int(), ctypes.c_int()
        for k, b in enumerate(blobs):
            lib.check(L.capf_jpeg_info(b, len(b), dev.index or 0, ctypes.byref(h), ctypes.byref(w)), "capf_jpeg_info")
            sizes[k] = (h.value, w.value)
        hs, ws = int(sizes[:, 0].max()), int(sizes[:, 1].max())
        frames = torch.zeros(n, hs, ws, 3, dtype=torch.uint8, device=dev)
        data = (ctypes.c_char_p * n)(*blobs)
        lens = (ctypes.c_size_t * n)(*[len(b) for b in blobs])
        got = (ctypes.c_int * (2 * n))()
        lib.check(L.capf_jpeg_decode_batch(data, lens, n, frames.data_ptr(), hs, ws, got, dev.index or 0, torch.cuda.current_stream(dev).cuda_stream),
                  "capf_jpeg_decode_batch")
        torch.cuda.current_stream(dev).synchronize()            # `blobs` must outlive the decode
        return frames, torch.from_numpy(sizes).to(dev)

    def batch(self, indices, device, decode: str = "cv2") -> dict:
        """What ``__getitem__`` returns for `indices` (collated), with the crop done on `device`:
        images uint8 [B,H,W,3] BGR crops, keypoints_3d_gt [B,1,17,3], keypoints_2d_cpn [B,17,2], keypoints_2d_cpn_crop [B,17,2];
        plus the raw material (`frames`, `sizes`, `trans`) for frontend.flip_test_forward_from_frames.
        decode="cv2": frames decoded on the host by cv2.imread like the reference (bit-identical frames);
        decode="nvjpeg": decoded on the GPU (decode_frames)."""
        from ..utils import img
        shots = [self.labels[i] for i in indices]
        trans = np.stack([img.get_affine_transform(s["center"], s["scale"], 0, self.image_shape) for s in shots])
        if decode == "nvjpeg":
            frames_dev, sizes_dev = self.decode_frames(indices, device)
        elif decode == "cv2":
            frames = self.read_frames(indices)
            hs, ws = max(f.shape[0] for f in frames), max(f.shape[1] for f in frames)
            stack = np.zeros((len(frames), hs, ws, 3), dtype=np.uint8)
            for k, f in enumerate(frames):
                stack[k, :f.shape[0], :f.shape[1]] = f
            sizes = torch.tensor([[f.shape[0], f.shape[1]] for f in frames], dtype=torch.int32)
            frames_dev, sizes_dev = torch.from_numpy(stack).to(device), sizes.to(device)
        else:
            raise ValueError("batch: decode must be 'cv2' or 'nvjpeg'")
        as_t = lambda key, extra=(): torch.from_numpy(np.stack([np.asarray(s[key], dtype=np.float32) for s in shots]).reshape(len(shots), *extra, 17, -1)).to(device)
        return {"images": img.crop_images(frames_dev, trans, self.image_shape, sizes=sizes_dev), "frames": frames_dev, "sizes": sizes_dev, "trans": trans,
                "keypoints_3d_gt": as_t("joints_3d", (1,)), "keypoints_2d_cpn": as_t("joints_2d_cpn"), "keypoints_2d_cpn_crop": as_t("joints_2d_cpn_crop")}

    def evaluate_using_pred(self, keypoints_gt, keypoints_3d_predicted):
        return evaluate_using_pred(keypoints_gt, keypoints_3d_predicted, self.labels_action_idx)
