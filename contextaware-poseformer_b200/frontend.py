"""Pre-processing + flip-test front end of the lifting path ("next" row f1).

Mirrors what the reference does immediately before and after ``model(images, kp2d, kp2d_crop)`` at evaluation time:

* ``data_prefetcher.preload`` (mvn/datasets/utils.py:33-82): uint8 BGR HWC crops -> RGB, ``/ 255``, mean/std (HRNet) or
  mean only (CPN), and for the flip test a copy mirrored along W together with mirrored keypoints
  (``x -> -x`` for the screen-normalised 2D pose, ``x -> 192 - x - 1`` for the crop pixels, left/right joints swapped);
* the flip-test merge in ``one_epoch_full`` (train.py:170-181): the mirrored prediction is un-mirrored (x negated,
  left/right joints swapped) and averaged with the plain one.

The image transform is one CUDA kernel (``CAPF_OP_PREPROCESS_U8``: 3 bytes in, 12 bytes out per pixel, bit-exact with the
torch expression); the host uploads 4x fewer bytes than with fp32 images.  The two flip-test passes run as ONE forward
of 2*B frames (frames are independent), i.e. one CUDA graph launch instead of two.
"""
import ctypes

import numpy as np
import torch

from . import lib

JOINTS_LEFT = [4, 5, 6, 11, 12, 13]      # mvn/datasets/utils.py:12
JOINTS_RIGHT = [1, 2, 3, 14, 15, 16]     # mvn/datasets/utils.py:13
CROP_WIDTH = 192                          # the hard-coded crop width of the flip (utils.py:56,75)


_NORM_CACHE = {}


def normalisation_params(backbone: str, device) -> torch.Tensor:
    """[mean R,G,B | std R,G,B] exactly as data_prefetcher.__init__ builds them (utils.py:24-29), fp32.  Built once per
    (backbone, device) like the prefetcher does in its constructor: the upload is a synchronous pageable copy."""
    key = (backbone, str(torch.device(device)))
    if key not in _NORM_CACHE:
        _NORM_CACHE[key] = _normalisation_params(backbone, device)
    return _NORM_CACHE[key]


def _normalisation_params(backbone: str, device) -> torch.Tensor:
    if backbone in ("hrnet_32", "hrnet_48"):
        mean = torch.tensor([0.485, 0.456, 0.406])
        std = torch.tensor([0.229, 0.224, 0.225])
    elif backbone == "cpn":
        mean = torch.tensor([122.7717, 115.9465, 102.9801])
        mean /= 255.
        std = torch.ones(3)
    else:
        raise ValueError(f"unknown backbone {backbone!r}")
    return torch.cat([mean, std]).to(device=device, dtype=torch.float32)


def preprocess(images_u8: torch.Tensor, backbone: str, mirror: bool = False, out: torch.Tensor = None) -> torch.Tensor:
    """uint8 [B,H,W,3] BGR (device) -> fp32 [B,H,W,3] RGB normalised, optionally mirrored along W."""
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
        raise ValueError("preprocess expects uint8 images [B,H,W,3]")
    if not images_u8.is_cuda:
        raise lib.CapfError("preprocess runs on a B200 through libcapf_b200; got a CPU tensor (no CPU path)")
    images_u8 = images_u8.contiguous()
    B, H, W, _ = images_u8.shape
    if out is None:
        out = torch.empty(B, H, W, 3, dtype=torch.float32, device=images_u8.device)
    elif out.dtype != torch.float32 or tuple(out.shape) != (B, H, W, 3) or not out.is_contiguous():
        raise ValueError("preprocess: `out` must be a contiguous fp32 [B,H,W,3] tensor")
    ms = normalisation_params(backbone, images_u8.device)
    op = lib.CapfOp()
    op.kind, op.dtype_in, op.dtype_out = lib.OP_PREPROCESS_U8, lib.F32, lib.F32
    for n, v in enumerate([B, H, W, 1 if mirror else 0, 0 if backbone == "cpn" else 1]):
        op.i[n] = v
    op.inp[0], op.inp[1], op.out[0] = images_u8.data_ptr(), ms.data_ptr(), out.data_ptr()
    st = torch.cuda.current_stream(images_u8.device).cuda_stream
    lib.check(lib.load().capf_op_run(ctypes.byref(op), images_u8.device.index or 0, st), "preprocess_u8")
    return out


def flip_keypoints(kp2d: torch.Tensor, kp2d_crop: torch.Tensor):
    """Mirrored copies of the 2D inputs (utils.py:69-77): x negated / reflected in the 192-wide crop, L/R joints swapped."""
    k = kp2d.clone()
    k[..., 0] *= -1
    k[..., JOINTS_LEFT + JOINTS_RIGHT, :] = k[..., JOINTS_RIGHT + JOINTS_LEFT, :]
    c = kp2d_crop.clone()
    c[:, :, 0] = CROP_WIDTH - c[:, :, 0] - 1
    c[:, JOINTS_LEFT + JOINTS_RIGHT] = c[:, JOINTS_RIGHT + JOINTS_LEFT]
    return k, c


def root_relative(keypoints_3d_gt: torch.Tensor) -> torch.Tensor:
    """The 3D target as the prefetcher hands it on (utils.py:52-53): joints 1.. relative to joint 0, joint 0 zeroed.
    [B,1,17,3] -> new fp32 tensor (the reference edits the loader's tensor in place)."""
    gt = keypoints_3d_gt.clone()
    gt[:, :, 1:] -= gt[:, :, :1]
    gt[:, :, 0] = 0
    return gt.float()


def merge_flip_test(pred: torch.Tensor, pred_flip: torch.Tensor) -> torch.Tensor:
    """train.py:177-180: un-mirror the second prediction and average.  pred, pred_flip: [B,1,17,3]."""
    pf = pred_flip.clone()
    pf[:, :, :, 0] *= -1
    pf[:, :, JOINTS_LEFT + JOINTS_RIGHT] = pf[:, :, JOINTS_RIGHT + JOINTS_LEFT]
    return torch.mean(torch.cat((pred, pf), dim=1), dim=1, keepdim=True)


def flip_test_forward(model, images_u8: torch.Tensor, kp2d: torch.Tensor, kp2d_crop: torch.Tensor) -> torch.Tensor:
    """Evaluation-time forward with flip test from raw uint8 crops: [B,H,W,3] u8, [B,17,2], [B,17,2] -> [B,1,17,3].

    Equivalent to data_prefetcher.preload(flip_test=True) + the two model calls + merge of train.py:170-181; like the
    reference it leaves the caller's ``kp2d_crop`` untouched (train.py passes ``.clone()``).  Both passes share one
    forward of 2*B frames."""
    B, H, W, _ = images_u8.shape
    dev = images_u8.device
    backbone = model.backbone_type
    static = model.static_inputs(2 * B, H, W, dev)["images"]          # write both halves straight into the plan's input
    preprocess(images_u8, backbone, mirror=False, out=static[:B])
    preprocess(images_u8, backbone, mirror=True, out=static[B:])
    kf, cf = flip_keypoints(kp2d.float(), kp2d_crop.float())
    kp2 = torch.cat([kp2d.float(), kf], dim=0)
    crop2 = torch.cat([kp2d_crop.float().clone(), cf], dim=0)
    pred2 = model(static, kp2, crop2)
    return merge_flip_test(pred2[:B], pred2[B:])


def flip_test_forward_from_frames(model, frames_u8: torch.Tensor, trans, kp2d: torch.Tensor, kp2d_crop: torch.Tensor,
                                  sizes: torch.Tensor = None, image_shape=(CROP_WIDTH, 256)) -> torch.Tensor:
    """The whole per-batch chain of an evaluation step from decoded camera frames: crop_image (human36m.py:569-571) +
    data_prefetcher.preload(flip_test=True) + both model calls + merge (train.py:170-181).

    frames_u8: uint8 [B,Hs,Ws,3] BGR on the GPU (as cv2.imread delivers them), trans: [B,2,3] frame -> crop maps
    (mvn.utils.img.get_affine_transform(center, scale, 0, image_shape)), image_shape = (W, H) of the crop.  Two kernels
    write the straight and the mirrored normalised crops into the plan's input; nothing uint8 or fp32 is staged between."""
    from .mvn.utils import img
    B = frames_u8.shape[0]
    dev = frames_u8.device
    wo, ho = int(image_shape[0]), int(image_shape[1])
    backbone = model.backbone_type
    static = model.static_inputs(2 * B, ho, wo, dev)["images"]
    minv = torch.from_numpy(np.stack([img.invert_affine(t) for t in np.asarray(trans, dtype=np.float64).reshape(B, 2, 3)])).to(dev)
    img.crop_images(frames_u8, None, image_shape, sizes=sizes, normalise=backbone, mirror=False, out=static[:B], minv=minv)
    img.crop_images(frames_u8, None, image_shape, sizes=sizes, normalise=backbone, mirror=True, out=static[B:], minv=minv)
    kf, cf = flip_keypoints(kp2d.float(), kp2d_crop.float())
    kp2 = torch.cat([kp2d.float(), kf], dim=0)
    crop2 = torch.cat([kp2d_crop.float().clone(), cf], dim=0)
    pred2 = model(static, kp2, crop2)
    return merge_flip_test(pred2[:B], pred2[B:])
