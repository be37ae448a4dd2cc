"""Host mirror of the reference's crop helpers (mvn/utils/img.py:16-69) over libcapf_b200.

``crop_image(image, center, scale, output_size)`` keeps the reference's name and argument meaning; the warp itself --
``cv2.warpAffine(image, trans, output_size, flags=cv2.INTER_LINEAR)`` in the reference, one CPU call per frame inside
``Human36MSingleViewDataset.__getitem__`` (human36m.py:569-571) -- is ``CAPF_OP_WARP_AFFINE_U8`` on the GPU, for a whole
batch of decoded frames at once, returning the same bytes OpenCV does.  ``crop_images(..., normalise=backbone)`` goes
straight to the normalised fp32 RGB tensor ``CA_PF.forward`` takes (crop + ``data_prefetcher.preload`` in one kernel).

Only rotation-free boxes are built here (the reference always passes rot = 0 and no shift); ``crop_images`` itself takes
arbitrary 2x3 maps.  There is no CPU path: CPU tensors are rejected.
"""
import ctypes

import numpy as np
import torch

from ... import lib


def get_3rd_point(a, b):
    """img.py:11-13."""
    direct = a - b
    return b + np.array([-direct[1], direct[0]], dtype=np.float32)


def get_affine_transform(center, scale, rot, output_size, shift=None, inv=0):
    """img.py:16-48 for the arguments the reference uses (rot = 0, no shift, inv = 0): the 2x3 float64 frame -> crop map
    through three float32 point pairs (cv2.getAffineTransform = one 6x6 solve in float64)."""
    if rot != 0 or inv or (shift is not None and np.any(np.asarray(shift) != 0)):
        raise NotImplementedError("get_affine_transform: the reference's crop path uses rot = 0, shift = 0, inv = 0")
    center = np.array(center)
    scale_tmp = np.array(scale) * 200.0
    src_w, dst_w, dst_h = scale_tmp[0], output_size[0], output_size[1]
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center
    src[1, :] = center + np.array([0, (src_w - 1) * -0.5], np.float32)
    dst[0, :] = [(dst_w - 1) * 0.5, (dst_h - 1) * 0.5]
    dst[1, :] = np.array([(dst_w - 1) * 0.5, (dst_h - 1) * 0.5]) + np.array([0, (dst_w - 1) * -0.5], np.float32)
    src[2, :] = get_3rd_point(src[0, :], src[1, :])
    dst[2, :] = get_3rd_point(dst[0, :], dst[1, :])
    a, b = np.zeros((6, 6)), np.zeros(6)
    for i in range(3):
        a[i, 0:2], a[i, 2] = src[i], 1.0
        a[i + 3, 3:5], a[i + 3, 5] = src[i], 1.0
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(a, b).reshape(2, 3)


def invert_affine(trans):
    """frame -> crop map to the crop -> frame map the kernel samples with, in OpenCV's operation order (float64)."""
    m = np.array(trans, dtype=np.float64).reshape(6).copy()
    det = m[0] * m[4] - m[1] * m[3]
    det = 1.0 / det if det != 0 else 0.0
    a11, a22 = m[4] * det, m[0] * det
    m[0], m[1], m[3], m[4] = a11, m[1] * -det, m[3] * -det, a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def crop_images(frames: torch.Tensor, trans, output_size, sizes: torch.Tensor = None, normalise: str = None, mirror: bool = False,
                out: torch.Tensor = None, minv: torch.Tensor = None) -> torch.Tensor:
    """frames: uint8 [B,Hs,Ws,3] on the GPU (frames of different size padded to one storage, live (h, w) in `sizes`
    int32 [B,2]); trans: [B,2,3] frame -> crop maps; output_size = (W, H) like cv2.  Returns uint8 [B,H,W,3], or with
    ``normalise="hrnet_32"|"hrnet_48"|"cpn"`` the fp32 RGB tensor of data_prefetcher.preload (mirrored along W if asked).
    `minv` (CUDA float64 [B,6], rows = invert_affine(trans[b])) replaces `trans` when the caller keeps the maps on the
    device (no host work, no copy: the call is then one kernel launch)."""
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError("crop_images expects uint8 frames [B,Hs,Ws,3]")
    if not frames.is_cuda:
        raise lib.CapfError("crop_images runs on a B200 through libcapf_b200; got a CPU tensor (no CPU path)")
    frames = frames.contiguous()
    B, Hs, Ws, _ = frames.shape
    wo, ho = int(output_size[0]), int(output_size[1])
    if minv is None:
        trans = np.asarray(trans, dtype=np.float64).reshape(B, 2, 3)
        minv = torch.from_numpy(np.stack([invert_affine(t) for t in trans])).to(frames.device)
    elif minv.dtype != torch.float64 or tuple(minv.shape) != (B, 6) or not minv.is_cuda or not minv.is_contiguous():
        raise ValueError("crop_images: `minv` must be a contiguous CUDA float64 [B,6] tensor")
    if sizes is not None:
        if sizes.dtype != torch.int32 or tuple(sizes.shape) != (B, 2) or not sizes.is_cuda:
            raise ValueError("crop_images: `sizes` must be a CUDA int32 [B,2] tensor of (h, w)")
        sizes = sizes.contiguous()
    odt = torch.uint8 if normalise is None else torch.float32
    if out is None:
        out = torch.empty(B, ho, wo, 3, dtype=odt, device=frames.device)
    elif out.dtype != odt or tuple(out.shape) != (B, ho, wo, 3) or not out.is_contiguous():
        raise ValueError(f"crop_images: `out` must be a contiguous {odt} [B,{ho},{wo},3] tensor")
    op = lib.CapfOp()
    op.kind, op.dtype_in, op.dtype_out = lib.OP_WARP_AFFINE_U8, lib.F32, lib.F32
    vals = [B, Hs, Ws, ho, wo, 0 if normalise is None else 1, 1 if mirror else 0, 0 if normalise == "cpn" else 1]
    for n, v in enumerate(vals):
        op.i[n] = v
    op.inp[0], op.inp[1] = frames.data_ptr(), minv.data_ptr()
    op.inp[2] = sizes.data_ptr() if sizes is not None else None
    ms = None
    if normalise is not None:
        from ... import frontend
        ms = frontend.normalisation_params(normalise, frames.device)
        op.inp[3] = ms.data_ptr()
    elif mirror:
        raise ValueError("crop_images: mirror applies to the normalised output (the flip-test copy) only")
    op.out[0] = out.data_ptr()
    st = torch.cuda.current_stream(frames.device).cuda_stream
    lib.check(lib.load().capf_op_run(ctypes.byref(op), frames.device.index or 0, st), "warp_affine_u8")
    return out


def crop_image(image: torch.Tensor, center, scale, output_size) -> torch.Tensor:
    """img.py:51-69 for one frame: uint8 [H,W,3] (GPU) -> uint8 [output H, output W, 3]."""
    trans = get_affine_transform(center, scale, 0, output_size)
    return crop_images(image[None], trans[None], output_size)[0]
