"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference lifting path (the parity oracle).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference legs may import
this module, and only as the checker / reported baseline.  The product (contextaware-poseformer_b200) never does.

What is restated: ``CA_PF.forward`` of /root/reference/ContextPose/mvn/models/conpose.py:30-42 and everything
under it, as plain functions over a ``state_dict`` (fp32, torch-CPU arithmetic = the same ATen kernels the
reference itself calls; the reference has no native code of its own).  Each function cites
the reference lines it follows.  The bilinear gather is additionally restated element-by-element in numpy
(``grid_sample_records``) so that the *integer* part of the sampler (corner indices, in-bounds masks) can be
compared bit-exactly with the CUDA kernels.

Pinning: the reference ships no tests or golden vectors, so this oracle is pinned to
outputs of the *reference itself* run in the authoring container: ``oracle/gen_golden.py`` imports the
unmodified reference, feeds it the seeded protocol of ``oracle/protocol.py`` and commits the results under
``tests/golden/``; ``tests/test_oracle.py`` checks this file against them (and against the live reference when
/root/reference is present).
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


# ------------------------------------------------------------------------------------------------------
# building blocks
# ------------------------------------------------------------------------------------------------------
def conv_bn(sd, x, conv, bn, stride=1, relu=False):
    """nn.Conv2d(bias=False) -> nn.BatchNorm2d in eval mode [-> ReLU] (e.g. pose_hrnet.py:82-84)."""
    w = sd[conv + ".weight"]
    y = F.conv2d(x, w, None, stride, w.shape[-1] // 2)
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     False, 0.0, BN_EPS)
    return F.relu(y) if relu else y


def basic_block(sd, x, p):
    """BasicBlock.forward, pose_hrnet.py:79-95 (no downsample inside HRNet branches)."""
    y = conv_bn(sd, x, p + ".conv1", p + ".bn1", relu=True)
    y = conv_bn(sd, y, p + ".conv2", p + ".bn2")
    return F.relu(y + x)


def bottleneck(sd, x, p, stride=1, expansion_conv3=True):
    """Bottleneck.forward, pose_hrnet.py:116-136 / networks/resnet.py:73-93 / networks/refineNet.py:26-45."""
    y = conv_bn(sd, x, p + ".conv1", p + ".bn1", relu=True)
    y = conv_bn(sd, y, p + ".conv2", p + ".bn2", stride=stride, relu=True)
    y = conv_bn(sd, y, p + ".conv3", p + ".bn3")
    r = x
    if (p + ".downsample.0.weight") in sd:
        r = conv_bn(sd, x, p + ".downsample.0", p + ".downsample.1", stride=stride)
    return F.relu(y + r)


# ------------------------------------------------------------------------------------------------------
# HRNet (pose_hrnet.py)
# ------------------------------------------------------------------------------------------------------
def _hr_module(sd, xs, p, n_br, n_blocks, multi_scale_output):
    """HighResolutionModule.forward, pose_hrnet.py:285-303.  Returns (branch outputs, fused outputs)."""
    br = []
    for i in range(n_br):
        t = xs[i]
        for b in range(n_blocks[i]):
            t = basic_block(sd, t, f"{p}.branches.{i}.{b}")
        br.append(t)
    fused = []
    for i in range(n_br if multi_scale_output else 1):
        y = None
        for j in range(n_br):
            f = f"{p}.fuse_layers.{i}.{j}"
            if j == i:
                t = br[j]
            elif j > i:     # 1x1 conv + BN + nearest upsample (:235-246)
                t = conv_bn(sd, br[j], f + ".0", f + ".1")
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:           # chain of stride-2 3x3 convs, ReLU on all but the last (:250-277)
                t = br[j]
                for k in range(i - j):
                    t = conv_bn(sd, t, f"{f}.{k}.0", f"{f}.{k}.1", stride=2, relu=(k != i - j - 1))
            y = t if y is None else y + t
        fused.append(F.relu(y))
    return br, fused


def hrnet_forward(sd, x, cfg, prefix="backbone."):
    """PoseHighResolutionNet.forward, pose_hrnet.py:464-501.  x: [B,3,H,W] -> 4 NCHW maps.

    Levels 1..3 are the *branch outputs of stage4's first module before fusion*: the module overwrites the list it
    is given (:289-290) and forward returns entries of that same list (:501)."""
    S = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = conv_bn(S, x, "conv1", "bn1", stride=2, relu=True)
    x = conv_bn(S, x, "conv2", "bn2", stride=2, relu=True)
    for b in range(4):
        x = bottleneck(S, x, f"layer1.{b}")
    ys = [x]
    kept = None
    for si, key in enumerate(("STAGE2", "STAGE3", "STAGE4")):
        st = cfg[key]
        n_br, chans = st["NUM_BRANCHES"], st["NUM_CHANNELS"]
        xs = []
        for i in range(n_br):       # transitions (:372-411; applied in forward :473-495)
            tn = f"transition{si + 1}.{i}"
            if (tn + ".0.weight") in S:                       # same-resolution 3x3 conv
                xs.append(conv_bn(S, ys[-1], tn + ".0", tn + ".1", relu=True))
            elif (tn + ".0.0.weight") in S:                   # new branch: stride-2 chain from the last map
                t = ys[-1]
                j = 0
                while (f"{tn}.{j}.0.weight") in S:
                    t = conv_bn(S, t, f"{tn}.{j}.0", f"{tn}.{j}.1", stride=2, relu=True)
                    j += 1
                xs.append(t)
            else:
                xs.append(ys[i])
        for m in range(st["NUM_MODULES"]):
            multi = not (key == "STAGE4" and m == st["NUM_MODULES"] - 1)
            br, xs = _hr_module(S, xs, f"stage{si + 2}.{m}", n_br, st["NUM_BLOCKS"], multi)
            if key == "STAGE4" and m == 0:
                kept = br
        ys = xs
    return [ys[0], kept[1], kept[2], kept[3]]


# ------------------------------------------------------------------------------------------------------
# CPN-50 (networks/*.py)
# ------------------------------------------------------------------------------------------------------
def cpn_forward(sd, x, output_shape=(64, 48), prefix="backbone."):
    """CPN.forward, networks/network.py:16-22 -> 4 maps [B,256,64,48]."""
    S = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    # ResNet.forward, networks/resnet.py:136-147
    x = conv_bn(S, x, "resnet.conv1", "resnet.bn1", stride=2, relu=True)
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, (blocks, stride) in enumerate(((3, 1), (4, 2), (6, 2), (3, 2))):
        for b in range(blocks):
            x = bottleneck(S, x, f"resnet.layer{li + 1}.{b}", stride=stride if b == 0 else 1)
        feats.append(x)
    res_out = feats[::-1]
    # globalNet.forward, networks/globalNet.py:61-83 (predict heads are computed and discarded there: skipped)
    fms, up = [], None
    for i in range(4):
        lat = conv_bn(S, res_out[i], f"global_net.laterals.{i}.0", f"global_net.laterals.{i}.1", relu=True)
        feature = lat if i == 0 else lat + up
        fms.append(feature)
        if i != 3:
            u = F.interpolate(feature, scale_factor=2, mode="bilinear", align_corners=True)
            up = conv_bn(S, u, f"global_net.upsamples.{i}.1", f"global_net.upsamples.{i}.2")
    # refineNet.forward, networks/refineNet.py:72-88
    outs = []
    for i in range(4):
        t = fms[i]
        for k in range(3 - i):
            t = bottleneck(S, t, f"refine_net.cascade.{i}.{k}")
        outs.append(F.interpolate(t, size=tuple(output_shape), mode="bilinear", align_corners=True))
    return outs


# ------------------------------------------------------------------------------------------------------
# bilinear gather, restated element-wise (ATen/native/GridSampler.h:27-36 + cuda/GridSampler.cu bilinear branch)
# ------------------------------------------------------------------------------------------------------
def grid_sample_records(grid_xy: np.ndarray, H: int, W: int, border: bool):
    """grid_xy [...,2] f32 in normalised coords -> (x_nw, y_nw int32, mask uint8 bit0..3 = nw,ne,sw,se in-bounds,
    weights [...,4] f32).  align_corners=True.  All arithmetic in float32, one rounding per operation."""
    f = np.float32
    g = grid_xy.astype(np.float32)
    ix = ((g[..., 0] + f(1)) / f(2)) * f(W - 1)
    iy = ((g[..., 1] + f(1)) / f(2)) * f(H - 1)
    if border:     # clip_coordinates
        ix = np.minimum(f(W - 1), np.maximum(ix, f(0)))
        iy = np.minimum(f(H - 1), np.maximum(iy, f(0)))
    fx, fy = np.floor(ix), np.floor(iy)
    x1, y1 = fx + f(1), fy + f(1)
    w = np.stack([(x1 - ix) * (y1 - iy), (ix - fx) * (y1 - iy), (x1 - ix) * (iy - fy), (ix - fx) * (iy - fy)], -1).astype(np.float32)
    x0 = np.clip(fx, -1e6, 1e6).astype(np.int32)
    y0 = np.clip(fy, -1e6, 1e6).astype(np.int32)
    xl, xr = (x0 >= 0) & (x0 < W), (x0 + 1 >= 0) & (x0 + 1 < W)
    yt, yb = (y0 >= 0) & (y0 < H), (y0 + 1 >= 0) & (y0 + 1 < H)
    mask = (xl & yt).astype(np.uint8) | ((xr & yt).astype(np.uint8) << 1) | ((xl & yb).astype(np.uint8) << 2) | ((xr & yb).astype(np.uint8) << 3)
    return x0, y0, mask, w


def grid_sample_gather(feat_nchw: torch.Tensor, grid_xy: np.ndarray, border: bool):
    """Pure-numpy bilinear gather for feat [B,C,H,W] and grid [B,P,2] -> [B,P,C] (used on small cases to pin
    F.grid_sample and the CUDA samplers to the same corner set)."""
    B, C, H, W = feat_nchw.shape
    x0, y0, mask, w = grid_sample_records(grid_xy, H, W, border)
    fm = feat_nchw.numpy()
    out = np.zeros(grid_xy.shape[:-1] + (C,), np.float32)
    for k, (dx, dy) in enumerate(((0, 0), (1, 0), (0, 1), (1, 1))):
        ok = (mask >> k) & 1
        xx = np.clip(x0 + dx, 0, W - 1)
        yy = np.clip(y0 + dy, 0, H - 1)
        b_idx = np.arange(B).reshape(B, *([1] * (x0.ndim - 1)))
        vals = fm[b_idx, :, yy, xx]                       # [..., C]
        out = out + (vals * w[..., k:k + 1]).astype(np.float32) * ok[..., None].astype(np.float32)
    return out


# ------------------------------------------------------------------------------------------------------
# PoseTransformer (pose_dformer.py)
# ------------------------------------------------------------------------------------------------------
def _lin(sd, x, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(sd, x, p, eps):
    return F.layer_norm(x, x.shape[-1:], sd[p + ".weight"], sd[p + ".bias"], eps)


def _mlp(sd, x, p):
    """Mlp.forward, pose_dformer.py:25-31 (dropout p=0)."""
    return _lin(sd, F.gelu(_lin(sd, x, p + ".fc1")), p + ".fc2")


def _attention(sd, x, p, heads):
    """Attention.forward, pose_dformer.py:47-59."""
    B, N, C = x.shape
    qkv = _lin(sd, x, p + ".qkv").reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = (q @ k.transpose(-2, -1)) * ((C // heads) ** -0.5)
    a = a.softmax(dim=-1)
    return _lin(sd, (a @ v).transpose(1, 2).reshape(B, N, C), p + ".proj")


def _drop(branch, scale):
    """timm DropPath (pose_dformer.py:71,101) with a given per-sample scale = bernoulli(keep) / keep over the leading dimension of
    `branch`; None = identity (eval mode, or a block whose rate is 0)."""
    if scale is None:
        return branch
    return branch * scale.view((-1,) + (1,) * (branch.ndim - 1))


def _block(sd, x, p, heads=8, drop=(None, None)):
    """Block.forward, pose_dformer.py:76-79 with LayerNorm eps 1e-6 (:166); DropPath is identity in eval."""
    x = x + _drop(_attention(sd, _ln(sd, x, p + ".norm1", 1e-6), p + ".attn", heads), drop[0])
    return x + _drop(_mlp(sd, _ln(sd, x, p + ".norm2", 1e-6), p + ".mlp"), drop[1])


def _context_block(sd, x, ref, feats, p, heads=4, samples=4, drop=(None, None)):
    """DeformableBlock.forward, pose_dformer.py:115-141 (LayerNorm eps 1e-5: default nn.LayerNorm, :89,:95)."""
    x0, xl = x[:, :1], x[:, 1:]
    b, l, pj, c = xl.shape
    t = _ln(sd, xl + x0, p + ".norm1", 1e-5)
    w = _lin(sd, t, p + ".attention_weights").view(b, l, pj, heads, samples).softmax(-1).unsqueeze(-1)
    off = _lin(sd, t, p + ".sampling_offsets").reshape(b, l, pj, heads * samples, 2).tanh()
    pos = off + ref.view(b, 1, pj, 1, -1)
    sampled = []
    for idx, fm in enumerate(feats):
        s = F.grid_sample(fm, pos[:, idx], padding_mode="border", align_corners=True).permute(0, 2, 3, 1)
        sampled.append(_lin(sd, s, f"{p}.embed_proj.{idx}"))
    s = torch.stack(sampled, 1)
    s = (w * s.view(b, l, pj, heads, samples, -1)).sum(-2).view(b, l, pj, -1)
    xl = xl + _drop(s, drop[0])
    xl = xl + _drop(_mlp(sd, _ln(sd, xl, p + ".norm2", 1e-5), p + ".mlp"), drop[1])
    return torch.cat([x0, xl], 1), pos


def lifter_forward(sd, kp2d, ref, feats, levels=4, prefix="volume_net.", trace=None, context=True, depth=None, drop=None):
    """PoseTransformer.forward, pose_dformer.py:210-241.  kp2d, ref: [B,17,2]; feats: 4 NCHW maps -> [B,1,17,3].

    context=False, depth=config.depth restates the MPI-INF-3DHP variant (ContextPose_mpi/model/pose_dformer.py:236-262):
    no DeformableBlocks, `depth` res / joint blocks; the caller applies that variant's output layout (mpi_output).
    drop: train-mode DropPath scales {"context_blocks" | "res_blocks" | "joint_blocks": [(s_branch1, s_branch2)] * depth}."""
    depth = levels if depth is None else depth
    dp = lambda grp, i: (drop[grp][i] if drop is not None else (None, None))
    S = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    b, p, _ = kp2d.shape
    x = _lin(S, kp2d, "coord_embed")
    fr = [F.grid_sample(fm, ref.unsqueeze(-2), align_corners=True).squeeze(-1).permute(0, 2, 1) for fm in feats]
    fr = [_lin(S, fr[i], f"feat_embed.{i}") for i in range(len(feats))]
    x = torch.stack([x, *fr], 1) + S["Spatial_pos_embed"]
    if trace is not None:
        trace["tokens_embed"] = x.clone()
    for i in range(levels if context else 0):
        x, pos = _context_block(S, x, ref, feats, f"context_blocks.{i}", drop=dp("context_blocks", i))
        if trace is not None and i == 0:
            trace["deform_pos0"] = pos.clone()
    if trace is not None:
        trace["tokens_context"] = x.clone()
    x = x.permute(0, 2, 1, 3).reshape(b * p, levels + 1, -1)          # 'b l p c -> (b p) l c'
    for i in range(depth):
        x = _block(S, x, f"res_blocks.{i}", drop=dp("res_blocks", i))
    if trace is not None:
        trace["tokens_res"] = x.clone()
    x = x.reshape(b, p, -1)                                           # '(b p) l c -> b p (l c)'
    for i in range(depth):
        x = _block(S, x, f"joint_blocks.{i}", drop=dp("joint_blocks", i))
    if trace is not None:
        trace["tokens_joint"] = x.clone()
    x = _lin(S, _ln(S, x, "head.0", 1e-5), "head.1")
    return x.view(b, 1, p, -1)


# ------------------------------------------------------------------------------------------------------
# whole path
# ------------------------------------------------------------------------------------------------------
def normalize_crop_(crop: torch.Tensor):
    """conpose.py:34-35, in place on the caller's tensor (hard-coded 192x256 crop)."""
    crop[..., :2] /= torch.tensor([192 // 2, 256 // 2], device=crop.device)
    crop[..., :2] -= torch.tensor([1, 1], device=crop.device)
    return crop


def mpi_forward(sd, bb_cfg, depth, images, kp2d, crop):
    """VolumetricTriangulationNet.forward of the MPI-INF-3DHP tree (ContextPose_mpi/model/conpose.py:30-42 +
    pose_dformer.py:236-262): HRNet features, lifter without context blocks, output (x[b,3,1,17,1], None)."""
    with torch.no_grad():
        x = images.permute(0, 3, 1, 2).contiguous()
        ref = normalize_crop_(crop)
        feats = hrnet_forward(sd, x, bb_cfg)
        y = lifter_forward(sd, kp2d, ref, feats, context=False, depth=depth)           # [b,1,p,3]
        b, _, p, _ = y.shape
        return y.reshape(b, 1, p, 3, 1).permute(0, 3, 1, 2, 4).contiguous(), None


def ca_pf_forward(sd, backbone, bb_cfg, images, kp2d, crop, trace=None):
    """CA_PF.forward, conpose.py:30-42.  images [B,H,W,3]; mutates `crop` like the reference.  fp32."""
    with torch.no_grad():
        x = images.permute(0, 3, 1, 2).contiguous()
        ref = normalize_crop_(crop)
        feats = cpn_forward(sd, x) if backbone == "cpn" else hrnet_forward(sd, x, bb_cfg)
        if trace is not None:
            trace["features"] = feats
        return lifter_forward(sd, kp2d, ref, feats, trace=trace)


# ------------------------------------------------------------------------------------------------------
# f1: pre-processing + flip-test front end (TEST-ONLY restatement of mvn/datasets/utils.py:15-88).  PINNED:
# oracle/gen_golden_prefetch.py runs the reference's own data_prefetcher on the CPU with its CUDA entry points stubbed
# and tests/test_frontend.py checks these functions against its outputs exactly (tests/golden/prefetch_cases.npz).
# flip_test_merge (train.py:177-180, code inside a script function) is a line-by-line restatement.
# ------------------------------------------------------------------------------------------------------
JOINTS_LEFT = [4, 5, 6, 11, 12, 13]      # mvn/datasets/utils.py:12
JOINTS_RIGHT = [1, 2, 3, 14, 15, 16]     # mvn/datasets/utils.py:13


def prefetch_images(images_u8: torch.Tensor, backbone: str) -> torch.Tensor:
    """mvn/datasets/utils.py:45-50 on a uint8 [B,H,W,3] BGR batch -> fp32 RGB normalised."""
    images = torch.flip(images_u8, [-1])
    if backbone in ("hrnet_32", "hrnet_48"):
        mean = torch.tensor([0.485, 0.456, 0.406])
        std = torch.tensor([0.229, 0.224, 0.225])
        images = (images / 255.0 - mean) / std
    else:
        mean = torch.tensor([122.7717, 115.9465, 102.9801]).view(1, 1, 1, 3)
        mean /= 255.
        images = images / 255.0 - mean
    return images.float()


def prefetch_flip_test(images_u8, kp2d, kp2d_crop, backbone):
    """mvn/datasets/utils.py:66-80: the stacked (plain, mirrored) inputs of the flip test."""
    images = prefetch_images(images_u8, backbone)
    images = torch.stack([images, torch.flip(images, [2])], dim=1)
    k = kp2d.clone()
    k[..., 0] *= -1
    k[..., JOINTS_LEFT + JOINTS_RIGHT, :] = k[..., JOINTS_RIGHT + JOINTS_LEFT, :]
    c = kp2d_crop.clone()
    c[:, :, 0] = 192 - c[:, :, 0] - 1
    c[:, JOINTS_LEFT + JOINTS_RIGHT] = c[:, JOINTS_RIGHT + JOINTS_LEFT]
    return images, torch.stack([kp2d, k], dim=1), torch.stack([kp2d_crop, c], dim=1)


def prefetch_targets(keypoints_3d_gt):
    """mvn/datasets/utils.py:52-53: the 3D target made root-relative (joint 0 subtracted from joints 1.., then zeroed)."""
    gt = keypoints_3d_gt.clone()
    gt[:, :, 1:] -= gt[:, :, :1]
    gt[:, :, 0] = 0
    return gt.float()


def flip_test_merge(pred, pred_flip):
    """train.py:177-180."""
    pred_flip = pred_flip.clone()
    pred_flip[:, :, :, 0] *= -1
    pred_flip[:, :, JOINTS_LEFT + JOINTS_RIGHT] = pred_flip[:, :, JOINTS_RIGHT + JOINTS_LEFT]
    return torch.mean(torch.cat((pred, pred_flip), dim=1), dim=1, keepdim=True)


# ------------------------------------------------------------------------------------------------------
# f4: crop_image (mvn/utils/img.py:16-69) -- the per-frame CPU work of Human36MSingleViewDataset.__getitem__
# (human36m.py:554-584).  The arithmetic lives in OpenCV (cv2.getAffineTransform / cv2.warpAffine; opencv-python is a
# reference dependency, 4.13.0 in this image): restated from its published algorithm (imgproc/imgwarp.cpp,
# WarpAffineInvoker + remapBilinear with the 15-bit bilinear table) and PINNED two ways: tests/golden/crop_cases.npz
# holds outputs of the reference's own crop_image (oracle/gen_golden_crop.py), and tests/test_crop.py compares against
# cv2.warpAffine directly wherever cv2 imports.
# ------------------------------------------------------------------------------------------------------
def affine_transform(center, scale, output_size):
    """get_affine_transform(center, scale, 0, output_size) (img.py:16-48): the 2x3 frame -> crop map through three point
    pairs held in float32, solved in float64 like cv2.getAffineTransform."""
    center = np.array(center)
    scale_tmp = np.array(scale) * 200.0
    src_w, dst_w, dst_h = scale_tmp[0], output_size[0], output_size[1]
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = center + np.array([0, (src_w - 1) * -0.5], np.float32)
    dst[0] = [(dst_w - 1) * 0.5, (dst_h - 1) * 0.5]
    dst[1] = np.array([(dst_w - 1) * 0.5, (dst_h - 1) * 0.5]) + np.array([0, (dst_w - 1) * -0.5], np.float32)
    for p in (src, dst):                                   # get_3rd_point: b + (-(a-b).y, (a-b).x), float32
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], dtype=np.float32)
    a = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        a[i, 0:2], a[i, 2] = src[i], 1.0
        a[i + 3, 3:5], a[i + 3, 5] = src[i], 1.0
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(a, b).reshape(2, 3)


def invert_affine(trans):
    """The crop -> frame map cv2.warpAffine derives from `trans` (no WARP_INVERSE_MAP), same operation order, float64."""
    m = np.array(trans, dtype=np.float64).reshape(6).copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0], m[1], m[3], m[4] = a11, m[1] * -d, m[3] * -d, a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def warp_affine_u8(image, trans, dsize):
    """cv2.warpAffine(image, trans, dsize, flags=cv2.INTER_LINEAR) for uint8 HWC images, constant 0 border: bit-exact."""
    wo, ho = int(dsize[0]), int(dsize[1])
    m = invert_affine(trans)
    rnd = lambda v: np.rint(v).astype(np.int64)            # cvRound: half to even
    x, y = np.arange(wo), np.arange(ho)
    adelta, bdelta = rnd(m[0] * x * 1024), rnd(m[3] * x * 1024)
    x0, y0 = rnd((m[1] * y + m[2]) * 1024) + 16, rnd((m[4] * y + m[5]) * 1024) + 16
    xf, yf = (x0[:, None] + adelta[None, :]) >> 5, (y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(xf >> 5, -32768, 32767), np.clip(yf >> 5, -32768, 32767)
    fx, fy = (xf & 31)[..., None], (yf & 31)[..., None]
    h, w = image.shape[:2]

    def corner(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        return image[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].astype(np.int64) * ok[..., None]

    t = (corner(sy, sx) * ((32 - fx) * (32 - fy)) + corner(sy, sx + 1) * (fx * (32 - fy))
         + corner(sy + 1, sx) * ((32 - fx) * fy) + corner(sy + 1, sx + 1) * (fx * fy))
    return ((t + 512) >> 10).astype(np.uint8)


def crop_image(image, center, scale, output_size):
    """img.py:51-69."""
    return warp_affine_u8(image, affine_transform(center, scale, output_size), output_size)


# ------------------------------------------------------------------------------------------------------
# f4: evaluation reducer -- evaluate_using_pred (mvn/datasets/human36m.py:358-422) over MPJPE / P_MPJPE / MPJVE
# (mvn/models/loss.py:16-22, :25-68, :87-101).  Pinned by tests/golden/eval_cases.npz (oracle/gen_golden_eval.py ran the
# reference's own method on seeded poses).
# ------------------------------------------------------------------------------------------------------
H36M_ACTIONS = ["Directions", "Discussion", "Eating", "Greeting", "Phoning", "Posing", "Purchases", "Sitting", "SittingDown",
                "Smoking", "TakingPhoto", "Waiting", "Walking", "WalkingDog", "WalkingTogether"]
H36M_ACTION_NAMES = [f"{a}-{t}" for a in H36M_ACTIONS for t in (1, 2)]       # human36m.py:18-33


def mpjpe(pred, gt):
    """loss.py:16-22 (torch, fp32)."""
    return float(torch.mean(torch.norm(pred - gt, dim=len(gt.shape) - 1)))


def p_mpjpe(pred, gt):
    """loss.py:25-68 on numpy arrays [F,J,3] (fp32 in the reference's call, human36m.py:373)."""
    mu_x, mu_y = np.mean(gt, axis=1, keepdims=True), np.mean(pred, axis=1, keepdims=True)
    x0, y0 = gt - mu_x, pred - mu_y
    norm_x = np.sqrt(np.sum(x0 ** 2, axis=(1, 2), keepdims=True))
    norm_y = np.sqrt(np.sum(y0 ** 2, axis=(1, 2), keepdims=True))
    x0, y0 = x0 / norm_x, y0 / norm_y
    u, s, vt = np.linalg.svd(np.matmul(x0.transpose(0, 2, 1), y0))
    v = vt.transpose(0, 2, 1)
    sign = np.sign(np.expand_dims(np.linalg.det(np.matmul(v, u.transpose(0, 2, 1))), axis=1))
    v[:, :, -1] *= sign
    s[:, -1] *= sign.flatten()
    r = np.matmul(v, u.transpose(0, 2, 1))
    a = np.expand_dims(np.sum(s, axis=1, keepdims=True), axis=2) * norm_x / norm_y
    t = mu_x - a * np.matmul(mu_y, r)
    return np.mean(np.linalg.norm(a * np.matmul(pred, r) + t - gt, axis=len(gt.shape) - 1))


def mpjve(pred, gt):
    """loss.py:87-101."""
    return np.mean(np.linalg.norm(np.diff(pred, axis=0) - np.diff(gt, axis=0), axis=len(gt.shape) - 1))


def evaluate_using_pred(keypoints_gt, keypoints_3d_predicted, labels_action_idx, action_names=None):
    """human36m.py:358-422: per-action scores with the two trials of an action merged (frame-count weighted)."""
    action_names = action_names or H36M_ACTION_NAMES
    scores = {}
    for k, name in enumerate(action_names):
        m = torch.from_numpy(labels_action_idx == k)
        n = int(m.sum())
        p, g = keypoints_3d_predicted[m], keypoints_gt[m]
        pn, gn = p.squeeze().cpu().numpy(), g.squeeze().cpu().numpy()
        scores[name] = {"MPJPE": n * mpjpe(p, g), "P_MPJPE": n * p_mpjpe(pn, gn), "MPJVE": n * mpjve(pn, gn), "frame_count": n}
    for base in [x[:-2] for x in action_names if x.endswith("-1")]:
        both = [scores.pop(f"{base}-{t}") for t in (1, 2)]
        scores[base] = {k: both[0][k] + both[1][k] for k in both[0]}
    return {k: {m: v[m] / v["frame_count"] for m in ("MPJPE", "P_MPJPE", "MPJVE")} for k, v in scores.items()}


# ------------------------------------------------------------------------------------------------------
# f2 groundwork: one training step of volume_net (train.py:186-201; optimiser train.py:337-345: AdamW over
# volume_net.parameters(), lr = config.train.volume_net_lr, weight_decay 0.1, no gradient clipping by default).
# Gradients come from autograd over the restated forward above (every op there is a differentiable torch op; the backbone
# is frozen, conpose.py:23-25).  DropPath is the identity here (eval-mode blocks): the stochastic-depth mask of train mode
# is a separate, seeded concern.  Pinned by tests/golden/grad_hrnet32_b2_128x96.npz (oracle/gen_golden_grad.py ran
# autograd and torch.optim.AdamW on the unmodified reference model).
# ------------------------------------------------------------------------------------------------------
def volume_net_loss_and_grads(sd, backbone, bb_cfg, images, kp2d, crop, gt, drop=None):
    """MPJPE loss (loss.py:16-22) of the forward and d loss / d volume_net parameters.  Returns (loss, {name: grad}).
    drop: DropPath scales of a train-mode step (see lifter_forward); None = eval-mode blocks."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if k.startswith("volume_net.") and v.is_floating_point()}
    sd2 = dict(sd)
    sd2.update(leaves)
    with torch.no_grad():                                   # frozen backbone (conpose.py:23-25): features are constants
        x = images.permute(0, 3, 1, 2).contiguous()
        ref = normalize_crop_(crop.clone())
        feats = cpn_forward(sd, x) if backbone == "cpn" else hrnet_forward(sd, x, bb_cfg)
    with torch.enable_grad():
        pred = lifter_forward(sd2, kp2d, ref, feats, drop=drop)
        loss = torch.mean(torch.norm(pred - gt, dim=len(gt.shape) - 1))
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    return float(loss.detach()), {k: (g if g is not None else torch.zeros_like(v)) for (k, v), g in zip(leaves.items(), grads)}


def adamw_step(param, grad, exp_avg, exp_avg_sq, step, lr, weight_decay=0.1, betas=(0.9, 0.999), eps=1e-8):
    """torch.optim.AdamW's update (decoupled weight decay; the defaults train.py:345 leaves in place), one tensor, step
    counted from 1.  Returns (param, exp_avg, exp_avg_sq)."""
    param = param * (1 - lr * weight_decay)
    exp_avg = exp_avg * betas[0] + grad * (1 - betas[0])
    exp_avg_sq = exp_avg_sq * betas[1] + grad * grad * (1 - betas[1])
    denom = (exp_avg_sq.sqrt() / (1 - betas[1] ** step) ** 0.5) + eps
    return param - (lr / (1 - betas[0] ** step)) * exp_avg / denom, exp_avg, exp_avg_sq
