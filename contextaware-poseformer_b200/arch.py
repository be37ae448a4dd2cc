"""Topology of the three backbones and the lifter, written once as *walkers*.

A walker drives a visitor through the network in execution order.  Two visitors exist:
  * ``ModuleVisitor`` (mvn/models/_tree.py) registers nn.Conv2d / nn.BatchNorm2d parameter holders under the
    reference's dotted names, so ``state_dict()`` has exactly the reference's keys and shapes;
  * ``ProgramBuilder`` (program.py) emits the flat op program that libcapf_b200 executes.
Keeping one description for both is what guarantees the checkpoint layout and the executed graph agree.

Reference behaviour being restated (paths relative to /root/reference/ContextPose/mvn/models/):
  HRNet : pose_hrnet.py:314-370 (construction), :372-411 (transitions), :432-462 (stages),
          :285-303 (module forward incl. the in-place list update), :464-501 (forward / returned maps)
  CPN   : networks/network.py:8-22, networks/resnet.py:95-147, networks/globalNet.py:5-83,
          networks/refineNet.py:3-88
"""

RELU, NONE, GELU = "relu", "none", "gelu"


class T:
    """Shape-only activation handle: NHWC tensor [N,H,W,C] (visitor-specific payload in .ref)."""
    __slots__ = ("H", "W", "C", "ref")

    def __init__(self, H, W, C, ref=None):
        self.H, self.W, self.C, self.ref = H, W, C, ref


# ------------------------------------------------------------------------------------------------------
# HRNet-W32 / W48
# ------------------------------------------------------------------------------------------------------
def _stage_cfgs(cfg):
    out = []
    for key in ("STAGE2", "STAGE3", "STAGE4"):
        st = cfg[key]
        if st["BLOCK"] != "BASIC" or st["FUSE_METHOD"] != "SUM":
            raise NotImplementedError(f"{key}: only BLOCK=BASIC / FUSE_METHOD=SUM (the shipped HRNet configs)")
        out.append((int(st["NUM_MODULES"]), int(st["NUM_BRANCHES"]), [int(b) for b in st["NUM_BLOCKS"]],
                    [int(c) for c in st["NUM_CHANNELS"]]))
    return out


def walk_hrnet(v, x, cfg):
    """x: T of the NHWC image.  Returns the 4 maps of pose_hrnet.py:501."""
    x = v.conv("conv1", "bn1", x, 64, k=3, stride=2, act=RELU)
    x = v.conv("conv2", "bn2", x, 64, k=3, stride=2, act=RELU)
    for b in range(4):                                   # layer1 = 4 Bottleneck(planes 64)  (:328, :413-430)
        q = f"layer1.{b}"
        ident = x
        if x.C != 256:
            ident = v.conv(q + ".downsample.0", q + ".downsample.1", x, 256, k=1, act=NONE)
        t = v.conv(q + ".conv1", q + ".bn1", x, 64, k=1, act=RELU)
        t = v.conv(q + ".conv2", q + ".bn2", t, 64, k=3, act=RELU)
        x = v.conv(q + ".conv3", q + ".bn3", t, 256, k=1, act=RELU, residual=ident)

    ys = [x]
    kept = None
    for si, (n_mod, n_br, n_blk, chans) in enumerate(_stage_cfgs(cfg)):
        sname = f"stage{si + 2}"
        # ---- transition (:372-411 / forward :473-495): every transition conv reads ys[-1]
        xs = []
        n_pre = len(ys)
        for i in range(n_br):
            tn = f"transition{si + 1}.{i}"
            if i < n_pre:
                if chans[i] != ys[i].C:
                    xs.append(v.conv(tn + ".0", tn + ".1", ys[-1], chans[i], k=3, act=RELU))
                else:
                    xs.append(ys[i])
            else:
                t = ys[-1]
                c_pre = ys[-1].C
                for j in range(i + 1 - n_pre):
                    cout = chans[i] if j == i - n_pre else c_pre
                    t = v.conv(f"{tn}.{j}.0", f"{tn}.{j}.1", t, cout, k=3, stride=2, act=RELU)
                xs.append(t)
        # ---- HighResolutionModules
        for m in range(n_mod):
            multi = not (si == 2 and m == n_mod - 1)     # stage4's last module is single-output (:446-449)
            br = []
            for i in range(n_br):
                t = xs[i]
                for blk in range(n_blk[i]):              # BasicBlock (:66-95)
                    q = f"{sname}.{m}.branches.{i}.{blk}"
                    if t.C != chans[i]:
                        raise NotImplementedError("branch channel change (never produced by the reference configs)")
                    u = v.conv(q + ".conv1", q + ".bn1", t, chans[i], k=3, act=RELU)
                    t = v.conv(q + ".conv2", q + ".bn2", u, chans[i], k=3, act=RELU, residual=t)
                br.append(t)
            if si == 2 and m == 0:
                # HighResolutionModule.forward overwrites the caller's list (:289-290); CA_PF therefore
                # receives these *pre-fusion* branch outputs for levels 1..3 (:501).
                kept = br
            outs = []
            for i in range(n_br if multi else 1):        # fuse (:225-280, :294-301)
                terms = []
                for j in range(n_br):
                    f = f"{sname}.{m}.fuse_layers.{i}.{j}"
                    if j == i:
                        terms.append((br[j], 0))
                    elif j > i:
                        t = v.conv(f + ".0", f + ".1", br[j], chans[i], k=1, act=NONE)
                        terms.append((t, j - i))         # nearest upsample by 2**(j-i)
                    else:
                        t = br[j]
                        for k in range(i - j):
                            last = k == i - j - 1
                            t = v.conv(f"{f}.{k}.0", f"{f}.{k}.1", t, chans[i] if last else chans[j],
                                       k=3, stride=2, act=NONE if last else RELU)
                        terms.append((t, 0))
                outs.append(v.fuse(terms, relu=True))
            xs = outs
        ys = xs
    return [ys[0], kept[1], kept[2], kept[3]]


# ------------------------------------------------------------------------------------------------------
# CPN-50
# ------------------------------------------------------------------------------------------------------
def _res_bottleneck(v, q, x, planes, stride, has_down):
    ident = x
    if has_down:
        ident = v.conv(q + ".downsample.0", q + ".downsample.1", x, planes * 4, k=1, stride=stride, act=NONE)
    t = v.conv(q + ".conv1", q + ".bn1", x, planes, k=1, act=RELU)
    t = v.conv(q + ".conv2", q + ".bn2", t, planes, k=3, stride=stride, act=RELU)
    return v.conv(q + ".conv3", q + ".bn3", t, planes * 4, k=1, act=RELU, residual=ident)


def walk_cpn(v, x, output_shape=(64, 48), num_class=17):
    """CPN50(output_shape, num_class) forward (network.py:16-22).  Returns 4 maps [N,64,48,256]."""
    # ResNet-50 (resnet.py:95-147)
    x = v.conv("resnet.conv1", "resnet.bn1", x, 64, k=7, stride=2, act=RELU)
    x = v.maxpool(x)
    feats = []
    inpl = 64
    for li, (planes, blocks, stride) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))):
        for b in range(blocks):
            s = stride if b == 0 else 1
            x = _res_bottleneck(v, f"resnet.layer{li + 1}.{b}", x, planes, s, has_down=(b == 0 and (s != 1 or inpl != planes * 4)))
            inpl = planes * 4
        feats.append(x)
    res_out = feats[::-1]                                # [x4, x3, x2, x1]

    # GlobalNet (globalNet.py:61-83); the `predict` heads are evaluated and dropped by the reference (:71):
    # their parameters are kept in the state_dict, their FLOPs are skipped.
    fms = []
    g = "global_net"
    for i in range(4):
        lat = v.conv(f"{g}.laterals.{i}.0", f"{g}.laterals.{i}.1", res_out[i], 256, k=1, act=RELU)
        if i == 0:
            feature = lat
        else:
            # upsamples[i-1] = Upsample(x2, bilinear, align_corners) -> 1x1 conv -> BN of the coarser feature
            # (:38-45); `laterals(x) + up` (:66) is emitted as that conv with the lateral as its residual.
            u = v.bilinear(fms[-1], fms[-1].H * 2, fms[-1].W * 2)
            feature = v.conv(f"{g}.upsamples.{i - 1}.1", f"{g}.upsamples.{i - 1}.2", u, 256, k=1, act=NONE, residual=lat)
        fms.append(feature)
        v.dead_conv(f"{g}.predict.{i}.0", 256, 256, 1)
        v.dead_bn(f"{g}.predict.{i}.1", 256)
        v.dead_conv(f"{g}.predict.{i}.3", 256, num_class, 3)
        v.dead_bn(f"{g}.predict.{i}.5", num_class)

    # RefineNet (refineNet.py:72-88): cascade[i] = (3-i) Bottleneck(256,128) + bilinear resize to output_shape
    outs = []
    for i in range(4):
        t = fms[i]
        for k in range(3 - i):
            q = f"refine_net.cascade.{i}.{k}"
            ident = v.conv(q + ".downsample.0", q + ".downsample.1", t, 256, k=1, act=NONE)
            u = v.conv(q + ".conv1", q + ".bn1", t, 128, k=1, act=RELU)
            u = v.conv(q + ".conv2", q + ".bn2", u, 128, k=3, act=RELU)
            t = v.conv(q + ".conv3", q + ".bn3", u, 256, k=1, act=RELU, residual=ident)
        outs.append(v.bilinear(t, output_shape[0], output_shape[1]))
    q = "refine_net.final_predict"
    v.dead_conv(q + ".0.conv1", 1024, 128, 1); v.dead_bn(q + ".0.bn1", 128)
    v.dead_conv(q + ".0.conv2", 128, 128, 3); v.dead_bn(q + ".0.bn2", 128)
    v.dead_conv(q + ".0.conv3", 128, 256, 1); v.dead_bn(q + ".0.bn3", 256)
    v.dead_conv(q + ".0.downsample.0", 1024, 256, 1); v.dead_bn(q + ".0.downsample.1", 256)
    v.dead_conv(q + ".1", 256, num_class, 3); v.dead_bn(q + ".2", num_class)
    return outs


def feature_dims(backbone: str, base_dim: int):
    """Channel list of the 4 maps as PoseTransformer expects them (pose_dformer.py:177-180)."""
    if backbone in ("hrnet_32", "hrnet_48"):
        return [base_dim, base_dim * 2, base_dim * 4, base_dim * 8]
    if backbone == "cpn":
        return [base_dim] * 4
    raise ValueError(f"unknown backbone {backbone!r}")
