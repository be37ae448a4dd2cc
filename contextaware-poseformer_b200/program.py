"""Forward programs: the flat op list libcapf_b200 executes for one (backbone, precision, B, H, W).

``ProgramBuilder`` is the visitor that arch.walk_* drive; ``build_forward_program`` appends the sampler and
lifter (PoseTransformer.forward, pose_dformer.py:210-241).  A ``Program`` is pure description (shapes, buffer
liveness, weight-packing recipes) and needs neither a GPU nor the shared library, so the CPU test-suite can
check it against the oracle with a reference interpreter (tests/interp.py).  ``Plan`` binds a program to
device memory + packed weights and runs it through the C ABI.

Precision policies
  fp32 : every tensor f32, CUDA-core kernels                        (parity mode, <=1e-5 of the oracle)
  fp16 : backbone activations/weights f16 (f32 accumulate), lifter token stream f32 with f16 GEMM operands
  bf16 : same with bfloat16
  bf16x3: fp32 storage everywhere, every conv / Linear on the tensor cores with SPLIT operands: x = hi + lo and w = Wh + Wl
         in bfloat16 (hi = bf16(v), lo = bf16(v - hi)), three products hi*Wh + lo*Wh + hi*Wl accumulated in fp32 -- operand
         precision 2^-16, the tensor-core mode that meets the reference's fp32 results to 1e-3 (observed ~1e-5)
"""
import os
import re
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch

from . import arch, lib

J = 17  # joints (pose_dformer.py:145)

_TORCH_DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16, "i32": torch.int32}
_ITEMSIZE = {"f32": 4, "f16": 2, "bf16": 2, "i32": 4}


@dataclass(eq=False)
class Buf:
    name: str
    shape: tuple
    dtype: str
    role: str = "act"            # act | input | output
    base: Optional["Buf"] = None  # views share storage with base at `offset` elements
    offset: int = 0

    @property
    def root(self):
        return self.base.root if self.base is not None else self

    @property
    def root_offset(self):
        return self.offset + (self.base.root_offset if self.base is not None else 0)

    @property
    def numel(self):
        return int(np.prod(self.shape))

    @property
    def nbytes(self):
        return self.numel * _ITEMSIZE[self.dtype]

    def view(self, offset, shape, name=None):
        return Buf(name or f"{self.name}[{offset}:]", tuple(shape), self.dtype, self.role, self, int(offset))


@dataclass(eq=False)
class WSlot:
    """A packed parameter blob; ``pack(state_dict) -> cpu tensor`` of `shape`/`dtype`."""
    name: str
    shape: tuple
    dtype: str
    pack: Callable


@dataclass(eq=False)
class Op:
    kind: int
    dtype_in: str
    dtype_out: str
    i: List[int]
    f: List[float]
    ins: list
    outs: list
    tag: str = ""
    flops: int = 0
    nbytes: int = 0        # algorithmic bytes of one launch: inputs + outputs (+ residual, weights), each counted once
    lane: int = 0          # execution lane (stream): HRNet branch chains run side by side, see assign_lanes()


@dataclass
class Program:
    backbone: str
    precision: str
    B: int
    H: int
    W: int
    ops: List[Op] = field(default_factory=list)
    inputs: dict = field(default_factory=dict)
    outputs: dict = field(default_factory=dict)
    feature_maps: list = field(default_factory=list)
    n_backbone_ops: int = 0

    def flops(self):
        return sum(o.flops for o in self.ops)


# ------------------------------------------------------------------------------------------------------
# weight packing (host side, float64 folding so the fold itself adds no error)
# ------------------------------------------------------------------------------------------------------
BN_EPS = 1e-5  # nn.BatchNorm2d default, used throughout the reference


def _fold(state, wkey, bnkey):
    w = state[wkey].detach().to("cpu", torch.float64)
    cout = w.shape[0]
    if bnkey is None:
        return w, torch.zeros(cout, dtype=torch.float64)
    g = state[bnkey + ".weight"].detach().to("cpu", torch.float64)
    b = state[bnkey + ".bias"].detach().to("cpu", torch.float64)
    m = state[bnkey + ".running_mean"].detach().to("cpu", torch.float64)
    v = state[bnkey + ".running_var"].detach().to("cpu", torch.float64)
    s = g / torch.sqrt(v + BN_EPS)
    return w * s.view(-1, 1, 1, 1), b - m * s


def conv_weight_packer(wkey, bnkey, dtype, layout):
    def pack(state):
        w, _ = _fold(state, wkey, bnkey)
        cout = w.shape[0]
        if layout == "kc":      # SIMT kernel: [KH*KW*Cin][Cout], k = (r*KW+s)*Cin+ci
            wp = w.permute(2, 3, 1, 0).reshape(-1, cout)
        else:                   # tcgen05 kernel: [Cout][KH*KW*Cin]  (K-major B operand)
            wp = w.permute(0, 2, 3, 1).reshape(cout, -1)
        return wp.contiguous().to(_TORCH_DT[dtype])
    return pack


def _split_bf16(w):
    """w (float64/32) -> (hi, lo) bfloat16 with hi + lo == w to ~2^-17."""
    w32 = w.to(torch.float32)
    hi = w32.to(torch.bfloat16)
    lo = (w32 - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def split_weight_packer(get_w):
    """B operand of a split-operand GEMM (CAPF_OP_CONV2D i[18] = 1): [Cout][taps * (Wh | Wh)][taps * Wl], matching the A
    walk of the kernel: pass 0 over the (hi | lo) planes of every tap, pass 1 over the hi plane of every tap."""
    def pack(state):
        w = get_w(state)                                    # [Cout, KH, KW, Cin]
        cout = w.shape[0]
        hi, lo = _split_bf16(w)
        p0 = torch.cat([hi, hi], dim=-1).reshape(cout, -1)  # per tap: Wh for the hi plane, Wh for the lo plane
        p1 = lo.reshape(cout, -1)
        return torch.cat([p0, p1], dim=1).contiguous()
    return pack


def conv_bias_packer(wkey, bnkey):
    def pack(state):
        return _fold(state, wkey, bnkey)[1].to(torch.float32)
    return pack


def linear_weight_packer(wkeys, dtype, layout):
    def pack(state):
        w = torch.cat([state[k].detach().to("cpu", torch.float32) for k in wkeys], dim=0)   # [out, in]
        wp = w.t() if layout == "kc" else w
        return wp.contiguous().to(_TORCH_DT[dtype])
    return pack


def vec_packer(keys):
    def pack(state):
        return torch.cat([state[k].detach().to("cpu", torch.float32).reshape(-1) for k in keys]).contiguous()
    return pack


# ------------------------------------------------------------------------------------------------------
# builder
# ------------------------------------------------------------------------------------------------------
class ProgramBuilder:
    def __init__(self, program: Program, shapes: dict, use_tc: bool = False):
        self.p = program
        self.shapes = shapes          # state_dict key -> shape (for channel counts)
        self.prefix = "backbone."
        prec = program.precision
        self.act_dt = {"fp32": "f32", "fp16": "f16", "bf16": "bf16", "bf16x3": "f32"}[prec]
        self.use_tc = use_tc and prec != "fp32"
        self.split = self.use_tc and prec == "bf16x3"
        self._split_cache = {}
        self._n = 0

    # -- helpers ---------------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype, role="act"):
        self._n += 1
        return Buf(f"{name}#{self._n}", tuple(int(s) for s in shape), dtype, role)

    def _emit(self, *a, **k):
        op = Op(*a, **k)
        self.p.ops.append(op)
        return op

    def _conv_impl(self, cin, cout, k, stride, dt_in):
        """Kernel family for a conv/linear.  tcgen05 needs 16-bit operands and TMA-friendly channel counts."""
        if self.split and dt_in == "f32":      # fp32 tensors, split into bf16 hi|lo planes in front of the GEMM
            return lib.IMPL_TCGEN05 if (cin % 16 == 0 and cout % 16 == 0 and stride <= 2 and k <= 7) else lib.IMPL_SIMT
        if not self.use_tc or dt_in == "f32":
            return lib.IMPL_SIMT
        if cin % 16 or cout % 16 or stride > 2 or k > 7:
            return lib.IMPL_SIMT
        return lib.IMPL_TCGEN05

    # -- visitor API used by arch.walk_* -----------------------------------------------------------------
    def conv(self, cname, bname, x, cout, k=1, stride=1, act=arch.NONE, residual=None):
        B = self.p.B
        wkey = self.prefix + cname + ".weight"
        bnkey = self.prefix + bname
        pad = k // 2
        Ho = (x.H + 2 * pad - k) // stride + 1
        Wo = (x.W + 2 * pad - k) // stride + 1
        src: Buf = x.ref
        dt_in = src.dtype
        dt_out = self.act_dt
        impl = self._conv_impl(x.C, cout, k, stride, dt_in)
        wdt = "f32" if dt_in == "f32" else dt_in
        layout = "ck" if impl == lib.IMPL_TCGEN05 else "kc"
        K = k * k * x.C
        b = WSlot(f"b:{cname}", (cout,), "f32", conv_bias_packer(wkey, bnkey))
        out = self._buf(cname, (B, Ho, Wo, cout), dt_out)
        acode = {arch.NONE: lib.ACT_NONE, arch.RELU: lib.ACT_RELU, arch.GELU: lib.ACT_GELU}[act]
        extra = []
        if self.split and impl == lib.IMPL_TCGEN05 and dt_in == "f32":
            # backbone tensors are written once (every conv gets a fresh buffer): share the planes between the consumers
            key = id(src)
            if key not in self._split_cache:
                self._split_cache[key] = self._split_planes(src, B * x.H * x.W, x.C, (B, x.H, x.W, 2 * x.C), self.prefix + cname)
            src = self._split_cache[key]
            dt_in = wdt = "bf16"
            w = WSlot(f"w:{cname}", (cout, 3 * K), "bf16",
                      split_weight_packer(lambda st: _fold(st, wkey, bnkey)[0].permute(0, 2, 3, 1)))
            extra = [1, 0, 0, 0, 0, 1]              # i[13] = per-tap kernel, i[18] = split operands
        else:
            w = WSlot(f"w:{cname}", (K, cout) if layout == "kc" else (cout, K), wdt, conv_weight_packer(wkey, bnkey, wdt, layout))
        self._emit(lib.OP_CONV2D, dt_in, dt_out,
                   [B, x.H, x.W, x.C, cout, k, k, stride, pad, Ho, Wo, acode, impl] + extra, [],
                   [src, w, b, residual.ref if residual is not None else None], [out],
                   tag=self.prefix + cname, flops=2 * B * Ho * Wo * cout * K)
        self.p.ops[-1].nbytes = src.nbytes + out.nbytes + (out.nbytes if residual is not None else 0) + K * cout * _ITEMSIZE[wdt] + 4 * cout
        return arch.T(Ho, Wo, cout, out)

    def _split_planes(self, src: Buf, rows, C, shape, tag):
        """CAPF_OP_CAST in split mode: fp32 [rows][C] -> bf16 [rows][hi (C) | lo (C)], the A operand of a split GEMM."""
        out = self._buf("split", shape, "bf16")
        n = rows * C
        self._emit(lib.OP_CAST, "f32", "bf16", [n & 0x7fffffff, n >> 31, C], [], [src], [out], tag=tag + ".split")
        self.p.ops[-1].nbytes = 8 * n
        return out

    def fuse(self, terms, relu=True):
        B = self.p.B
        t0 = terms[0][0]
        assert terms[0][1] == 0 or True
        # output resolution = resolution of the shift-0 term
        ref_t = next(t for t, s in terms if s == 0)
        H, W, Cc = ref_t.H, ref_t.W, ref_t.C
        out = self._buf("fuse", (B, H, W, Cc), self.act_dt)
        shifts = [s for _, s in terms] + [0] * (4 - len(terms))
        self._emit(lib.OP_FUSE_SUM, self.act_dt, self.act_dt, [B, H, W, Cc, len(terms)] + shifts + [1 if relu else 0], [],
                   [t.ref for t, _ in terms], [out], tag="fuse")
        return arch.T(H, W, Cc, out)

    def maxpool(self, x):
        B = self.p.B
        Ho, Wo = (x.H + 2 - 3) // 2 + 1, (x.W + 2 - 3) // 2 + 1
        out = self._buf("maxpool", (B, Ho, Wo, x.C), x.ref.dtype)
        self._emit(lib.OP_MAXPOOL, x.ref.dtype, x.ref.dtype, [B, x.H, x.W, x.C, Ho, Wo], [], [x.ref], [out], tag="maxpool")
        return arch.T(Ho, Wo, x.C, out)

    def bilinear(self, x, Ho, Wo):
        B = self.p.B
        out = self._buf("bilinear", (B, Ho, Wo, x.C), x.ref.dtype)
        self._emit(lib.OP_BILINEAR, x.ref.dtype, x.ref.dtype, [B, x.H, x.W, x.C, Ho, Wo], [], [x.ref], [out], tag="bilinear")
        return arch.T(Ho, Wo, x.C, out)

    def dead_conv(self, *a):
        pass

    def dead_bn(self, *a):
        pass

    # -- token-matrix ops for the lifter -------------------------------------------------------------------
    def linear(self, x: Buf, rows, cin, wkeys, bkeys, cout, act=arch.NONE, residual: Buf = None, out: Buf = None,
               out_dtype=None, tag=""):
        dt_in = x.dtype
        dt_out = out_dtype or (out.dtype if out is not None else self.act_dt)
        impl = self._conv_impl(cin, cout, 1, 1, dt_in)
        wdt = "f32" if dt_in == "f32" else dt_in
        layout = "ck" if impl == lib.IMPL_TCGEN05 else "kc"
        b = WSlot(f"b:{tag}", (cout,), "f32", vec_packer(bkeys)) if bkeys else None
        if out is None:
            out = self._buf(tag, (rows, cout), dt_out)
        acode = {arch.NONE: lib.ACT_NONE, arch.RELU: lib.ACT_RELU, arch.GELU: lib.ACT_GELU}[act]
        extra = []
        if self.split and impl == lib.IMPL_TCGEN05 and dt_in == "f32":
            x = self._split_planes(x, rows, cin, (rows, 2 * cin), tag)
            dt_in = wdt = "bf16"
            w = WSlot(f"w:{tag}", (cout, 3 * cin), "bf16", split_weight_packer(
                lambda st: torch.cat([st[k].detach().to("cpu", torch.float32) for k in wkeys], dim=0).view(cout, 1, 1, cin)))
            extra = [1, 0, 0, 0, 0, 1]
        else:
            w = WSlot(f"w:{tag}", (cin, cout) if layout == "kc" else (cout, cin), wdt, linear_weight_packer(wkeys, wdt, layout))
        self._emit(lib.OP_CONV2D, dt_in, dt_out, [rows, 1, 1, cin, cout, 1, 1, 1, 0, 1, 1, acode, impl] + extra, [],
                   [x, w, b, residual], [out], tag=tag, flops=2 * rows * cin * cout)
        self.p.ops[-1].nbytes = rows * cin * _ITEMSIZE[dt_in] + rows * cout * _ITEMSIZE[dt_out] * (2 if residual is not None else 1) + \
            cin * cout * _ITEMSIZE[wdt] + 4 * cout
        return out

    def layernorm(self, x: Buf, rows, D, prefix, eps, out_dtype, x0: Buf = None, period=0, tag=""):
        g = WSlot(f"g:{tag}", (D,), "f32", vec_packer([prefix + ".weight"]))
        b = WSlot(f"b:{tag}", (D,), "f32", vec_packer([prefix + ".bias"]))
        out = self._buf(tag, (rows, D), out_dtype)
        self._emit(lib.OP_LAYERNORM, "f32", out_dtype, [rows, D, period], [eps], [x, g, b, x0], [out], tag=tag)
        return out


# ------------------------------------------------------------------------------------------------------
# whole-forward program
# ------------------------------------------------------------------------------------------------------
def build_forward_program(backbone: str, bb_cfg, pf_cfg, shapes: dict, B: int, H: int, W: int, precision: str = "fp32",
                          use_tc: bool = False, debug_records: bool = False, backbone_only: bool = False,
                          variant: str = "h36m") -> Program:
    """CA_PF.forward (conpose.py:30-42) after the crop normalisation, as a Program.

    inputs : images [B,H,W,3] f32 (NHWC as the caller passes them -- no permute is needed, conpose.py:32),
             kp2d [B*17,2] f32, ref [B*17,2] f32 (already normalised, conpose.py:34-35)
    output : out [B*17,3] f32  (== [B,1,17,3])
    """
    if precision not in ("fp32", "fp16", "bf16", "bf16x3"):
        raise ValueError(f"precision {precision!r}")
    prog = Program(backbone, precision, B, H, W)
    pb = ProgramBuilder(prog, shapes, use_tc)
    images = Buf("images", (B, H, W, 3), "f32", "input")
    prog.inputs["images"] = images
    x = arch.T(H, W, 3, images)
    if backbone in ("hrnet_32", "hrnet_48"):
        maps = arch.walk_hrnet(pb, x, bb_cfg)
    elif backbone == "cpn":
        maps = arch.walk_cpn(pb, x)
    else:
        raise ValueError(backbone)
    prog.feature_maps = [m.ref for m in maps]
    if pb.use_tc:
        fuse_basic_blocks(prog)
        fuse_downsample(prog)
        fuse_expand_reduce(prog)
        fuse_siblings(prog)
        chunk_prefix(prog)
    prog.n_backbone_ops = len(prog.ops)
    if backbone_only:
        for k, m in enumerate(maps):
            m.ref.role = "output"
            prog.outputs[f"map{k}"] = m.ref
        assign_lanes(prog, _lanes_enabled())
        return prog

    if variant not in ("h36m", "mpi"):
        raise ValueError(f"variant {variant!r}")
    D = int(pf_cfg["embed_dim_ratio"])
    levels = int(pf_cfg["levels"])
    # H36M tree: `levels` doubles as the block depth (pose_dformer.py:169); the MPI-INF-3DHP tree has its own `depth` and
    # no DeformableBlocks (ContextPose_mpi/model/pose_dformer.py:199,211-222)
    depth = levels if variant == "h36m" else int(pf_cfg["depth"])
    n_context = levels if variant == "h36m" else 0
    if levels != 4:
        raise NotImplementedError("the sampler kernels are specialised for the reference's 4 feature levels")
    dims = arch.feature_dims(backbone, int(pf_cfg["base_dim"]))
    for l in range(levels):
        assert maps[l].C == dims[l], (maps[l].C, dims[l])
    R = B * J
    S = levels + 1
    E = D * S
    adt = pb.act_dt                     # GEMM operand dtype in the lifter
    vn = "volume_net."
    kp2d = Buf("kp2d", (R, 2), "f32", "input")
    ref = Buf("ref", (R, 2), "f32", "input")
    prog.inputs["kp2d"], prog.inputs["ref"] = kp2d, ref

    # ---- token stream X[slab][b*17+j][D] f32, level-major ------------------------------------------------
    X = pb._buf("X", (S, R, D), "f32")
    pb._emit(lib.OP_EMBED_COORD, "f32", "f32", [B, J, D, S], [],
             [kp2d, WSlot("w:coord_embed", (D, 2), "f32", vec_packer([vn + "coord_embed.weight"])),
              WSlot("b:coord_embed", (D,), "f32", vec_packer([vn + "coord_embed.bias"])),
              WSlot("pos", (S, J, D), "f32", vec_packer([vn + "Spatial_pos_embed"]))], [X], tag="embed_coord")

    map_geo = []
    for l in range(levels):
        map_geo += [maps[l].H, maps[l].W, maps[l].C]
    map_geo += [0] * (12 - len(map_geo))

    # ---- (a7) reference-point gather + feat_embed ---------------------------------------------------------
    offs = [0]
    for l in range(levels):
        offs.append(offs[-1] + R * dims[l])
    samp = pb._buf("ref_sampled", (offs[-1],), adt)
    rec0 = pb._buf("ref_corners", (levels, R, 8), "i32") if debug_records else None
    if rec0 is not None:
        rec0.role = "output"
        prog.outputs["ref_corners"] = rec0
    pb._emit(lib.OP_REF_SAMPLE, maps[0].ref.dtype, adt, [B, J, levels] + map_geo + offs[:4], [],
             [ref] + [m.ref for m in maps], [samp, rec0], tag="ref_sample")
    for l in range(levels):
        slab = X.view((1 + l) * R * D, (R, D), f"X[{1 + l}]")
        pb.linear(samp.view(offs[l], (R, dims[l])), R, dims[l], [f"{vn}feat_embed.{l}.weight"], [f"{vn}feat_embed.{l}.bias"],
                  D, residual=slab, out=slab, tag=f"{vn}feat_embed.{l}")

    # ---- (a8) 4 x DeformableBlock -------------------------------------------------------------------------
    X0 = X.view(0, (R, D), "X[0]")
    Xl = X.view(R * D, (levels * R, D), "X[1:]")
    goffs = [0]
    for l in range(levels):
        goffs.append(goffs[-1] + R * 4 * dims[l])
    for i in range(n_context):
        q = f"{vn}context_blocks.{i}"
        t = pb.layernorm(Xl, levels * R, D, q + ".norm1", 1e-5, adt, x0=X0, period=R, tag=q + ".norm1")
        ow = pb.linear(t, levels * R, D, [q + ".attention_weights.weight", q + ".sampling_offsets.weight"],
                       [q + ".attention_weights.bias", q + ".sampling_offsets.bias"], 48, out_dtype="f32", tag=q + ".ow")
        g = pb._buf("deform_sampled", (goffs[-1],), adt)
        rec = pb._buf("deform_corners", (levels, R, 16, 8), "i32") if (debug_records and i == 0) else None
        if rec is not None:
            rec.role = "output"
            prog.outputs["deform_corners"] = rec
        pb._emit(lib.OP_DEFORM_SAMPLE, maps[0].ref.dtype, adt, [B, J, levels] + map_geo + goffs[:4], [],
                 [ref] + [m.ref for m in maps] + [ow], [g, rec], tag=q + ".sample")
        for l in range(levels):
            slab4 = X.view((1 + l) * R * D, (R * 4, D // 4), f"X[{1 + l}] as heads")
            pb.linear(g.view(goffs[l], (R * 4, dims[l])), R * 4, dims[l], [f"{q}.embed_proj.{l}.weight"],
                      [f"{q}.embed_proj.{l}.bias"], D // 4, residual=slab4, out=slab4, tag=f"{q}.embed_proj.{l}")
        t = pb.layernorm(Xl, levels * R, D, q + ".norm2", 1e-5, adt, tag=q + ".norm2")
        hdn = pb.linear(t, levels * R, D, [q + ".mlp.fc1.weight"], [q + ".mlp.fc1.bias"], 2 * D, act=arch.GELU, tag=q + ".mlp.fc1")
        pb.linear(hdn, levels * R, 2 * D, [q + ".mlp.fc2.weight"], [q + ".mlp.fc2.bias"], D, residual=Xl, out=Xl, tag=q + ".mlp.fc2")

    # ---- (a9) transformer blocks ---------------------------------------------------------------------------
    def block(q, x: Buf, rows, dim, heads, groups, seq, tok_stride, grp_stride):
        t = pb.layernorm(x, rows, dim, q + ".norm1", 1e-6, adt, tag=q + ".norm1")
        qkv = pb.linear(t, rows, dim, [q + ".attn.qkv.weight"], [q + ".attn.qkv.bias"], 3 * dim, tag=q + ".attn.qkv")
        att = pb._buf(q + ".attn", (rows, dim), adt)
        hd = dim // heads
        pb._emit(lib.OP_ATTENTION, adt, adt, [groups, seq, heads, hd, tok_stride, grp_stride], [float(hd) ** -0.5],
                 [qkv], [att], tag=q + ".attn", flops=4 * groups * heads * seq * seq * hd)
        pb.linear(att, rows, dim, [q + ".attn.proj.weight"], [q + ".attn.proj.bias"], dim, residual=x, out=x, tag=q + ".attn.proj")
        t = pb.layernorm(x, rows, dim, q + ".norm2", 1e-6, adt, tag=q + ".norm2")
        hdn = pb.linear(t, rows, dim, [q + ".mlp.fc1.weight"], [q + ".mlp.fc1.bias"], 2 * dim, act=arch.GELU, tag=q + ".mlp.fc1")
        pb.linear(hdn, rows, 2 * dim, [q + ".mlp.fc2.weight"], [q + ".mlp.fc2.bias"], dim, residual=x, out=x, tag=q + ".mlp.fc2")

    Xall = X.view(0, (S * R, D), "X[:]")
    for i in range(depth):     # res_blocks: attention over the S level-tokens of one joint (:231-234)
        block(f"{vn}res_blocks.{i}", Xall, S * R, D, 8, R, S, R, 1)
    Y = pb._buf("Y", (R, E), "f32")
    pb._emit(lib.OP_LEVELS_TO_JOINT, "f32", "f32", [R, S, D], [], [X], [Y], tag="levels_to_joint")
    for i in range(depth):     # joint_blocks: attention over the 17 joints of a frame (:235-238)
        block(f"{vn}joint_blocks.{i}", Y, R, E, 8, B, J, 1, J)

    # ---- (a10) head ----------------------------------------------------------------------------------------
    out = Buf("out", (R, 3), "f32", "output")
    pb._emit(lib.OP_LAYERNORM, "f32", "f32", [R, E, 0, 3], [1e-5],
             [Y, WSlot("g:head.0", (E,), "f32", vec_packer([vn + "head.0.weight"])),
              WSlot("b:head.0", (E,), "f32", vec_packer([vn + "head.0.bias"])), None,
              WSlot("w:head.1", (3, E), "f32", vec_packer([vn + "head.1.weight"])),
              WSlot("b:head.1", (3,), "f32", vec_packer([vn + "head.1.bias"]))], [out],
             tag=vn + "head", flops=2 * R * E * 3)
    prog.outputs["out"] = out
    if pb.use_tc:
        fuse_mlp(prog, prog.n_backbone_ops)
    assign_lanes(prog, _lanes_enabled())
    return prog


def _lanes_enabled():
    return os.environ.get("CAPF_STREAMS", "1") != "0"


# ------------------------------------------------------------------------------------------------------
# peephole: HRNet BasicBlocks of the 32-channel branch -> one fused op (csrc/capf_tc_block.cu)
# ------------------------------------------------------------------------------------------------------
FUSED_BLOCK_CHANNELS = 32
FUSED_BLOCK_MAX_W = 128        # wider bands (input + intermediate, double buffered) do not fit one SM's shared memory
# 64-channel blocks run on CTA pairs (csrc/capf_tc_block64.cu: each CTA keeps half of the 2 x 72 KB of weights); CAPF_FUSE_BLOCKS64=0
# keeps the two-kernel form for them
FUSED_BLOCK_MAX_W_BY_C = {32: FUSED_BLOCK_MAX_W, 64: 64}


def _same_buf(a, b):
    return isinstance(a, Buf) and isinstance(b, Buf) and a.root is b.root and a.root_offset == b.root_offset and a.shape == b.shape


def fuse_basic_blocks(prog: Program):
    """conv3x3-BN-ReLU -> conv3x3-BN-(+x)-ReLU pairs (pose_hrnet.py:79-95) with 32 or 64 channels, 16-bit tensors and tcgen05
    kernels become one CAPF_OP_BASICBLOCK: the intermediate tensor disappears from the program (and from HBM).  Returns the
    number of fused pairs.  CAPF_FUSE_BLOCKS=0 keeps the two-kernel form."""
    if os.environ.get("CAPF_FUSE_BLOCKS", "1") == "0":
        return 0
    max_w = dict(FUSED_BLOCK_MAX_W_BY_C)
    if os.environ.get("CAPF_FUSE_BLOCKS64", "1") == "0":
        max_w.pop(64)
    readers = {}
    for op in prog.ops:
        for b in op.ins:
            if isinstance(b, Buf):
                readers[b.root] = readers.get(b.root, 0) + 1
    out, k, fused = [], 0, 0
    ops = prog.ops
    while k < len(ops):
        a = ops[k]
        b = ops[k + 1] if k + 1 < len(ops) else None
        ok = (b is not None and a.kind == lib.OP_CONV2D and b.kind == lib.OP_CONV2D
              and a.dtype_in == a.dtype_out == b.dtype_in == b.dtype_out and a.dtype_in in ("f16", "bf16")
              and a.i[12] == lib.IMPL_TCGEN05 and b.i[12] == lib.IMPL_TCGEN05
              and a.i[3] == a.i[4] == b.i[3] == b.i[4] and a.i[3] in max_w
              and a.i[5:9] == [3, 3, 1, 1] and b.i[5:9] == [3, 3, 1, 1] and a.i[:3] == b.i[:3] and a.i[2] <= max_w[a.i[3]]
              and a.i[11] == lib.ACT_RELU and b.i[11] == lib.ACT_RELU
              and a.ins[3] is None and _same_buf(b.ins[0], a.outs[0]) and _same_buf(b.ins[3], a.ins[0])
              and readers.get(a.outs[0].root, 0) == 1 and a.outs[0].role == "act")
        if ok:
            x, y = a.ins[0], b.outs[0]
            wbytes = sum(int(np.prod(w.shape)) * _ITEMSIZE[w.dtype] for w in (a.ins[1], b.ins[1]))
            out.append(Op(lib.OP_BASICBLOCK, a.dtype_in, a.dtype_out, [a.i[0], a.i[1], a.i[2], a.i[3]], [],
                          [x, a.ins[1], a.ins[2], b.ins[1], b.ins[2]], [y], tag=b.tag.rsplit(".", 1)[0] + ".block",
                          flops=a.flops + b.flops, nbytes=x.nbytes + y.nbytes + wbytes))
            fused += 1
            k += 2
        else:
            out.append(a)
            k += 1
    prog.ops[:] = out
    return fused


# ------------------------------------------------------------------------------------------------------
# peephole: Bottleneck conv3 + downsample conv -> one GEMM over the concatenated inputs
# ------------------------------------------------------------------------------------------------------
def fuse_downsample(prog: Program):
    """A Bottleneck whose shortcut is a 1x1 conv + BN (pose_hrnet.py:116-136 with `downsample`, :421-427; networks/resnet.py,
    networks/refineNet.py:17-21) computes relu(bn3(conv3(t)) + bn_d(conv_d(x))): two per-pixel GEMMs into the same output.
    As separate ops the shortcut tensor (256 channels: 537 MB at bs = 256 for HRNet's layer1.0) is written by one kernel and read
    back as the residual of the other.  Here both become ONE CAPF_OP_CONV2D with two A operands (i[19] = Cin2, in[5] = x):
    out = act([t | x] . [W3 | Wd]^T + b3 + bd), accumulated in fp32 -- the shortcut never exists in memory.  Stride-1 shortcuts
    only (a strided shortcut is not a plain row-major matrix).  CAPF_FUSE_DOWNSAMPLE=0 keeps the two-op form."""
    if os.environ.get("CAPF_FUSE_DOWNSAMPLE", "1") == "0":
        return 0
    ops = prog.ops
    readers = {}
    for op in ops:
        for b in op.ins:
            if isinstance(b, Buf):
                readers[b.root] = readers.get(b.root, 0) + 1

    def plain_1x1(op):
        return (op.kind == lib.OP_CONV2D and op.i[12] == lib.IMPL_TCGEN05 and op.i[5:9] == [1, 1, 1, 0] and len(op.i) <= 13
                and op.dtype_in in ("f16", "bf16") and op.i[3] % 16 == 0)

    producer = {}
    for k, op in enumerate(ops):
        for b in op.outs:
            if isinstance(b, Buf):
                producer[b.root] = k
    drop, fused = set(), 0
    for k, b in enumerate(ops):
        if not plain_1x1(b) or not isinstance(b.ins[3], Buf):
            continue
        ka = producer.get(b.ins[3].root)
        if ka is None or ka in drop:
            continue
        a = ops[ka]
        if (not plain_1x1(a) or a.ins[3] is not None or a.i[11] != lib.ACT_NONE or a.i[0:3] != b.i[0:3] or a.i[4] != b.i[4]
                or a.dtype_in != b.dtype_in or a.dtype_out != b.dtype_out or readers.get(a.outs[0].root, 0) != 1
                or a.outs[0].role != "act" or not _same_buf(a.outs[0], b.ins[3])):
            continue
        wa, wb, ba, bb = a.ins[1], b.ins[1], a.ins[2], b.ins[2]
        cin1, cin2, cout = b.i[3], a.i[3], b.i[4]
        w = WSlot(wb.name + "+" + wa.name, (cout, cin1 + cin2), wb.dtype,
                  (lambda pb_, pa_: lambda st: torch.cat([pb_(st), pa_(st)], dim=1).contiguous())(wb.pack, wa.pack))
        bias = WSlot(bb.name + "+" + ba.name, (cout,), "f32", (lambda pb_, pa_: lambda st: (pb_(st) + pa_(st)).contiguous())(bb.pack, ba.pack))
        i = list(b.i) + [0] * (20 - len(b.i))
        i[19] = cin2
        ops[k] = Op(lib.OP_CONV2D, b.dtype_in, b.dtype_out, i, list(b.f), [b.ins[0], w, bias, None, None, a.ins[0]], list(b.outs),
                    tag=b.tag + "+" + a.tag.rsplit(".", 2)[-2] + "." + a.tag.rsplit(".", 1)[-1], flops=a.flops + b.flops,
                    nbytes=b.ins[0].nbytes + a.ins[0].nbytes + b.outs[0].nbytes + (cin1 + cin2) * cout * _ITEMSIZE[wb.dtype] + 4 * cout)
        drop.add(ka)
        fused += 1
    prog.ops[:] = [op for k, op in enumerate(ops) if k not in drop]
    return fused


# ------------------------------------------------------------------------------------------------------
# peephole: Bottleneck conv3 (+ residual) followed by conv1 of the next Bottleneck -> one kernel
# ------------------------------------------------------------------------------------------------------
def fuse_expand_reduce(prog: Program):
    """Inside a stack of Bottlenecks (pose_hrnet.py:421-427 layer1; networks/resnet.py layer1) block i ends with
    y = relu(bn3(conv3(t)) + x) -- 1x1, 64 -> 256 -- and block i + 1 starts with relu(bn1(conv1(y))) -- 1x1, 256 -> 64: two per-pixel
    GEMMs.  As separate ops the second one is nothing but a read of the 256-channel tensor the first one just wrote (537 MB at
    bs = 256).  Both become ONE CAPF_OP_EXPAND_REDUCE (csrc/capf_tc_chain.cu): y is written once (it is block i + 1's residual) and
    never read back for conv1.  Bit-identical results.  CAPF_FUSE_CHAIN=0 keeps the two-op form."""
    if os.environ.get("CAPF_FUSE_CHAIN", "1") == "0" or int(os.environ.get("CAPF_CHUNK", "0") or 0) > 0:
        return 0
    ops = prog.ops

    def plain_1x1(op, cin, cout, act):
        return (op.kind == lib.OP_CONV2D and op.i[12] == lib.IMPL_TCGEN05 and op.i[5:9] == [1, 1, 1, 0] and op.i[3] == cin and op.i[4] == cout
                and op.i[11] == act and not any(op.i[13:]) and all(x is None for x in op.ins[4:]))

    out, k, fused = [], 0, 0
    while k < len(ops):
        a = ops[k]
        b = ops[k + 1] if k + 1 < len(ops) else None
        ok = (b is not None and plain_1x1(a, 64, 256, lib.ACT_RELU) and plain_1x1(b, 256, 64, lib.ACT_RELU)
              and a.dtype_in == a.dtype_out == b.dtype_in == b.dtype_out and a.dtype_in in ("f16", "bf16")
              and isinstance(a.ins[3], Buf) and b.ins[3] is None and _same_buf(b.ins[0], a.outs[0]) and a.i[0:3] == b.i[0:3]
              and a.i[9:11] == a.i[1:3] and getattr(a, "lane", 0) == getattr(b, "lane", 0))
        if ok:
            rows = a.i[0] * a.i[1] * a.i[2]
            wbytes = sum(int(np.prod(w.shape)) * _ITEMSIZE[w.dtype] for w in (a.ins[1], b.ins[1]))
            op = Op(lib.OP_EXPAND_REDUCE, a.dtype_in, a.dtype_out, [rows, 64, 256, 64], [],
                    [a.ins[0], a.ins[1], a.ins[2], a.ins[3], b.ins[1], b.ins[2]], [a.outs[0], b.outs[0]],
                    tag=a.tag + "+" + ".".join(b.tag.rsplit(".", 2)[-2:]), flops=a.flops + b.flops,
                    nbytes=a.ins[0].nbytes + a.ins[3].nbytes + a.outs[0].nbytes + b.outs[0].nbytes + wbytes)
            out.append(op)
            fused += 1
            k += 2
        else:
            out.append(a)
            k += 1
    prog.ops[:] = out
    return fused


# ------------------------------------------------------------------------------------------------------
# peephole: Mlp of a 128-wide transformer block (fc1 + GELU -> fc2 + residual) -> one kernel
# ------------------------------------------------------------------------------------------------------
def fuse_mlp(prog: Program, first: int = 0):
    """The Mlp of the DeformableBlocks and of the res blocks (pose_dformer.py:25-31, :138-141, :78) is x + fc2(gelu(fc1(norm2(x)))) with
    128 -> 256 -> 128 features: two GEMM launches whose hidden tensor makes an HBM round trip and which each pay the fixed cost of a
    tcgen05 kernel for ~1 us of tensor work.  Both become ONE CAPF_OP_MLP (csrc/capf_tc_mlp.cu): the hidden tile stays in shared
    memory, the fp32 token stream is updated in place.  Bit-identical results.  CAPF_FUSE_MLP=0 keeps the two-op form."""
    if os.environ.get("CAPF_FUSE_MLP", "1") == "0":
        return 0
    ops = prog.ops
    readers = {}
    for op in ops:
        for b in op.ins:
            if isinstance(b, Buf):
                readers[b.root] = readers.get(b.root, 0) + 1

    def rows_linear(op, cin, cout, act):
        return (op.kind == lib.OP_CONV2D and op.i[12] == lib.IMPL_TCGEN05 and op.i[1:3] == [1, 1] and op.i[5:9] == [1, 1, 1, 0] and op.i[3] == cin
                and op.i[4] == cout and op.i[11] == act and not any(op.i[13:]) and all(x is None for x in op.ins[4:]) and op.ins[2] is not None)

    out, k, fused = list(ops[:first]), first, 0
    while k < len(ops):
        a = ops[k]
        b = ops[k + 1] if k + 1 < len(ops) else None
        ok = (b is not None and rows_linear(a, 128, 256, lib.ACT_GELU) and rows_linear(b, 256, 128, lib.ACT_NONE)
              and a.dtype_in in ("f16", "bf16") and a.dtype_out == a.dtype_in and b.dtype_in == a.dtype_in and b.dtype_out == "f32"
              and a.ins[3] is None and isinstance(b.ins[3], Buf) and _same_buf(b.ins[3], b.outs[0]) and _same_buf(b.ins[0], a.outs[0])
              and a.i[0] == b.i[0] and readers.get(a.outs[0].root, 0) == 1 and a.outs[0].role == "act")
        if ok:
            rows = a.i[0]
            wbytes = sum(int(np.prod(w.shape)) * _ITEMSIZE[w.dtype] for w in (a.ins[1], b.ins[1]))
            out.append(Op(lib.OP_MLP, a.dtype_in, "f32", [rows, 128, 256, 128], [],
                          [a.ins[0], a.ins[1], a.ins[2], b.ins[3], b.ins[1], b.ins[2]], [b.outs[0]],
                          tag=a.tag.rsplit(".", 1)[0], flops=a.flops + b.flops,
                          nbytes=a.ins[0].nbytes + 2 * b.outs[0].nbytes + wbytes))
            fused += 1
            k += 2
        else:
            out.append(a)
            k += 1
    prog.ops[:] = out
    return fused


# ------------------------------------------------------------------------------------------------------
# horizontal fusion: sibling convolutions of a fuse layer that read the same branch -> one GEMM, one tensor per sibling
# ------------------------------------------------------------------------------------------------------
SIBLING_MAX_COUT = 256      # one column tile of the per-tap kernel (N <= 256)


def fuse_siblings(prog: Program):
    """The fuse layers of a HighResolutionModule (pose_hrnet.py:235-277) start several convolutions from the same branch output:
    branch 0 of a 4-branch module feeds three 3x3 / stride-2 convs (32 -> 64 towards output 1, 32 -> 32 + ReLU as the first step of
    the chains towards outputs 2 and 3), branch 3 feeds three 1x1 convs (256 -> 32 / 64 / 128), and so on.  As separate launches
    each of them fetches the same input again (nine times per launch for the 3x3 ones: the per-tap kernel is bound by that feed).
    Siblings with identical geometry become ONE CAPF_OP_CONV2D over the Cout-concatenated weights with OUTPUT SEGMENTS (i[20..23],
    per-segment ReLU mask i[14]): every sibling still gets its own dense tensor, so no consumer changes.  A column of the GEMM is
    computed exactly as before (same K order), results are bit-identical.  Only convs that run on the per-tap kernel anyway (1x1, or
    stride 2) are merged; 3x3 / stride-1 convs keep their halo kernels.  CAPF_FUSE_SIBLINGS=0 keeps the separate launches."""
    if os.environ.get("CAPF_FUSE_SIBLINGS", "1") == "0":
        return 0
    ops = prog.ops

    def key(op):
        if (op.kind != lib.OP_CONV2D or len(op.i) > 13 or op.i[12] != lib.IMPL_TCGEN05 or op.dtype_in not in ("f16", "bf16")
                or op.dtype_out != op.dtype_in or len(op.ins) != 4 or op.ins[3] is not None or op.i[11] not in (lib.ACT_NONE, lib.ACT_RELU)
                or not isinstance(op.ins[0], Buf) or len(op.outs) != 1 or op.outs[0].role != "act" or op.i[4] % 16
                or (op.i[5] == 3 and op.i[7] == 1)):
            return None
        src = op.ins[0]
        return (id(src.root), src.root_offset, tuple(src.shape), tuple(op.i[0:4]), tuple(op.i[5:11]), op.dtype_in)

    groups = {}
    for k, op in enumerate(ops):
        kk = key(op)
        if kk is not None:
            groups.setdefault(kk, []).append(k)
    replace, drop, fused = {}, set(), 0
    for members in groups.values():
        batch = []
        for k in members + [None]:
            if k is not None and len(batch) < 4 and sum(ops[m].i[4] for m in batch) + ops[k].i[4] <= SIBLING_MAX_COUT:
                batch.append(k)
                continue
            if len(batch) > 1:
                first = ops[batch[0]]
                sib = [ops[m] for m in batch]
                couts = [o.i[4] for o in sib]
                wdt = first.ins[1].dtype
                K = first.i[3] * first.i[5] * first.i[6]
                w = WSlot("+".join(o.ins[1].name for o in sib), (sum(couts), K), wdt,
                          (lambda packs: lambda st: torch.cat([pk(st) for pk in packs], dim=0).contiguous())([o.ins[1].pack for o in sib]))
                bias = WSlot("+".join(o.ins[2].name for o in sib), (sum(couts),), "f32",
                             (lambda packs: lambda st: torch.cat([pk(st) for pk in packs]).contiguous())([o.ins[2].pack for o in sib]))
                relu = any(o.i[11] == lib.ACT_RELU for o in sib)
                i = list(first.i) + [0] * (24 - len(first.i))
                i[4] = sum(couts)
                i[11] = lib.ACT_RELU if relu else lib.ACT_NONE
                i[14] = sum(1 << n for n, o in enumerate(sib) if relu and o.i[11] != lib.ACT_RELU)
                i[20] = len(sib)
                for n, c in enumerate(couts[:-1]):
                    i[21 + n] = c
                src = first.ins[0]
                replace[batch[0]] = Op(lib.OP_CONV2D, first.dtype_in, first.dtype_out, i, [], [src, w, bias, None], [o.outs[0] for o in sib],
                                       tag=first.tag + "".join("+" + ".".join(o.tag.split(".")[-4:]) for o in sib[1:]),
                                       flops=sum(o.flops for o in sib),
                                       nbytes=src.nbytes + sum(o.outs[0].nbytes for o in sib) + sum(couts) * K * _ITEMSIZE[wdt] + 4 * sum(couts))
                drop.update(batch[1:])
                fused += len(batch) - 1
            batch = [k] if k is not None else []
    prog.ops[:] = [replace.get(k, op) for k, op in enumerate(ops) if k not in drop]
    return fused


# ------------------------------------------------------------------------------------------------------
# batch-chunked schedule of the high-resolution prefix
# ------------------------------------------------------------------------------------------------------
_PREFIX_RE = re.compile(r"^backbone\.(conv1|conv2|layer1\.|transition1\.|resnet\.conv1|resnet\.layer1\.)")


def chunk_prefix(prog: Program, chunk: int = None):
    """The stem, layer1 and transition1 of HRNet (pose_hrnet.py:464-472; ResNet-50's stem and layer1 for CPN) work on the
    largest tensors of the network -- [B, H/4, W/4, 256] is 537 MB at bs = 256 -- and are bound by HBM: every op streams its
    whole input from and its whole output to DRAM.  Frames are independent, so the same ops are issued per CHUNK of frames
    (all prefix ops for frames [0, c), then [c, 2c), ...): a chunk's tensors (c * 2 MB) are still in the 126 MB L2 when the
    next op reads them.  Pure scheduling: every op becomes ceil(B / c) ops on sub-views of the same buffers; results are
    bit-identical.

    MEASURED on B200 (bs = 256, HRNet-32, profiles/r2_chunk_sweep.txt): it does NOT pay -- 11.72 ms/step unchunked vs 12.85 /
    12.19 / 11.99 / 11.84 ms at 16 / 27 / 37 / 64 frames per chunk: the ~5 us of ramp + tail every extra launch costs
    outweighs the DRAM traffic saved.  The pass is therefore OFF by default; CAPF_CHUNK=<frames> turns it on for experiments."""
    env = os.environ.get("CAPF_CHUNK")
    if env is None and chunk is None:
        return 0
    if env is not None:
        chunk = int(env) if int(env) > 0 else None
        if int(env) == 0:
            return 0
    def in_prefix(op):
        if op.kind == lib.OP_MAXPOOL:
            return True
        return op.kind in (lib.OP_CONV2D, lib.OP_CAST) and bool(_PREFIX_RE.match(op.tag or ""))

    n = 0
    while n < len(prog.ops) and in_prefix(prog.ops[n]):
        n += 1
    if n == 0:
        return 0
    B = prog.B
    if chunk is None:
        # frames whose widest prefix tensor (H/4 x W/4 x 256 x 2 B) keeps ~1/2 of the L2 free for the op's other operand
        per_frame = max(int(np.prod(b.shape[1:])) * _ITEMSIZE[b.dtype] for op in prog.ops[:n] for b in op.outs if isinstance(b, Buf))
        chunk = max(1, int(56e6 // per_frame))
    if chunk <= 0 or chunk >= B:
        return 0
    head = prog.ops[:n]
    out = []
    for c0 in range(0, B, chunk):
        c = min(chunk, B - c0)
        for op in head:
            def sub(b):
                if not isinstance(b, Buf):
                    return b
                per = int(np.prod(b.shape[1:]))
                assert b.shape[0] == B, (op.tag, b.shape)
                return b.view(c0 * per, (c,) + tuple(b.shape[1:]), f"{b.name}[{c0}:{c0 + c}]")
            ins = [sub(b) for b in op.ins]
            outs = [sub(b) for b in op.outs]
            i = list(op.i)
            if op.kind == lib.OP_CAST:
                cnt = ins[0].numel
                i[0], i[1] = cnt & 0x7fffffff, cnt >> 31
            else:
                i[0] = c
            o = Op(op.kind, op.dtype_in, op.dtype_out, i, list(op.f), ins, outs, tag=op.tag, flops=op.flops * c // B,
                   nbytes=op.nbytes * c // B)
            o.pin_lane0 = True
            out.append(o)
    prog.ops[:n] = out
    return len(out)


# ------------------------------------------------------------------------------------------------------
# execution lanes: the branches of a HighResolutionModule are independent chains (pose_hrnet.py:289-290)
# ------------------------------------------------------------------------------------------------------
MAX_LANES = 4
_LANE_RE = re.compile(r"\.stage\d+\.\d+\.branches\.(\d+)\.|\.stage\d+\.\d+\.fuse_layers\.\d+\.(\d+)\.|\.transition\d+\.(\d+)\."
                      r"|\.refine_net\.cascade\.(\d+)\.")


# the per-level Linears of the lifter (pose_dformer.py:190-193 feat_embed, :108-110 embed_proj) read and write disjoint slabs
# of the token stream: level l on lane l, four ~2 us GEMMs side by side instead of in a row
_LIFTER_LANE_RE = re.compile(r"\.feat_embed\.(\d+)$|\.embed_proj\.(\d+)$")


def assign_lanes(prog: Program, enable: bool = True):
    """Lane (stream) of every op.  Branch b of an HRNet stage -- its BasicBlocks, the fuse-layer convolutions that read
    it, the transition that creates it and the fuse-sum that produces its next input -- runs on lane b; everything else
    (stem, layer1, CPN, the lifter) on lane 0.  Kernels of one lane are ordered; lanes only meet through the data
    dependencies op_schedule() turns into events.  One-wave kernels that leave SMs idle (C = 128 / 256 branches: 256 or
    128 tiles on 148 SMs) then overlap with the other branches instead of serialising behind them."""
    for op in prog.ops:
        op.lane = 0
    if not enable:
        return
    producer_lane = {}
    for op in prog.ops:
        lane = 0
        if getattr(op, "pin_lane0", False):
            pass
        elif op.kind == lib.OP_CONV2D:
            m = _LANE_RE.search(op.tag) or _LIFTER_LANE_RE.search(op.tag)
            if m:
                lane = int(next(g for g in m.groups() if g is not None))
        elif op.kind == lib.OP_BILINEAR:
            # CPN: the resize that ends a RefineNet cascade (refineNet.py:72-88) stays on the cascade's lane; the four cascades are
            # independent chains of small kernels (8x8 ... 64x64 maps) that then overlap instead of queueing
            src = op.ins[0]
            lane = producer_lane.get(src.root, 0) if isinstance(src, Buf) else 0
        elif op.kind == lib.OP_FUSE_SUM:
            # output index i = position of the identity term: terms are ordered j = 0..n-1, j > i carry a shift
            shifts = op.i[5:5 + op.i[4]]
            first_up = next((k for k, sft in enumerate(shifts) if sft > 0), len(shifts))
            ident = op.ins[first_up - 1]
            lane = producer_lane.get(ident.root, 0) if isinstance(ident, Buf) else 0
        op.lane = min(lane, MAX_LANES - 1)
        for b in op.outs:
            if isinstance(b, Buf):
                producer_lane[b.root] = op.lane


def op_clocks(prog: Program):
    """Vector clocks of the lane-parallel execution.  Returns (vc, waits): vc[k][l] = index of the latest op of lane l
    that is guaranteed complete before op k starts (-1: none); waits[k] = ops of OTHER lanes whose completion op k must
    wait for explicitly (at most one per lane: lanes are in-order)."""
    n = len(prog.ops)
    vc = [None] * n
    waits = [[] for _ in range(n)]
    last_on_lane = [-1] * MAX_LANES
    writers, readers = {}, {}       # root buffer -> [(op index, first byte, end byte)]: views of one root conflict only where they overlap

    def span(b):
        lo = b.root_offset * _ITEMSIZE[b.dtype]
        return b.root, lo, lo + b.nbytes

    def hits(table, r, lo, hi):
        return [d for d, a, z in table.get(r, ()) if a < hi and lo < z]

    for k, op in enumerate(prog.ops):
        deps = set()
        ins = [span(b) for b in op.ins if isinstance(b, Buf)]
        outs = [span(b) for b in op.outs if isinstance(b, Buf)]
        for r, lo, hi in ins:
            deps.update(hits(writers, r, lo, hi))           # read after write
        for r, lo, hi in outs:
            deps.update(hits(writers, r, lo, hi))           # write after write
            deps.update(hits(readers, r, lo, hi))           # write after read
        prev = last_on_lane[op.lane]
        clock = list(vc[prev]) if prev >= 0 else [-1] * MAX_LANES
        need = [-1] * MAX_LANES
        for d in deps:
            need[prog.ops[d].lane] = max(need[prog.ops[d].lane], d)
        for l in range(MAX_LANES):
            if l != op.lane and need[l] > clock[l]:
                waits[k].append(need[l])
        for d in waits[k]:
            clock = [max(a, b) for a, b in zip(clock, vc[d])]
        clock[op.lane] = k
        vc[k] = clock
        last_on_lane[op.lane] = k
        for r, lo, hi in ins:
            readers.setdefault(r, []).append((k, lo, hi))
        for r, lo, hi in outs:
            writers.setdefault(r, []).append((k, lo, hi))
    return vc, waits


# ------------------------------------------------------------------------------------------------------
# memory planning: greedy reuse of dead activation buffers (exact-size pools)
# ------------------------------------------------------------------------------------------------------
def _dying_residual(op, k, last_use, keep_alive):
    """The root buffer of op's residual operand if this conv / Linear is its last reader and it may be overwritten."""
    if op.kind not in (lib.OP_CONV2D, lib.OP_EXPAND_REDUCE) or len(op.ins) < 4 or not isinstance(op.ins[3], Buf) or not op.outs:
        return None
    res = op.ins[3]
    root = res.root
    if (last_use.get(root) != k or root in keep_alive or res.root_offset != 0 or res.numel * _ITEMSIZE[res.dtype] != root.nbytes
            or any(isinstance(x, Buf) and x.root is root for n, x in enumerate(op.ins) if n != 3)
            or any(isinstance(o, Buf) and o.root is root for o in op.outs)):
        return None
    return root


def plan_memory(prog: Program):
    """Returns (assignment: root Buf -> pool slot id, slots: list of (nbytes)).

    A slot is handed to a new buffer only if every op that touched its previous tenants is ordered before the new
    buffer's producer by the lane clocks (op_clocks) -- with a single lane this is plain liveness in program order."""
    vc, _ = op_clocks(prog)
    last_use = {}
    for k, op in enumerate(prog.ops):
        for b in list(op.ins) + list(op.outs):
            if isinstance(b, Buf):
                last_use[b.root] = k
    keep_alive = {b.root for b in prog.inputs.values()} | {b.root for b in prog.outputs.values()}
    assign, slots, free = {}, [], {}
    slot_users = []                 # slot -> per-lane latest op index that touched it
    for k, op in enumerate(prog.ops):
        donor = _dying_residual(op, k, last_use, keep_alive)
        for b in op.outs:
            if isinstance(b, Buf) and b.root not in assign:
                r = b.root
                if (donor is not None and r not in keep_alive and r.nbytes == donor.nbytes and b.root_offset == 0
                        and all(vc[k][l] >= slot_users[assign[donor]][l] for l in range(MAX_LANES))):
                    # y = act(conv(x) + res) with res read for the last time here: update res in place.  Every kernel
                    # reads a residual element in the thread group that later writes it (tests: residual_updated_in_place),
                    # and the write then lands on lines the read just brought into L2.
                    assign[r] = assign[donor]
                    last_use[donor] = -1                      # the slot now belongs to the output; do not free it below
                    donor = None
                    continue
                pool = free.get(r.nbytes, [])
                pick = None
                if r not in keep_alive:
                    for cand in pool:
                        if all(vc[k][l] >= slot_users[cand][l] for l in range(MAX_LANES)):
                            pick = cand
                            break
                if pick is not None:
                    pool.remove(pick)
                    assign[r] = pick
                else:
                    slots.append(r.nbytes)
                    slot_users.append([-1] * MAX_LANES)
                    assign[r] = len(slots) - 1
        for b in op.ins:
            if isinstance(b, Buf) and b.root not in assign:   # program inputs
                slots.append(b.root.nbytes)
                slot_users.append([-1] * MAX_LANES)
                assign[b.root] = len(slots) - 1
        touched = set(x.root for x in list(op.ins) + list(op.outs) if isinstance(x, Buf))
        for b in touched:
            u = slot_users[assign[b]]
            u[op.lane] = max(u[op.lane], k)
        for b in touched:
            if last_use[b] == k and b not in keep_alive:
                free.setdefault(b.nbytes, []).append(assign[b])
    return assign, slots


# ------------------------------------------------------------------------------------------------------
# device plan
# ------------------------------------------------------------------------------------------------------
class BufferStore:
    """Storage for every program buffer according to plan_memory (shared by Plan and tests/interp.py so the
    CPU test-suite exercises the same aliasing the GPU sees)."""

    def __init__(self, prog: Program, device):
        self.assign, slots = plan_memory(prog)
        self.slots = [torch.empty(max(n, 16), dtype=torch.uint8, device=device) for n in slots]
        self.nbytes = sum(slots)

    def tensor(self, b: Buf) -> torch.Tensor:
        raw = self.slots[self.assign[b.root]]
        it = _ITEMSIZE[b.dtype]
        start = b.root_offset * it
        return raw[start:start + b.numel * it].view(_TORCH_DT[b.dtype]).view(b.shape)

    def ptr(self, b: Buf) -> int:
        return self.slots[self.assign[b.root]].data_ptr() + b.root_offset * _ITEMSIZE[b.dtype]


def weight_slots(prog: Program):
    out = {}
    for op in prog.ops:
        for w in op.ins:
            if isinstance(w, WSlot):
                out.setdefault(id(w), w)
    return out


class Plan:
    """A Program bound to device buffers and packed weights, executed through libcapf_b200."""

    def __init__(self, prog: Program, state: dict, device):
        import ctypes as C
        self.prog = prog
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise lib.CapfError("Plan needs a CUDA device: libcapf_b200 has no CPU path")
        L = lib.load()
        self.mem = BufferStore(prog, self.device)
        self.workspace_bytes = self.mem.nbytes
        self._wslots = weight_slots(prog)
        self._wtensors = {}
        self.repack(state)
        arr = (lib.CapfOp * len(prog.ops))()
        for k, op in enumerate(prog.ops):
            c = arr[k]
            c.kind = op.kind
            c.dtype_in = lib.DTYPE_CODE[op.dtype_in]
            c.dtype_out = lib.DTYPE_CODE[op.dtype_out]
            for n, v in enumerate(op.i):
                c.i[n] = int(v)
            for n, v in enumerate(op.f):
                c.f[n] = float(v)
            for n, b in enumerate(op.ins):
                c.inp[n] = self._ptr(b)
            for n, b in enumerate(op.outs):
                c.out[n] = self._ptr(b)
        self._ops = arr
        h = C.c_void_p()
        lib.check(L.capf_plan_create(arr, len(prog.ops), self.device.index or 0, C.byref(h)), "capf_plan_create")
        self._h = h
        self._L = L
        self._graph = None
        self.post = None           # optional callable(torch stream): enqueued after every full run (e.g. the output all-gather)

    def _ptr(self, b):
        if b is None:
            return None
        if isinstance(b, WSlot):
            return self._wtensors[id(b)].data_ptr()
        return self.mem.ptr(b)

    def tensor(self, b: Buf) -> torch.Tensor:
        """Typed torch view of a program buffer (inputs/outputs/feature maps)."""
        return self.mem.tensor(b)

    def repack(self, state):
        """(Re)pack every weight blob from `state` into its fixed device tensor (pointers stay valid)."""
        for k, w in self._wslots.items():
            t = w.pack(state)
            assert tuple(t.shape) == tuple(w.shape) or t.numel() == int(np.prod(w.shape)), (w.name, t.shape, w.shape)
            if k in self._wtensors:
                self._wtensors[k].copy_(t.reshape(self._wtensors[k].shape), non_blocking=False)
            else:
                self._wtensors[k] = t.to(self.device)

    # ---- lane-parallel schedule ---------------------------------------------------------------------
    def _schedule(self):
        """Segments (lane, first, count, wait_ops, record) of one full run: maximal runs of consecutive ops of one lane
        with the cross-lane waits in front and an event recorded behind when another lane depends on the last op."""
        if getattr(self, "_segs", None) is not None:
            return self._segs
        ops = self.prog.ops
        _, waits = op_clocks(self.prog)
        recorded = set(d for w in waits for d in w)
        last_of_lane = {}
        for k, op in enumerate(ops):
            last_of_lane[op.lane] = k
        self._tails = [k for l, k in last_of_lane.items() if l != 0]
        recorded.update(self._tails)
        segs = []
        for k, op in enumerate(ops):
            if segs and segs[-1][0] == op.lane and not waits[k] and not segs[-1][4]:
                lane, first, count, w, _ = segs[-1]
                segs[-1] = (lane, first, count + 1, w, k in recorded)
            else:
                segs.append((op.lane, k, 1, list(waits[k]), k in recorded))
        self._segs = segs
        n_side = max([op.lane for op in ops] + [0])
        self._side = [torch.cuda.Stream(self.device) for _ in range(n_side)]
        self._events = {k: torch.cuda.Event() for k in recorded}
        return segs

    def run(self, first=0, count=-1, stream=None):
        """Enqueue ops [first, first + count) (default: all).  A full run spreads the lanes of the program over side
        streams that fork from / join back into `stream` (a torch.cuda.Stream; default: the current one), so it is
        still one stream-ordered, graph-capturable unit of work for the caller.  Partial runs are single-stream."""
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        s = st.cuda_stream
        n_ops = len(self.prog.ops)
        if os.environ.get("CAPF_DEBUG_SYNC", "0") != "0":      # locate a faulting op: one launch + sync at a time
            n = n_ops - first if count < 0 else count
            for k in range(first, first + n):
                lib.check(self._L.capf_plan_run(self._h, k, 1, s), "capf_plan_run")
                try:
                    torch.cuda.synchronize(self.device)
                except Exception as e:  # noqa: BLE001
                    op = self.prog.ops[k]
                    raise lib.CapfError(f"op {k} ({op.tag}, kind {op.kind}, i={op.i}) faulted: {e}") from e
            return
        full = first == 0 and (count < 0 or count == n_ops)
        segs = self._schedule() if full else None
        if not full or not self._side:
            lib.check(self._L.capf_plan_run(self._h, first, count, s), "capf_plan_run")
            if full and self.post is not None:
                self.post(st)
            return
        streams = [st] + self._side
        for lane, k0, n, waits, record in segs:
            ls = streams[lane]
            for w in waits:
                ls.wait_event(self._events[w])
            lib.check(self._L.capf_plan_run(self._h, k0, n, ls.cuda_stream), "capf_plan_run")
            if record:
                self._events[k0 + n - 1].record(ls)
        for k in self._tails:                                   # join: the caller's stream owns the result again
            st.wait_event(self._events[k])
        if self.post is not None:
            self.post(st)

    def op_kernel(self, k: int) -> str:
        """Name (and tile shape) of the kernel op k launches, as reported by the library."""
        import ctypes as C
        buf = C.create_string_buffer(160)
        lib.check(self._L.capf_plan_op_kernel(self._h, k, buf, 160), "capf_plan_op_kernel")
        return buf.value.decode()

    @property
    def num_launches(self):
        return len(self.prog.ops)

    def time_ops(self, passes: int = 2):
        """Per-op device time (ms) of one in-order pass, CUDA events on the launching stream around every op.
        The passes are enqueued back to back WITHOUT synchronising in between and the last one is returned: while the
        GPU works through the earlier passes the host runs ahead, so the measured pass is not host-launch-bound (the gap
        between two events is the kernel's own duration incl. its launch latency, not the host's issue interval) and
        caches are in their steady in-step state (not warm per op)."""
        st = torch.cuda.current_stream(self.device)
        n = len(self.prog.ops)
        passes = max(1, passes)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(passes)]
        for ev in evs:
            ev[0].record(st)
            for k in range(n):
                lib.check(self._L.capf_plan_run(self._h, k, 1, st.cuda_stream), "capf_plan_run")
                ev[k + 1].record(st)
        st.synchronize()
        ev = evs[-1]
        return [ev[k].elapsed_time(ev[k + 1]) for k in range(n)]

    def time_op_repeated(self, k: int, reps: int = 20):
        """Average device time (ms) of `reps` back-to-back launches of op k (must be idempotent: no in-place residual)
        between two CUDA events on the launching stream -- the kernel's steady duration without the per-op event and
        host launch gaps that time_ops() includes.  Operands stay in their in-step cache state (L2-resident)."""
        op = self.prog.ops[k]
        if any(isinstance(b, Buf) and any(b.root is o.root for o in op.outs if isinstance(o, Buf)) for b in op.ins):
            raise ValueError("time_op_repeated: op updates its input in place")
        st = torch.cuda.current_stream(self.device)
        for _ in range(3):
            lib.check(self._L.capf_plan_run(self._h, k, 1, st.cuda_stream), "capf_plan_run")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            lib.check(self._L.capf_plan_run(self._h, k, 1, st.cuda_stream), "capf_plan_run")
        e1.record(st)
        st.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- CUDA graph ---------------------------------------------------------------------------------
    def capture(self):
        """Capture one full run into a CUDA graph (replayed by run_graph)."""
        torch.cuda.synchronize(self.device)
        side = torch.cuda.Stream(self.device)
        with torch.cuda.stream(side):
            self.run()                      # warm-up outside capture (lazy module load, func attributes)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        # thread_local: a collective in `post` keeps NCCL's watchdog thread alive next to the capture
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local" if self.post is not None else "global"):
            self.run(stream=side)
        self._graph = g
        return g

    def run_graph(self):
        if self._graph is None:
            self.capture()
        self._graph.replay()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.capf_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
