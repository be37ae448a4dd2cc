"""Training step of ``volume_net``: forward with saved activations, hand-written backward, AdamW.

The reference trains only the lifter (``fix_weights``: the backbone is frozen and kept in ``eval()``, conpose.py:22-25,
train.py:145-148) by calling ``model(images, kp2d, kp2d_crop)`` under autograd, ``loss.backward()`` and ``AdamW.step()``
(train.py:186-201, :337-345).  Here the same three calls work on ``capf_b200.CA_PF``:

* ``CA_PF.forward`` in ``volume_net.train()`` mode with grad enabled runs the (frozen) backbone through its inference plan
  and the lifter through :class:`LifterFunction` -- a ``torch.autograd.Function`` whose forward and backward are
  sequences of libcapf_b200 kernels (fp32 storage; the forward and dgrad GEMMs on the tensor cores with split bfloat16
  operands -- the `bf16x3` mode of the inference kernel, fp32-class --, wgrad and everything else in csrc/capf_train.cu, plus
  the inference samplers / LayerNorm / attention kernels);
  autograd only routes the gradients into ``p.grad`` of the 191 ``volume_net`` parameters.
* DropPath (timm ``drop_path``; rates ``linspace(0, 0.2, 4)``, pose_dformer.py:71,101,187): the per-sample keep masks
  are drawn with the same torch calls, in the same order and with the same shapes as the reference draws them, so a
  seeded run sees the same masks; the scaling itself is fused into the residual-add kernel.
* :class:`FusedAdamW` -- ``torch.optim.AdamW``'s update as ONE kernel over a flat parameter buffer (the parameters and
  their ``.grad`` become views of flat buffers); ``torch.optim.AdamW`` itself keeps working on the same parameters.

PyTorch is used for device memory, the autograd hand-over and RNG only; there is no CPU path.
"""
import ctypes
import math

import torch

from . import lib

J = 17
HEADS_CTX, SAMPLES = 4, 4


# ------------------------------------------------------------------------------------------------------
# thin launch helpers (one libcapf_b200 op each) on torch tensors
# ------------------------------------------------------------------------------------------------------
_DT = {torch.float32: lib.F32, torch.float16: lib.F16, torch.bfloat16: lib.BF16}


PROFILE = None      # set to a list to collect (kind, shape ints, start event, end event) per launch (tools/train_bench.py --profile)


def _run(kind, dt_in, dt_out, ints, floats, ins, outs, dev):
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _run_op(kind, dt_in, dt_out, ints, floats, ins, outs, dev)
        e1.record()
        PROFILE.append((kind, tuple(int(v) for v in ints[:5]), e0, e1))
        return
    _run_op(kind, dt_in, dt_out, ints, floats, ins, outs, dev)


def _run_op(kind, dt_in, dt_out, ints, floats, ins, outs, dev):
    op = lib.CapfOp()
    op.kind, op.dtype_in, op.dtype_out = kind, dt_in, dt_out
    for n, v in enumerate(ints):
        op.i[n] = int(v)
    for n, v in enumerate(floats):
        op.f[n] = float(v)
    for n, t in enumerate(ins):
        op.inp[n] = None if t is None else t.data_ptr()
    for n, t in enumerate(outs):
        op.out[n] = None if t is None else t.data_ptr()
    st = torch.cuda.current_stream(dev).cuda_stream
    lib.check(lib.load().capf_op_run(ctypes.byref(op), dev.index or 0, st), f"train op {kind}")


def _f32(*shape, dev):
    return torch.empty(*shape, dtype=torch.float32, device=dev)


def gemm(A, B, M, N, K, ta, tb, lda, ldb, out, ldc, bias=None, accumulate=False, splits=1):
    ws = None
    if splits > 1:
        ws = _f32(splits * M * N, dev=out.device)
    _run(lib.OP_GEMM_F32, lib.F32, lib.F32, [M, N, K, ta, tb, lda, ldb, ldc, splits, 1 if accumulate else 0], [], [A, B, bias], [out, ws], out.device)
    return out


# ---- tensor-core path of the forward and dgrad GEMMs: split bfloat16 operands (precision "bf16x3": hi*Wh + lo*Wh + hi*Wl, fp32
# accumulate, ~2^-16 operand precision) through the inference tcgen05 kernel.  Weights change every step, so they are packed on
# the device (CAPF_OP_CAST i[3]); plans (TMA descriptors) are cached by pointers -- the caching allocator hands a steady-state
# training loop the same addresses every step.  CAPF_TRAIN_TC=0 keeps everything on the fp32 CUDA-core GEMM.
import os as _os

USE_TC = _os.environ.get("CAPF_TRAIN_TC", "1") != "0"
_TC_PLANS = {}
_TC_PLANS_MAX = 2048
_PACKED = None          # split-packed weights of the step in flight ({("f" | "t", data_ptr): tensor}); set by LifterFunction


def _tc_ok(M, N, K):
    return USE_TC and K % 16 == 0 and N % 16 == 0 and M >= 64


def _split_rows(x, M, K):
    """fp32 [M][K] -> bf16 [M][hi (K) | lo (K)] (CAPF_OP_CAST split planes)."""
    xs = torch.empty(M, 2 * K, dtype=torch.bfloat16, device=x.device)
    n = M * K
    _run(lib.OP_CAST, lib.F32, lib.BF16, [n & 0x7fffffff, n >> 31, K], [], [x], [xs], x.device)
    return xs


def pack_split_weight(W, rows, C, transposed=False):
    """fp32 weight [rows][C] (or its transpose stored [C][rows]) -> bf16 [rows][hi | hi | lo] for a split-operand GEMM."""
    wp = torch.empty(rows, 3 * C, dtype=torch.bfloat16, device=W.device)
    n = rows * C
    _run(lib.OP_CAST, lib.F32, lib.BF16, [n & 0x7fffffff, n >> 31, C, 2 if transposed else 1], [], [W], [wp], W.device)
    return wp


def _tc_gemm(xs, wp, bias, M, K, N, out, residual=None):
    """out[M][N] = (hi + lo)[M][K] . W^T (+ bias) (+ residual), through tc_gemm_kernel with split operands."""
    import ctypes as C
    dev = out.device
    key = (xs.data_ptr(), wp.data_ptr(), 0 if bias is None else bias.data_ptr(), out.data_ptr(),
           0 if residual is None else residual.data_ptr(), M, K, N, dev.index or 0)
    h = _TC_PLANS.get(key)
    L = lib.load()
    if h is None:
        if len(_TC_PLANS) >= _TC_PLANS_MAX:
            for old in _TC_PLANS.values():
                L.capf_plan_destroy(old)
            _TC_PLANS.clear()
        op = lib.CapfOp()
        op.kind, op.dtype_in, op.dtype_out = lib.OP_CONV2D, lib.BF16, lib.F32
        for n, v in enumerate([M, 1, 1, K, N, 1, 1, 1, 0, 1, 1, lib.ACT_NONE, lib.IMPL_TCGEN05, 1, 0, 0, 0, 0, 1]):
            op.i[n] = v
        op.inp[0], op.inp[1] = xs.data_ptr(), wp.data_ptr()
        op.inp[2] = None if bias is None else bias.data_ptr()
        op.inp[3] = None if residual is None else residual.data_ptr()
        op.out[0] = out.data_ptr()
        h = C.c_void_p()
        arr = (lib.CapfOp * 1)(op)
        lib.check(L.capf_plan_create(arr, 1, dev.index or 0, C.byref(h)), "train tc plan")
        _TC_PLANS[key] = h
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.check(L.capf_plan_run(h, 0, 1, torch.cuda.current_stream(dev).cuda_stream), "train tc gemm")
        e1.record()
        PROFILE.append((lib.OP_CONV2D, (M, N, K), e0, e1))
        return out
    lib.check(L.capf_plan_run(h, 0, 1, torch.cuda.current_stream(dev).cuda_stream), "train tc gemm")
    return out


def linear(x, W, b, M, out=None, accumulate=False, ldc=None, packed=None):
    """y[M][N] = x[M][K] W[N][K]^T + b   (nn.Linear).  `packed`: dict caching the split-packed weights of this step."""
    N, K = W.shape
    if out is None:
        out = _f32(M, N, dev=x.device)
    packed = _PACKED if packed is None else packed
    if _tc_ok(M, N, K) and (ldc is None or ldc == N):
        wp = None if packed is None else packed.get(("f", W.data_ptr()))
        if wp is None:
            wp = pack_split_weight(W, N, K)
            if packed is not None:
                packed[("f", W.data_ptr())] = wp
        return _tc_gemm(_split_rows(x, M, K), wp, b, M, K, N, out, residual=out if accumulate else None)
    return gemm(x, W, M, N, K, 0, 1, K, K, out, ldc or N, b, accumulate)


def linear_bwd(x, W, dy, M, need_dx=True, need_db=True, packed=None):
    """-> (dx [M][K] or None, dW [N][K], db [N] or None) of y = x W^T + b."""
    N, K = W.shape
    dev = x.device
    dx = None
    packed = _PACKED if packed is None else packed
    if need_dx and _tc_ok(M, K, N):
        wt = None if packed is None else packed.get(("t", W.data_ptr()))
        if wt is None:
            wt = pack_split_weight(W, K, N, transposed=True)        # W^T [K][N]: the "weight" of dx = dy . W
            if packed is not None:
                packed[("t", W.data_ptr())] = wt
        dx = _tc_gemm(_split_rows(dy, M, N), wt, None, M, N, K, _f32(M, K, dev=dev))
    elif need_dx:
        dx = gemm(dy, W, M, K, N, 0, 0, N, K, _f32(M, K, dev=dev), K)
    tiles = ((N + 63) // 64) * ((K + 63) // 64)
    splits = max(1, min(64, (296 + tiles - 1) // tiles, M // 256))
    dW = gemm(dy, x, N, K, M, 1, 0, N, K, _f32(N, K, dev=dev), K, splits=splits)
    db = colsum(dy, M, N) if need_db else None
    return dx, dW, db


def colsum(x, M, N, ld=None, out=None, accumulate=False):
    dev = x.device
    if out is None:
        out = _f32(N, dev=dev)
    ws = _f32(256 * N, dev=dev) if M > 512 else None
    _run(lib.OP_COLSUM, lib.F32, lib.F32, [M, N, ld or N, 1 if accumulate else 0], [], [x], [out, ws], dev)
    return out


def layernorm(x, rows, D, gamma, beta, eps, x0=None, period=0):
    out = _f32(rows, D, dev=x.device)
    _run(lib.OP_LAYERNORM, lib.F32, lib.F32, [rows, D, period], [eps], [x, gamma, beta, x0], [out], x.device)
    return out


def layernorm_bwd(x, rows, D, gamma, eps, dy, dx, accumulate, x0=None, period=0):
    """dx (+)= LayerNorm backward; returns (dgamma, dbeta)."""
    dev = x.device
    nblocks = max(1, min(296, (rows + 7) // 8))
    part = _f32(nblocks, 2 * D, dev=dev)
    _run(lib.OP_LAYERNORM_BWD, lib.F32, lib.F32, [rows, D, period, 1 if accumulate else 0, nblocks], [eps], [x, gamma, dy, x0], [dx, part], dev)
    gb = colsum(part, nblocks, 2 * D)
    return gb[:D], gb[D:]


def gelu(h):
    y = torch.empty_like(h)
    n = h.numel()
    _run(lib.OP_GELU, lib.F32, lib.F32, [n & 0x7fffffff, n >> 31], [], [h], [y], h.device)
    return y


def gelu_bwd(h, dy):
    dh = torch.empty_like(h)
    n = h.numel()
    _run(lib.OP_GELU_BWD, lib.F32, lib.F32, [n & 0x7fffffff, n >> 31], [], [h, dy], [dh], h.device)
    return dh


def rows_axpy(t, y, rows, D, scale=None, mod=1, div=1, accumulate=True):
    _run(lib.OP_ROWS_AXPY, lib.F32, lib.F32, [rows, D, mod, div, 1 if accumulate else 0], [], [t, scale], [y], t.device)
    return y


def attention(qkv, rows, dim, heads, groups, seq, tok_stride, grp_stride):
    out = _f32(rows, dim, dev=qkv.device)
    hd = dim // heads
    _run(lib.OP_ATTENTION, lib.F32, lib.F32, [groups, seq, heads, hd, tok_stride, grp_stride], [float(hd) ** -0.5], [qkv], [out], qkv.device)
    return out


def attention_bwd(qkv, dout, rows, dim, heads, groups, seq, tok_stride, grp_stride):
    dqkv = _f32(rows, 3 * dim, dev=qkv.device)
    hd = dim // heads
    _run(lib.OP_ATTENTION_BWD, lib.F32, lib.F32, [groups, seq, heads, hd, tok_stride, grp_stride], [float(hd) ** -0.5], [qkv, dout], [dqkv], qkv.device)
    return dqkv


# ------------------------------------------------------------------------------------------------------
# DropPath masks, drawn like the reference draws them
# ------------------------------------------------------------------------------------------------------
def drop_path_rates(depth=4, drop_path_rate=0.2):
    """pose_dformer.py:187: dpr = linspace(0, drop_path_rate, depth); block i of every group uses dpr[i]."""
    return [float(v) for v in torch.linspace(0, drop_path_rate, depth)]


def draw_drop_path_scales(B, device, depth=4, drop_path_rate=0.2, generator=None):
    """Per-block scale vectors (bernoulli(keep) / keep, timm.drop_path with scale_by_keep) in the order the reference's forward
    consumes random numbers: context_blocks (mask per frame: x is [b, l, p, c]), res_blocks (per (frame, joint): x is
    [(b p), l, c]), joint_blocks (per frame), two draws per block (attention / sampling branch, then the MLP branch); a block with rate
    0 holds nn.Identity and draws nothing (pose_dformer.py:71,101).  Returns {group: [(s1, s2) or (None, None)] * depth}."""
    rates = drop_path_rates(depth, drop_path_rate)
    out = {}
    for group, n in (("context_blocks", B), ("res_blocks", B * J), ("joint_blocks", B)):
        rows = []
        for r in rates:
            if r <= 0.0:
                rows.append((None, None))
                continue
            keep = 1.0 - r
            pair = []
            for _ in range(2):
                m = torch.empty(n, dtype=torch.float32, device=device).bernoulli_(keep, generator=generator)
                pair.append(m.div_(keep))
            rows.append(tuple(pair))
        out[group] = rows
    return out


# ------------------------------------------------------------------------------------------------------
# the lifter: forward with saved activations, backward
# ------------------------------------------------------------------------------------------------------
class Geometry:
    def __init__(self, B, D, dims, map_hw, map_dtype):
        self.B, self.D, self.dims, self.map_hw, self.map_dtype = B, D, list(dims), list(map_hw), map_dtype
        self.R = B * J
        self.L = len(dims)
        self.E = D * (self.L + 1)
        geo = []
        for (h, w), c in zip(map_hw, dims):
            geo += [h, w, c]
        self.map_geo = geo + [0] * (12 - len(geo))
        self.offs = [0]
        self.goffs = [0]
        for c in dims:
            self.offs.append(self.offs[-1] + self.R * c)
            self.goffs.append(self.goffs[-1] + self.R * 4 * c)


def _mlp_fwd(P, q, x, rows, sv):
    sv["h"] = linear(x, P[q + ".mlp.fc1.weight"], P[q + ".mlp.fc1.bias"], rows)
    sv["a"] = gelu(sv["h"])
    return linear(sv["a"], P[q + ".mlp.fc2.weight"], P[q + ".mlp.fc2.bias"], rows)


def _mlp_bwd(P, G, q, t2, dm, rows, sv):
    da, G[q + ".mlp.fc2.weight"], G[q + ".mlp.fc2.bias"] = linear_bwd(sv["a"], P[q + ".mlp.fc2.weight"], dm, rows)
    dh = gelu_bwd(sv["h"], da)
    dt2, G[q + ".mlp.fc1.weight"], G[q + ".mlp.fc1.bias"] = linear_bwd(t2, P[q + ".mlp.fc1.weight"], dh, rows)
    return dt2


def lifter_forward(P, g: Geometry, maps, kp2d, ref, scales):
    """PoseTransformer.forward (pose_dformer.py:210-241) in training mode.  P: {name: fp32 tensor} of volume_net; maps: 4 NHWC
    feature maps (g.map_dtype); kp2d, ref: [R, 2] fp32; scales: draw_drop_path_scales() or None (no DropPath).
    Returns (out [R, 3], saved)."""
    dev = kp2d.device
    B, R, D, L, E = g.B, g.R, g.D, g.L, g.E
    S = L + 1
    mdt = _DT[g.map_dtype]
    sv = {"ctx": [], "res": [], "joint": []}
    X = _f32(S, R, D, dev=dev)
    _run(lib.OP_EMBED_COORD, lib.F32, lib.F32, [B, J, D, S], [],
         [kp2d, P["coord_embed.weight"], P["coord_embed.bias"], P["Spatial_pos_embed"]], [X], dev)
    samp = _f32(g.offs[-1], dev=dev)
    _run(lib.OP_REF_SAMPLE, mdt, lib.F32, [B, J, L] + g.map_geo + g.offs[:4], [], [ref] + list(maps), [samp, None], dev)
    for l in range(L):
        linear(samp[g.offs[l]:g.offs[l + 1]], P[f"feat_embed.{l}.weight"], P[f"feat_embed.{l}.bias"], R, out=X[1 + l], accumulate=True)
    sv["samp"] = samp
    Xl = X[1:]                                      # [L, R, D] contiguous view
    for i in range(L):
        q = f"context_blocks.{i}"
        s1, s2 = scales["context_blocks"][i] if scales else (None, None)
        b = {"Xin": X.clone(), "s1": s1, "s2": s2}
        b["t"] = layernorm(Xl, L * R, D, P[q + ".norm1.weight"], P[q + ".norm1.bias"], 1e-5, x0=X[0], period=R)
        b["ow"] = linear(b["t"], P[q + ".ow.weight"], P[q + ".ow.bias"], L * R)
        b["g"] = _f32(g.goffs[-1], dev=dev)
        _run(lib.OP_DEFORM_SAMPLE, mdt, lib.F32, [B, J, L] + g.map_geo + g.goffs[:4], [], [ref] + list(maps) + [b["ow"]], [b["g"], None], dev)
        u = _f32(L, R, D, dev=dev)
        for l in range(L):
            linear(b["g"][g.goffs[l]:g.goffs[l + 1]], P[f"{q}.embed_proj.{l}.weight"], P[f"{q}.embed_proj.{l}.bias"], R * 4, out=u[l])
        rows_axpy(u, Xl, L * R, D, s1, mod=R, div=J)
        b["Xmid"] = Xl.clone()
        b["t2"] = layernorm(Xl, L * R, D, P[q + ".norm2.weight"], P[q + ".norm2.bias"], 1e-5)
        m = _mlp_fwd(P, q, b["t2"], L * R, b)
        rows_axpy(m, Xl, L * R, D, s2, mod=R, div=J)
        sv["ctx"].append(b)

    def block(q, x, rows, dim, groups, seq, ts, gs, mod, div, s12):
        b = {"X0": x.clone(), "s1": s12[0], "s2": s12[1]}
        b["t"] = layernorm(x, rows, dim, P[q + ".norm1.weight"], P[q + ".norm1.bias"], 1e-6)
        b["qkv"] = linear(b["t"], P[q + ".attn.qkv.weight"], P[q + ".attn.qkv.bias"], rows)
        b["att"] = attention(b["qkv"], rows, dim, 8, groups, seq, ts, gs)
        o = linear(b["att"], P[q + ".attn.proj.weight"], P[q + ".attn.proj.bias"], rows)
        rows_axpy(o, x, rows, dim, s12[0], mod=mod, div=div)
        b["X1"] = x.clone()
        b["t2"] = layernorm(x, rows, dim, P[q + ".norm2.weight"], P[q + ".norm2.bias"], 1e-6)
        m = _mlp_fwd(P, q, b["t2"], rows, b)
        rows_axpy(m, x, rows, dim, s12[1], mod=mod, div=div)
        return b

    for i in range(L):      # res_blocks: 5 level tokens of one joint; DropPath mask per (frame, joint) = row % R
        sv["res"].append(block(f"res_blocks.{i}", X, S * R, D, R, S, R, 1, R, 1, scales["res_blocks"][i] if scales else (None, None)))
    Y = _f32(R, E, dev=dev)
    _run(lib.OP_LEVELS_TO_JOINT, lib.F32, lib.F32, [R, S, D], [], [X], [Y], dev)
    for i in range(L):      # joint_blocks: 17 joints of one frame; mask per frame = row // 17
        sv["joint"].append(block(f"joint_blocks.{i}", Y, R, E, B, J, 1, J, R, J, scales["joint_blocks"][i] if scales else (None, None)))
    sv["Y"] = Y
    sv["tY"] = layernorm(Y, R, E, P["head.0.weight"], P["head.0.bias"], 1e-5)
    out = linear(sv["tY"], P["head.1.weight"], P["head.1.bias"], R)
    return out, sv


def lifter_backward(P, g: Geometry, maps, kp2d, ref, sv, dout):
    """d loss / d parameters given dout [R, 3].  Returns {name: grad} for every entry of P."""
    dev = dout.device
    B, R, D, L, E = g.B, g.R, g.D, g.L, g.E
    S = L + 1
    mdt = _DT[g.map_dtype]
    G = {}
    dtY, G["head.1.weight"], G["head.1.bias"] = linear_bwd(sv["tY"], P["head.1.weight"], dout, R)
    dY = _f32(R, E, dev=dev)
    G["head.0.weight"], G["head.0.bias"] = layernorm_bwd(sv["Y"], R, E, P["head.0.weight"], 1e-5, dtY, dY, False)

    def block_bwd(q, b, dx, rows, dim, groups, seq, ts, gs, mod, div):
        """dx: gradient w.r.t. the block's output, updated in place to the gradient w.r.t. its input."""
        dm = rows_axpy(dx, _f32(rows, dim, dev=dev), rows, dim, b["s2"], mod=mod, div=div, accumulate=False)
        dt2 = _mlp_bwd(P, G, q, b["t2"], dm, rows, b)
        G[q + ".norm2.weight"], G[q + ".norm2.bias"] = layernorm_bwd(b["X1"], rows, dim, P[q + ".norm2.weight"], 1e-6, dt2, dx, True)
        do = rows_axpy(dx, _f32(rows, dim, dev=dev), rows, dim, b["s1"], mod=mod, div=div, accumulate=False)
        datt, G[q + ".attn.proj.weight"], G[q + ".attn.proj.bias"] = linear_bwd(b["att"], P[q + ".attn.proj.weight"], do, rows)
        dqkv = attention_bwd(b["qkv"], datt, rows, dim, 8, groups, seq, ts, gs)
        dt, G[q + ".attn.qkv.weight"], G[q + ".attn.qkv.bias"] = linear_bwd(b["t"], P[q + ".attn.qkv.weight"], dqkv, rows)
        G[q + ".norm1.weight"], G[q + ".norm1.bias"] = layernorm_bwd(b["X0"], rows, dim, P[q + ".norm1.weight"], 1e-6, dt, dx, True)

    for i in reversed(range(L)):
        block_bwd(f"joint_blocks.{i}", sv["joint"][i], dY, R, E, B, J, 1, J, R, J)
    dX = _f32(S, R, D, dev=dev)
    _run(lib.OP_JOINT_TO_LEVELS, lib.F32, lib.F32, [R, S, D], [], [dY], [dX], dev)
    for i in reversed(range(L)):
        block_bwd(f"res_blocks.{i}", sv["res"][i], dX, S * R, D, R, S, R, 1, R, 1)
    dXl = dX[1:]
    for i in reversed(range(L)):
        q = f"context_blocks.{i}"
        b = sv["ctx"][i]
        rows = L * R
        dm = rows_axpy(dXl, _f32(rows, D, dev=dev), rows, D, b["s2"], mod=R, div=J, accumulate=False)
        dt2 = _mlp_bwd(P, G, q, b["t2"], dm, rows, b)
        G[q + ".norm2.weight"], G[q + ".norm2.bias"] = layernorm_bwd(b["Xmid"], rows, D, P[q + ".norm2.weight"], 1e-5, dt2, dXl, True)
        du = rows_axpy(dXl, _f32(L, R, D, dev=dev), rows, D, b["s1"], mod=R, div=J, accumulate=False)
        dg = _f32(g.goffs[-1], dev=dev)
        for l in range(L):
            gl = b["g"][g.goffs[l]:g.goffs[l + 1]]
            W = P[f"{q}.embed_proj.{l}.weight"]                                  # [32][C_l]
            dgl, G[f"{q}.embed_proj.{l}.weight"], G[f"{q}.embed_proj.{l}.bias"] = linear_bwd(gl, W, du[l], R * 4)
            dg[g.goffs[l]:g.goffs[l + 1]].copy_(dgl.reshape(-1))
        dow = _f32(rows, 48, dev=dev)
        _run(lib.OP_DEFORM_BWD, mdt, lib.F32, [B, J, L] + g.map_geo + g.goffs[:4], [], [ref] + list(maps) + [b["ow"]], [dow, dg], dev)
        dt, G[q + ".ow.weight"], G[q + ".ow.bias"] = linear_bwd(b["t"], P[q + ".ow.weight"], dow, rows)
        dZ = _f32(L, R, D, dev=dev)
        Xin = b["Xin"]
        G[q + ".norm1.weight"], G[q + ".norm1.bias"] = layernorm_bwd(Xin[1:], rows, D, P[q + ".norm1.weight"], 1e-5, dt, dZ, False, x0=Xin[0], period=R)
        rows_axpy(dZ, dXl, rows, D)                       # d / d x_l
        for l in range(L):                                # d / d x_0: norm1 sees x_l + x_0 on every level (:120)
            rows_axpy(dZ[l], dX[0], R, D)
    for l in range(L):
        _, G[f"feat_embed.{l}.weight"], G[f"feat_embed.{l}.bias"] = linear_bwd(sv["samp"][g.offs[l]:g.offs[l + 1]], P[f"feat_embed.{l}.weight"],
                                                                                 dX[1 + l], R, need_dx=False)
    _, G["coord_embed.weight"], G["coord_embed.bias"] = linear_bwd(kp2d, P["coord_embed.weight"], dX[0], R, need_dx=False)
    dpos = _f32(S, J * D, dev=dev)
    for s in range(S):                                    # Spatial_pos_embed [1, S, J, D] is broadcast over the frames
        colsum(dX[s], B, J * D, out=dpos[s])
    G["Spatial_pos_embed"] = dpos.view(1, S, J, D)
    return G


# ------------------------------------------------------------------------------------------------------
# autograd hand-over
# ------------------------------------------------------------------------------------------------------
def _param_view(named):
    """volume_net parameters as the flat {name: contiguous fp32 tensor} the kernels read; attention_weights / sampling_offsets of a
    DeformableBlock are applied as ONE Linear to 48 columns (16 logits, then 32 offsets)."""
    P = {}
    for n, p in named.items():
        P[n] = p.detach().contiguous().float()
    for n in list(P):
        if n.endswith(".attention_weights.weight"):
            q = n[: -len(".attention_weights.weight")]
            P[q + ".ow.weight"] = torch.cat([P[q + ".attention_weights.weight"], P[q + ".sampling_offsets.weight"]], 0).contiguous()
            P[q + ".ow.bias"] = torch.cat([P[q + ".attention_weights.bias"], P[q + ".sampling_offsets.bias"]], 0).contiguous()
    return P


class LifterFunction(torch.autograd.Function):
    """volume_net(kp2d, ref, maps) with libcapf_b200 kernels on both sides of autograd."""

    @staticmethod
    def forward(ctx, geometry, maps, scales, names, kp2d, ref, *params):
        global _PACKED
        named = dict(zip(names, params))
        P = _param_view(named)
        packed = {}
        _PACKED = packed
        try:
            out, sv = lifter_forward(P, geometry, maps, kp2d, ref, scales)
        finally:
            _PACKED = None
        ctx.capf = (geometry, maps, names, P, sv, kp2d, ref, [tuple(p.shape) for p in params], packed)
        return out.view(geometry.B, 1, J, 3)

    @staticmethod
    def backward(ctx, dout):
        global _PACKED
        geometry, maps, names, P, sv, kp2d, ref, shapes, packed = ctx.capf
        _PACKED = packed
        try:
            G = lifter_backward(P, geometry, maps, kp2d, ref, sv, dout.contiguous().view(-1, 3).float())
        finally:
            _PACKED = None
        for n in list(G):
            if n.endswith(".ow.weight"):
                q = n[: -len(".ow.weight")]
                G[q + ".attention_weights.weight"], G[q + ".sampling_offsets.weight"] = G[n][:16], G[n][16:]
                G[q + ".attention_weights.bias"], G[q + ".sampling_offsets.bias"] = G[q + ".ow.bias"][:16], G[q + ".ow.bias"][16:]
        grads = [G[n].reshape(s).contiguous() for n, s in zip(names, shapes)]
        return (None, None, None, None, None, None) + tuple(grads)


def forward_train(model, images, kp2d, ref, scales="draw"):
    """CA_PF.forward for a training step: frozen backbone through its inference plan, lifter through LifterFunction.
    `ref` is the already normalised crop tensor.  scales: "draw" (reference DropPath), None (no DropPath), or a dict from
    draw_drop_path_scales()."""
    from . import program
    B, H, W, _ = images.shape
    dev = images.device
    from .mvn.models._runtime import default_use_tc, state_version
    key = ("backbone-only", B, H, W, model.precision, dev.index or 0)
    ent = model._plans.get(key)
    ver = state_version(model.backbone)
    if ent is None:
        state = model.state_dict()
        shapes = {k: tuple(v.shape) for k, v in state.items()}
        prog = program.build_forward_program(model.backbone_type, getattr(model.backbone, "cfg", None), None, shapes, B, H, W,
                                             model.precision, use_tc=default_use_tc(), backbone_only=True)
        ent = [program.Plan(prog, state, dev), ver]
        model._plans[key] = ent
    elif ent[1] != ver:
        ent[0].repack(model.state_dict())
        ent[1] = ver
    plan = ent[0]
    plan.tensor(plan.prog.inputs["images"]).copy_(images)
    plan.run()
    maps = [plan.tensor(f) for f in plan.prog.feature_maps]
    vn = model.volume_net
    g = Geometry(B, vn.embed_dim_ratio, vn.feature_dim_list, [(m.shape[1], m.shape[2]) for m in maps], maps[0].dtype)
    if scales == "draw":
        scales = draw_drop_path_scales(B, dev, vn.levels, vn.drop_path_rate) if vn.drop_path_rate > 0 else None
    names, params = zip(*[(n, p) for n, p in vn.named_parameters()])
    return LifterFunction.apply(g, maps, scales, names, kp2d.reshape(-1, 2).contiguous().float(), ref.reshape(-1, 2).contiguous().float(), *params)


# ------------------------------------------------------------------------------------------------------
# AdamW as one kernel over a flat buffer
# ------------------------------------------------------------------------------------------------------
class FusedAdamW:
    """``torch.optim.AdamW(params, lr, betas, eps, weight_decay)`` semantics (train.py:337-345: weight_decay 0.1, defaults
    otherwise) with the update of ALL parameters as one libcapf_b200 kernel.  The parameters' storage and their ``.grad``
    become views of two flat fp32 buffers; ``param_groups[0]['lr']`` may be changed between steps like train.py:411-413 does."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.params = [p for p in params if p.requires_grad]
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 for p in self.params):
            raise lib.CapfError("FusedAdamW needs fp32 CUDA parameters (there is no CPU path)")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p, self.flat_g = _f32(n, dev=dev), torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat_g), torch.zeros_like(self.flat_g)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                p.grad = self.flat_g[off:off + k].view(p.shape)
                off += k
        self.param_groups = [{"params": self.params, "lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay}]
        self.step_count = 0
        self._hp = _f32(8, dev=dev)

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()

    def _gather_grads(self):
        off = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None:
                self.flat_g[off:off + k].zero_()
                p.grad = self.flat_g[off:off + k].view(p.shape)
            elif p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                self.flat_g[off:off + k].copy_(p.grad.reshape(-1))
                p.grad = self.flat_g[off:off + k].view(p.shape)
            off += k

    @torch.no_grad()
    def step(self):
        self._gather_grads()
        g = self.param_groups[0]
        self.step_count += 1
        b1, b2 = g["betas"]
        hp = [g["lr"], b1, b2, g["eps"], g["weight_decay"], 1.0 - b1 ** self.step_count, math.sqrt(1.0 - b2 ** self.step_count), 0.0]
        self._hp.copy_(torch.tensor(hp, dtype=torch.float32), non_blocking=True)
        n = self.flat_p.numel()
        _run(lib.OP_ADAMW, lib.F32, lib.F32, [n & 0x7fffffff, n >> 31], [], [self._hp, self.flat_g, self.exp_avg_sq], [self.flat_p, self.exp_avg],
             self.flat_p.device)

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
