"""TEST-ONLY reference interpreter for ``Program``s (torch-CPU semantics of every op in include/capf_b200.h).

It executes exactly the op list, packed weights and *memory plan* the GPU executes, so the CPU suite can prove
host logic (graph construction, BN folding, weight layouts, buffer aliasing) against the oracle without a GPU.
It is not part of the product and is never imported by it.
"""
import numpy as np
import torch
import torch.nn.functional as F

from capf_b200 import lib, program
from capf_b200.program import Buf, WSlot

import capf_oracle


class Interp:
    def __init__(self, prog, state):
        self.prog = prog
        self.mem = program.BufferStore(prog, "cpu")
        self.w = {k: w.pack(state) for k, w in program.weight_slots(prog).items()}

    def t(self, b):
        if b is None:
            return None
        if isinstance(b, WSlot):
            return self.w[id(b)]
        return self.mem.tensor(b)

    def run(self):
        for op in self.prog.ops:
            getattr(self, "_op%d" % op.kind)(op)

    # ---- ops -------------------------------------------------------------------------------------------
    def _op1(self, op):   # CONV2D
        N, H, W, Cin, Cout, KH, KW, stride, pad, Ho, Wo, act, impl = op.i[:13]
        if len(op.i) > 18 and op.i[18] == 1:      # split operands (bf16x3): hi * Wh + lo * Wh + hi * Wl, fp32 accumulation
            xs = self.t(op.ins[0]).reshape(N, H, W, 2 * Cin).float()
            hi, lo = xs[..., :Cin].permute(0, 3, 1, 2), xs[..., Cin:].permute(0, 3, 1, 2)
            wp = self.t(op.ins[1]).float().reshape(Cout, -1)
            T = KH * KW
            wh = wp[:, :2 * T * Cin].reshape(Cout, KH, KW, 2, Cin)
            assert torch.equal(wh[:, :, :, 0], wh[:, :, :, 1])
            wh = wh[:, :, :, 0].permute(0, 3, 1, 2).contiguous()
            wl = wp[:, 2 * T * Cin:].reshape(Cout, KH, KW, Cin).permute(0, 3, 1, 2).contiguous()
            y = (F.conv2d((hi + lo).double(), wh.double(), None, stride, pad) + F.conv2d(hi.double(), wl.double(), None, stride, pad))
            y = y.float().permute(0, 2, 3, 1)
        else:
            x = self.t(op.ins[0]).reshape(N, H, W, Cin).float().permute(0, 3, 1, 2)
            w = self.t(op.ins[1]).float()
            cin2 = op.i[19] if len(op.i) > 19 else 0
            if cin2:                                  # second A operand: out = [x | x2] . w^T (1x1 only)
                x2 = self.t(op.ins[5]).reshape(N, H, W, cin2).float().permute(0, 3, 1, 2)
                x = torch.cat([x, x2], dim=1)
                Cin = Cin + cin2
            if impl == lib.IMPL_TCGEN05:
                w = w.reshape(Cout, KH, KW, Cin).permute(0, 3, 1, 2)
            else:
                w = w.reshape(KH, KW, Cin, Cout).permute(3, 2, 0, 1)
            y = F.conv2d(x, w.contiguous(), None, stride, pad).permute(0, 2, 3, 1)
        if op.ins[2] is not None:
            y = y + self.t(op.ins[2])
        if act == lib.ACT_GELU:
            y = F.gelu(y)
        if op.ins[3] is not None:
            y = y + self.t(op.ins[3]).reshape(N, Ho, Wo, Cout).float()
        nseg = op.i[20] if len(op.i) > 20 else 0
        if nseg > 1:                                  # output segments: channel ranges of the GEMM go to separate tensors
            widths = [op.i[21 + s] for s in range(nseg - 1)]
            widths.append(Cout - sum(widths))
            c0 = 0
            for s, wd in enumerate(widths):
                ys = y[..., c0:c0 + wd]
                if act == lib.ACT_RELU and not (op.i[14] >> s) & 1:
                    ys = F.relu(ys)
                out = self.t(op.outs[s])
                out.copy_(ys.reshape(out.shape).to(out.dtype))
                c0 += wd
            return
        if act == lib.ACT_RELU:
            y = F.relu(y)
        out = self.t(op.outs[0])
        out.copy_(y.reshape(out.shape).to(out.dtype))

    def _op2(self, op):   # FUSE_SUM
        N, H, W, C, nt = op.i[:5]
        acc = None
        for k in range(nt):
            s = op.i[5 + k]
            t = self.t(op.ins[k]).reshape(N, H >> s, W >> s, C).float()
            if s:
                t = t.repeat_interleave(1 << s, 1).repeat_interleave(1 << s, 2)
            acc = t if acc is None else acc + t
        if op.i[9]:
            acc = F.relu(acc)
        out = self.t(op.outs[0])
        out.copy_(acc.to(out.dtype))

    def _op3(self, op):   # MAXPOOL
        N, H, W, C, Ho, Wo = op.i[:6]
        x = self.t(op.ins[0]).float().permute(0, 3, 1, 2)
        out = self.t(op.outs[0])
        out.copy_(F.max_pool2d(x, 3, 2, 1).permute(0, 2, 3, 1).to(out.dtype))

    def _op4(self, op):   # BILINEAR
        N, H, W, C, Ho, Wo = op.i[:6]
        x = self.t(op.ins[0]).float().permute(0, 3, 1, 2)
        out = self.t(op.outs[0])
        out.copy_(F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=True).permute(0, 2, 3, 1).to(out.dtype))

    def _op5(self, op):   # LAYERNORM
        rows, D, period = op.i[:3]
        x = self.t(op.ins[0]).reshape(rows, D)
        if period:
            x = x + self.t(op.ins[3]).reshape(period, D).repeat(rows // period, 1)
        y = F.layer_norm(x, (D,), self.t(op.ins[1]), self.t(op.ins[2]), op.f[0])
        out = self.t(op.outs[0])
        if len(op.i) > 3 and op.i[3]:        # fused projection (the head: LayerNorm + Linear(D -> n), pose_dformer.py:205-208)
            n = op.i[3]
            y = y @ self.t(op.ins[4]).reshape(n, D).t() + self.t(op.ins[5]).reshape(n)
        out.copy_(y.reshape(out.shape).to(out.dtype))

    def _op6(self, op):   # ATTENTION
        groups, seq, heads, hd, ts, gs = op.i[:6]
        D = heads * hd
        qkv = self.t(op.ins[0]).float()
        rows = (torch.arange(groups).view(-1, 1) * gs + torch.arange(seq).view(1, -1) * ts)      # [G, seq]
        x = qkv[rows.reshape(-1)].view(groups, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
        a = ((x[0] @ x[1].transpose(-2, -1)) * op.f[0]).softmax(-1)
        y = (a @ x[2]).transpose(1, 2).reshape(groups * seq, D)
        out = self.t(op.outs[0])
        out[rows.reshape(-1)] = y.to(out.dtype)

    def _maps(self, op):
        B, Jn, nl = op.i[:3]
        geo = [(op.i[3 + 3 * l], op.i[4 + 3 * l], op.i[5 + 3 * l]) for l in range(nl)]
        maps = [self.t(op.ins[1 + l]).float().permute(0, 3, 1, 2) for l in range(nl)]
        return B, Jn, nl, geo, maps, [op.i[15 + l] for l in range(nl)]

    def _op7(self, op):   # REF_SAMPLE
        B, Jn, nl, geo, maps, offs = self._maps(op)
        ref = self.t(op.ins[0]).view(B, Jn, 2)
        out = self.t(op.outs[0])
        for l in range(nl):
            s = F.grid_sample(maps[l], ref.unsqueeze(-2), align_corners=True).squeeze(-1).permute(0, 2, 1)
            C = geo[l][2]
            out[offs[l]:offs[l] + B * Jn * C] = s.reshape(-1).to(out.dtype)
            if op.outs[1] is not None:
                x0, y0, m, _ = capf_oracle.grid_sample_records(ref.numpy(), geo[l][0], geo[l][1], border=False)
                rec = self.t(op.outs[1])
                rec[l, :, 0] = torch.from_numpy(x0.reshape(-1))
                rec[l, :, 1] = torch.from_numpy(y0.reshape(-1))
                rec[l, :, 2] = torch.from_numpy(m.reshape(-1).astype(np.int32))
                rec[l, :, 3:] = 0
                rec[l, :, 4:6] = torch.from_numpy(ref.numpy().reshape(-1, 2).view(np.int32))

    def _op8(self, op):   # DEFORM_SAMPLE
        B, Jn, nl, geo, maps, offs = self._maps(op)
        R = B * Jn
        ref = self.t(op.ins[0]).view(B, 1, Jn, 1, 2)
        ow = self.t(op.ins[5]).view(nl, B, Jn, 48)
        wts = ow[..., :16].reshape(nl, B, Jn, 4, 4).softmax(-1)
        pos = ow[..., 16:].reshape(nl, B, Jn, 16, 2).tanh().permute(1, 0, 2, 3, 4) + ref      # [B,nl,J,16,2]
        out = self.t(op.outs[0])
        for l in range(nl):
            C = geo[l][2]
            s = F.grid_sample(maps[l], pos[:, l], padding_mode="border", align_corners=True).permute(0, 2, 3, 1)  # [B,J,16,C]
            g = (s.reshape(B, Jn, 4, 4, C) * wts[l].unsqueeze(-1)).sum(-2)                                      # [B,J,4,C]
            out[offs[l]:offs[l] + R * 4 * C] = g.reshape(-1).to(out.dtype)
            if op.outs[1] is not None:
                x0, y0, m, _ = capf_oracle.grid_sample_records(pos[:, l].numpy(), geo[l][0], geo[l][1], border=True)
                rec = self.t(op.outs[1])
                rec[l, :, :, 0] = torch.from_numpy(x0.reshape(R, 16))
                rec[l, :, :, 1] = torch.from_numpy(y0.reshape(R, 16))
                rec[l, :, :, 2] = torch.from_numpy(m.reshape(R, 16).astype(np.int32))
                rec[l, :, :, 3:] = 0
                rec[l, :, :, 4:6] = torch.from_numpy(np.ascontiguousarray(pos[:, l].numpy()).reshape(R, 16, 2).view(np.int32))

    def _op9(self, op):   # EMBED_COORD
        B, Jn, D, S = op.i[:4]
        kp = self.t(op.ins[0]).view(B * Jn, 2)
        W, b, pos = self.t(op.ins[1]).view(D, 2), self.t(op.ins[2]), self.t(op.ins[3]).view(S, Jn, D)
        X = self.t(op.outs[0])
        X.copy_(pos.unsqueeze(1).expand(S, B, Jn, D).reshape(S, B * Jn, D))
        X[0] += F.linear(kp, W, b)

    def _op10(self, op):  # LEVELS_TO_JOINT
        R, S, D = op.i[:3]
        self.t(op.outs[0]).copy_(self.t(op.ins[0]).view(S, R, D).permute(1, 0, 2).reshape(R, S * D))

    def _op11(self, op):  # CROP_NORMALIZE
        c = self.t(op.outs[0]).view(-1, 2)
        c /= torch.tensor([96.0, 128.0])
        c -= 1.0

    def _op14(self, op):  # BASICBLOCK: relu(conv2(relu(conv1(x) + b1)) + b2 + x), 16-bit intermediate like the two-op form
        N, H, W, C = op.i[:4]
        x = self.t(op.ins[0]).reshape(N, H, W, C)
        xf = x.float().permute(0, 3, 1, 2)
        w1 = self.t(op.ins[1]).float().reshape(C, 3, 3, C).permute(0, 3, 1, 2).contiguous()
        w2 = self.t(op.ins[3]).float().reshape(C, 3, 3, C).permute(0, 3, 1, 2).contiguous()
        # the same expressions, in the same order, as the two CONV2D ops this op replaces (bias added after the convolution)
        u = F.relu(F.conv2d(xf, w1, None, 1, 1).permute(0, 2, 3, 1) + self.t(op.ins[2])).to(x.dtype).float().permute(0, 3, 1, 2)
        y = F.relu(F.conv2d(u, w2, None, 1, 1).permute(0, 2, 3, 1) + self.t(op.ins[4]) + x.float())
        out = self.t(op.outs[0])
        out.copy_(y.reshape(out.shape).to(out.dtype))

    def _op28(self, op):  # MLP: X = X + gelu(t W1^T + b1) W2^T + b2, the expressions of the two CONV2D ops it replaces (hidden rounded to dtype_in)
        rows, k1, n1, n2 = op.i[:4]
        t = self.t(op.ins[0]).reshape(rows, k1)
        w1 = self.t(op.ins[1]).float().reshape(n1, k1)
        w2 = self.t(op.ins[4]).float().reshape(n2, n1)
        h = F.conv2d(t.float().t().reshape(1, k1, rows, 1), w1.reshape(n1, k1, 1, 1)).reshape(n1, rows).t()
        if op.ins[2] is not None:
            h = h + self.t(op.ins[2])
        h = F.gelu(h).to(t.dtype)
        y = F.conv2d(h.float().t().reshape(1, n1, rows, 1), w2.reshape(n2, n1, 1, 1)).reshape(n2, rows).t()
        if op.ins[5] is not None:
            y = y + self.t(op.ins[5])
        y = y + self.t(op.ins[3]).reshape(rows, n2).float()
        out = self.t(op.outs[0])
        out.copy_(y.reshape(out.shape).to(out.dtype))

    def _op27(self, op):  # EXPAND_REDUCE: y = relu(t W3^T + b3 + x); u = relu(y W1^T + b1), the expressions of the two CONV2D ops it replaces
        rows, k1, n1, n2 = op.i[:4]
        t = self.t(op.ins[0]).reshape(rows, k1).float()
        x = self.t(op.ins[3]).reshape(rows, n1).float()
        w3 = self.t(op.ins[1]).float().reshape(n1, k1)
        w1 = self.t(op.ins[4]).float().reshape(n2, n1)
        y = F.conv2d(t.t().reshape(1, k1, rows, 1), w3.reshape(n1, k1, 1, 1)).reshape(n1, rows).t()
        if op.ins[2] is not None:
            y = y + self.t(op.ins[2])
        y = F.relu(y + x)
        yo = self.t(op.outs[0])
        yo.copy_(y.reshape(yo.shape).to(yo.dtype))
        u = F.conv2d(yo.reshape(rows, n1).float().t().reshape(1, n1, rows, 1), w1.reshape(n2, n1, 1, 1)).reshape(n2, rows).t()
        if op.ins[5] is not None:
            u = u + self.t(op.ins[5])
        uo = self.t(op.outs[1])
        uo.copy_(F.relu(u).reshape(uo.shape).to(uo.dtype))

    def _op12(self, op):  # CAST
        out = self.t(op.outs[0])
        if len(op.i) > 2 and op.i[2] > 0:         # split planes: rows of C fp32 -> [hi (C) | lo (C)] bf16
            C = op.i[2]
            x = self.t(op.ins[0]).reshape(-1, C).float()
            hi = x.to(torch.bfloat16)
            lo = (x - hi.float()).to(torch.bfloat16)
            out.copy_(torch.cat([hi, lo], dim=1).reshape(out.shape))
            return
        out.copy_(self.t(op.ins[0]).to(out.dtype))
