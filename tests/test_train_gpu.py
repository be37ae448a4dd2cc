"""GPU: the training step of volume_net -- kernels of csrc/capf_train.cu through the C ABI against
torch autograd, and the whole step (CA_PF.forward under autograd -> loss.backward() -> AdamW) against the fixtures the
unmodified reference produced (tests/golden/grad_*.npz) and the CPU oracle's autograd on the same inputs."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import capf_b200
import capf_oracle
import protocol
from capf_b200 import lib, train
from conftest import rel_l2
from gen_golden_grad import CASE, LR, TRAIN_SEED, make_target, positions

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# ---- kernels ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(300, 48, 128), (37, 3, 640), (4352, 1920, 640), (128, 128, 21760), (65, 70, 33)])
@pytest.mark.parametrize("tc", [False, True], ids=["cuda-core", "tensor-core"])
def test_linear_forward_backward_all_transposes_and_split_k(M, N, K, tc, monkeypatch):
    """y = x W^T + b, dx = dy W, dW = dy^T x (split over the rows, fixed-order reduce), db.  cuda-core: the fp32 GEMM of
    capf_train.cu for all three; tensor-core: forward and dgrad through the split-bf16 tcgen05 kernel where the shape allows
    (K, N % 16 == 0), wgrad stays fp32 -- fp32-class either way."""
    monkeypatch.setattr(train, "USE_TC", tc)
    x, W, dy = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(M, N, seed=3)
    b = rnd(N, seed=4)
    tol = 3e-5 * max(1.0, (max(K, N) / 1280.0) ** 0.5) if tc else 2e-6     # split-bf16 operands: 2^-16 per product, random-walk over the depth
    y = train.linear(x, W, b, M)
    assert rel_l2(y.cpu(), (x.double() @ W.double().t() + b.double()).float().cpu()) < tol
    dx, dW, db = train.linear_bwd(x, W, dy, M)
    assert rel_l2(dx.cpu(), (dy.double() @ W.double()).float().cpu()) < tol
    assert rel_l2(dW.cpu(), (dy.double().t() @ x.double()).float().cpu()) < 2e-6
    assert rel_l2(db.cpu(), dy.double().sum(0).float().cpu()) < 2e-6
    acc = y.clone()
    train.linear(x, W, b, M, out=acc, accumulate=True)
    assert rel_l2(acc.cpu(), (2 * y).cpu()) < 1e-6
    dW2 = train.linear_bwd(x, W, dy, M)[1]
    assert torch.equal(dW, dW2), "fixed summation order: bit-reproducible"


@pytest.mark.parametrize("rows,D,period", [(1000, 128, 0), (85 * 4, 128, 85), (300, 640, 0)])
def test_layernorm_backward_matches_autograd(rows, D, period):
    x = rnd(rows, D, seed=5).requires_grad_(True)
    x0 = rnd(period, D, seed=6).requires_grad_(True) if period else None
    gam, bet = (1 + 0.1 * rnd(D, seed=7)).requires_grad_(True), rnd(D, seed=8).requires_grad_(True)
    dy = rnd(rows, D, seed=9)
    z = x + (x0.repeat(rows // period, 1) if period else 0)
    F.layer_norm(z, (D,), gam, bet, 1e-5).backward(dy)
    dx = torch.ones(rows, D, device=DEV)
    dg, db = train.layernorm_bwd(x.detach(), rows, D, gam.detach(), 1e-5, dy, dx, True, x0=None if x0 is None else x0.detach(), period=period)
    assert rel_l2((dx - 1).cpu(), x.grad.cpu()) < 1e-5 and rel_l2(dg.cpu(), gam.grad.cpu()) < 1e-5 and rel_l2(db.cpu(), bet.grad.cpu()) < 1e-5
    if period:
        assert rel_l2((dx - 1).view(rows // period, period, D).sum(0).cpu(), x0.grad.cpu()) < 1e-5


def test_gelu_and_rows_axpy():
    h, dy = rnd(64, 256, seed=1).requires_grad_(True), rnd(64, 256, seed=2)
    F.gelu(h).backward(dy)
    assert rel_l2(train.gelu(h.detach()).cpu(), F.gelu(h.detach()).cpu()) < 1e-6
    assert rel_l2(train.gelu_bwd(h.detach(), dy).cpu(), h.grad.cpu()) < 1e-6
    t, y, s = rnd(5 * 34, 128, seed=3), rnd(5 * 34, 128, seed=4), rnd(2, seed=5)
    want = y + s[(torch.arange(5 * 34, device=DEV) % 34) // 17].view(-1, 1) * t
    assert rel_l2(train.rows_axpy(t, y.clone(), 5 * 34, 128, s, mod=34, div=17).cpu(), want.cpu()) < 1e-6


@pytest.mark.parametrize("groups,seq,heads,hd,ts,gs", [(34, 5, 8, 16, 34, 1), (3, 17, 8, 80, 1, 17)])
def test_attention_backward_matches_autograd(groups, seq, heads, hd, ts, gs):
    rows, D = groups * seq, heads * hd
    qkv = rnd(rows, 3 * D, seed=1).requires_grad_(True)
    dout = rnd(rows, D, seed=2)
    idx = (torch.arange(groups).view(-1, 1) * gs + torch.arange(seq).view(1, -1) * ts).reshape(-1).to(DEV)
    x = qkv[idx].view(groups, seq, 3, heads, hd).permute(2, 0, 3, 1, 4)
    a = ((x[0] @ x[1].transpose(-2, -1)) * hd ** -0.5).softmax(-1)
    y = torch.zeros(rows, D, device=DEV).index_copy(0, idx, (a @ x[2]).transpose(1, 2).reshape(groups * seq, D))
    y.backward(dout)
    got = train.attention(qkv.detach(), rows, D, heads, groups, seq, ts, gs)
    assert rel_l2(got.cpu(), y.detach().cpu()) < 1e-5
    dq = train.attention_bwd(qkv.detach(), dout, rows, D, heads, groups, seq, ts, gs)
    assert rel_l2(dq.cpu(), qkv.grad.cpu()) < 1e-5


@pytest.mark.parametrize("map_dtype", [torch.float32, torch.float16])
def test_deform_backward_matches_autograd(map_dtype):
    """grid_sample(padding_mode='border', align_corners=True) w.r.t. its grid, tanh and softmax (pose_dformer.py:124-135)."""
    B, dims, hw = 3, [32, 64, 128, 256], [(16, 12), (8, 6), (4, 3), (2, 2)]
    g = train.Geometry(B, 128, dims, hw, map_dtype)
    R = g.R
    maps = [rnd(B, h, w, c, seed=10 + l).to(map_dtype) for l, ((h, w), c) in enumerate(zip(hw, dims))]
    ref = (torch.rand(R, 2, generator=torch.Generator().manual_seed(1)) * 2.6 - 1.3).to(DEV)        # some points outside: border clip
    ow = rnd(4 * R, 48, seed=2).requires_grad_(True)
    dg = rnd(g.goffs[-1], seed=3)
    # torch composite of the forward
    wts = ow.view(4, B, 17, 48)[..., :16].reshape(4, B, 17, 4, 4).softmax(-1)
    pos = ow.view(4, B, 17, 48)[..., 16:].reshape(4, B, 17, 16, 2).tanh() + ref.view(1, B, 17, 1, 2)
    outs = []
    for l in range(4):
        s = F.grid_sample(maps[l].float().permute(0, 3, 1, 2), pos[l], padding_mode="border", align_corners=True).permute(0, 2, 3, 1)
        outs.append((s.reshape(B, 17, 4, 4, -1) * wts[l].unsqueeze(-1)).sum(-2).reshape(-1))
    torch.cat(outs).backward(dg)
    dow = torch.empty(4 * R, 48, device=DEV)
    mdt = train._DT[map_dtype]
    train._run(lib.OP_DEFORM_BWD, mdt, lib.F32, [B, 17, 4] + g.map_geo + g.goffs[:4], [], [ref] + maps + [ow.detach()], [dow, dg], DEV)
    assert rel_l2(dow.cpu(), ow.grad.cpu()) < 2e-5


# ---- the whole step ------------------------------------------------------------------------------------------------
def training_model(precision="fp32", drop_path_rate=None):
    backbone, B, H, W, wseed, iseed = CASE
    cfg = capf_b200.make_config(backbone)
    model = capf_b200.CA_PF(cfg, precision=precision)
    sd = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], wseed)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV)
    model.train()                      # train.py:145-148
    model.backbone.eval()
    model.volume_net.train()
    if drop_path_rate is not None:
        model.volume_net.drop_path_rate = drop_path_rate
    images, kp2d, crop = protocol.make_inputs(B, H, W, iseed)
    return model, sd, cfg, images, kp2d, crop, make_target(B, 99)


def check_against_fixture(model, loss, fixture, tol_norm):
    g = np.load(os.path.join(GOLD, fixture))
    assert abs(float(loss) - float(g["loss"])) < 2e-5 * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    named = dict(model.volume_net.named_parameters())
    worst = 0.0
    for k, n in enumerate(str(n) for n in g["names"]):
        gr = named[n].grad.reshape(-1).double().cpu()
        ref_norm = float(g[f"g{k}_norm"])
        assert abs(float(gr.norm()) - ref_norm) <= tol_norm * ref_norm + 1e-9, (n, float(gr.norm()), ref_norm)
        pos = positions(n, gr.numel())
        worst = max(worst, float(np.abs(gr[pos].numpy() - g[f"g{k}_samples"]).max() / (ref_norm / max(1.0, gr.numel() ** 0.5) + 1e-12)))
    return g, worst


@pytest.mark.parametrize("tc", [False, True], ids=["cuda-core-gemms", "tensor-core-gemms"])
def test_training_step_matches_reference_fixture_and_oracle(tc, monkeypatch):
    """Both GEMM paths of the step (CAPF_TRAIN_TC): all-fp32 CUDA-core GEMMs hold the reference to 2e-4 per tensor; with the forward /
    dgrad GEMMs on the tensor cores (split bf16 operands, 2^-16) the bound is 1e-3 -- the sampling-offset gradients amplify a
    1e-5 change of the sampled positions the most (observed 2.1e-4).
    model(images, kp, crop) under autograd -> MPJPE -> backward, blocks without DropPath (the eval-mode fixture): loss, all 191
    gradients vs the reference fixture AND element-wise vs the CPU oracle's autograd; then one optimiser step (torch.optim.AdamW
    on our gradients) vs the parameters the reference ended with."""
    monkeypatch.setattr(train, "USE_TC", tc)
    tol = 1e-3 if tc else 2e-4
    model, sd, cfg, images, kp2d, crop, gt = training_model("fp32", drop_path_rate=0.0)
    opt = torch.optim.AdamW([{"params": [p for p in model.volume_net.parameters() if p.requires_grad], "lr": LR}], weight_decay=0.1)
    crop_d = crop.to(DEV)
    pred = model(images.to(DEV), kp2d.to(DEV), crop_d)
    assert pred.requires_grad and pred.shape == (CASE[1], 1, 17, 3)
    c2 = crop.clone()
    capf_oracle.normalize_crop_(c2)
    assert torch.equal(crop_d.cpu(), c2), "conpose.py:34-35 in-place normalisation also in a training step"
    loss = torch.mean(torch.norm(pred - gt.to(DEV), dim=3))
    opt.zero_grad()
    loss.backward()
    g, worst = check_against_fixture(model, loss, "grad_hrnet32_b2_128x96.npz", tol)
    assert worst < 5e-2
    oloss, ograds = capf_oracle.volume_net_loss_and_grads(sd, CASE[0], cfg.model.backbone, images, kp2d, crop, gt)
    for n, p in model.volume_net.named_parameters():
        e = rel_l2(p.grad.cpu(), ograds["volume_net." + n])
        assert e < (5e-3 if tc else 2e-4), (n, e)
    assert all(p.grad is None for p in model.backbone.parameters())
    grads_before = {n: p.grad.detach().cpu().clone() for n, p in model.volume_net.named_parameters()}
    opt.step()
    named = dict(model.volume_net.named_parameters())
    for k, n in enumerate(str(n) for n in g["names"]):
        pos = positions(n, named[n].numel())
        # the first AdamW step moves an element by lr * g / (|g| + eps): only where |g| >> eps = 1e-8 is that insensitive to a
        # 1e-4 relative difference of g -- compare those elements with what the reference ended with, and every element with
        # the AdamW restatement applied to OUR gradient
        got = named[n].detach().reshape(-1)[pos].double().cpu().numpy()
        gref = g[f"g{k}_samples"]
        big = np.abs(gref) > 1e-6
        np.testing.assert_allclose(got[big], g[f"p{k}_after"][big], rtol=0, atol=1e-6)
        p0 = sd["volume_net." + n].reshape(-1).double()[pos]
        ours = grads_before[n].reshape(-1).double()[pos]
        want, _, _ = capf_oracle.adamw_step(p0, ours, torch.zeros_like(p0), torch.zeros_like(p0), 1, LR)
        np.testing.assert_allclose(got, want.numpy(), rtol=0, atol=2e-7)


def test_training_step_with_droppath_matches_reference_train_mode():
    """DropPath live: the masks the reference drew under TRAIN_SEED (same torch calls on the CPU generator), handed to the kernels."""
    model, sd, cfg, images, kp2d, crop, gt = training_model("fp32")
    torch.manual_seed(TRAIN_SEED)
    drop = train.draw_drop_path_scales(CASE[1], "cpu", 4, model.volume_net.drop_path_rate)
    drop_d = {k: [tuple(None if s is None else s.to(DEV) for s in pair) for pair in v] for k, v in drop.items()}
    ref = capf_oracle.normalize_crop_(crop.clone()).to(DEV)
    pred = train.forward_train(model, images.to(DEV), kp2d.to(DEV), ref, scales=drop_d)
    loss = torch.mean(torch.norm(pred - gt.to(DEV), dim=3))
    loss.backward()
    check_against_fixture(model, loss, "grad_train_hrnet32_b2_128x96.npz", 1e-3)
    # and the public call draws its own masks on the GPU generator: runs, differentiable, different from the eval-mode output
    for p in model.volume_net.parameters():
        p.grad = None
    torch.manual_seed(1)
    a = model(images.to(DEV), kp2d.to(DEV), crop.to(DEV))
    torch.manual_seed(1)
    b = model(images.to(DEV), kp2d.to(DEV), crop.to(DEV))
    assert torch.equal(a, b), "same seed, same masks"
    a.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.volume_net.parameters())


@pytest.mark.parametrize("precision", ["fp16", "bf16x3"])
def test_training_step_on_a_tensor_core_backbone(precision):
    """The frozen backbone may run in any inference precision; the lifter trains in fp32 on its feature maps."""
    model, sd, cfg, images, kp2d, crop, gt = training_model(precision, drop_path_rate=0.0)
    pred = model(images.to(DEV), kp2d.to(DEV), crop.to(DEV))
    torch.mean(torch.norm(pred - gt.to(DEV), dim=3)).backward()
    oloss, ograds = capf_oracle.volume_net_loss_and_grads(sd, CASE[0], cfg.model.backbone, images, kp2d, crop, gt)
    tol = 2e-2 if precision == "fp16" else 5e-4
    errs = [rel_l2(p.grad.cpu(), ograds["volume_net." + n]) for n, p in model.volume_net.named_parameters()]
    assert float(np.median(errs)) < tol and max(errs) < 10 * tol, (float(np.median(errs)), max(errs))


def test_fused_adamw_matches_torch_adamw():
    """One kernel over the flat buffer == torch.optim.AdamW (train.py:337-345: weight_decay 0.1), several steps with an lr change."""
    g = torch.Generator().manual_seed(0)
    shapes = [(128, 2), (128,), (1, 5, 17, 128), (1920, 640), (3,)]
    ours = [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    fo = train.FusedAdamW(ours, lr=LR, weight_decay=0.1)
    to = torch.optim.AdamW([{"params": theirs, "lr": LR}], weight_decay=0.1)
    for step in range(4):
        fo.zero_grad()
        to.zero_grad()
        for p, q in zip(ours, theirs):
            gr = torch.randn(*p.shape, generator=g).to(DEV) * (10.0 ** (step - 2))
            (p * gr).sum().backward()
            (q * gr).sum().backward()
        fo.step()
        to.step()
        if step == 1:
            fo.param_groups[0]["lr"] = to.param_groups[0]["lr"] = LR * 0.99          # train.py:411-413
        for p, q in zip(ours, theirs):
            assert (p - q).abs().max() < 2e-7 * max(1.0, float(q.abs().max()))
    assert ours[0].grad.data_ptr() == fo.flat_g.data_ptr() and ours[0].data_ptr() == fo.flat_p.data_ptr()
