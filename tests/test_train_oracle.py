"""f2 groundwork: the parity gate a volume_net training step will be held to.  The oracle's autograd
gradients and its AdamW restatement against what the unmodified reference produced under autograd + torch.optim.AdamW
(tests/golden/grad_hrnet32_b2_128x96.npz, oracle/gen_golden_grad.py).  CPU only: no backward kernel exists yet."""
import os

import numpy as np
import torch

import capf_b200
import capf_oracle
import protocol
from gen_golden_grad import CASE, LR, make_target, positions

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "grad_hrnet32_b2_128x96.npz")


def test_oracle_gradients_and_adamw_step_match_reference():
    g = np.load(GOLD)
    backbone, B, H, W, wseed, iseed = CASE
    cfg = capf_b200.make_config(backbone)
    model = capf_b200.CA_PF(cfg, precision="fp32")
    sd = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], wseed)
    images, kp2d, crop = protocol.make_inputs(B, H, W, iseed)
    loss, grads = capf_oracle.volume_net_loss_and_grads(sd, backbone, cfg.model.backbone, images, kp2d, crop, make_target(B, 99))
    assert abs(loss - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    names = [str(n) for n in g["names"]]
    assert len(names) == 191 and all(("volume_net." + n) in grads for n in names)
    worst = 0.0
    for k, n in enumerate(names):
        gr = grads["volume_net." + n].reshape(-1).double()
        pos = positions(n, gr.numel())
        ref_norm = float(g[f"g{k}_norm"])
        assert abs(float(gr.norm()) - ref_norm) <= 2e-4 * ref_norm + 1e-9, (n, float(gr.norm()), ref_norm)
        err = np.abs(gr[pos].numpy() - g[f"g{k}_samples"]).max() / (ref_norm / max(1.0, gr.numel() ** 0.5) + 1e-12)
        worst = max(worst, err)
        # one optimiser step from the reference's gradient samples: torch.optim.AdamW == the restatement
        p0 = sd["volume_net." + n].reshape(-1).double()[pos]
        gs = torch.from_numpy(g[f"g{k}_samples"])
        p1, _, _ = capf_oracle.adamw_step(p0, gs, torch.zeros_like(p0), torch.zeros_like(p0), 1, LR)
        np.testing.assert_allclose(p1.numpy(), g[f"p{k}_after"], rtol=0, atol=2e-7)
    assert worst < 5e-2, worst          # sampled elements agree to a few percent of the parameter's RMS gradient


def test_oracle_droppath_step_matches_reference_train_mode():
    """Train-mode step (train.py:145-148: volume_net.train(), DropPath rates linspace(0, 0.2, 4) live).  The masks the product draws
    (capf_b200.train.draw_drop_path_scales: same torch calls, order and shapes as the reference's DropPath modules) under the seed
    the fixture was generated with reproduce the reference's loss and gradients through the oracle."""
    from gen_golden_grad import TRAIN_SEED
    g = np.load(os.path.join(os.path.dirname(GOLD), "grad_train_hrnet32_b2_128x96.npz"))
    backbone, B, H, W, wseed, iseed = CASE
    cfg = capf_b200.make_config(backbone)
    model = capf_b200.CA_PF(cfg, precision="fp32")
    sd = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], wseed)
    images, kp2d, crop = protocol.make_inputs(B, H, W, iseed)
    torch.manual_seed(TRAIN_SEED)
    drop = capf_b200.train.draw_drop_path_scales(B, "cpu", 4, model.volume_net.drop_path_rate)
    assert drop["context_blocks"][0] == (None, None) and drop["res_blocks"][1][0].shape == (B * 17,) and drop["joint_blocks"][3][1].shape == (B,)
    loss, grads = capf_oracle.volume_net_loss_and_grads(sd, backbone, cfg.model.backbone, images, kp2d, crop, make_target(B, 99), drop=drop)
    assert abs(loss - float(g["loss"])) < 1e-5 * abs(float(g["loss"])), (loss, float(g["loss"]))
    eval_loss = float(np.load(GOLD)["loss"])
    assert abs(float(g["loss"]) - eval_loss) > 1e-3          # the masks did something
    for k, n in enumerate(str(n) for n in g["names"]):
        gr = grads["volume_net." + n].reshape(-1).double()
        ref_norm = float(g[f"g{k}_norm"])
        assert abs(float(gr.norm()) - ref_norm) <= 2e-4 * ref_norm + 1e-9, (n, float(gr.norm()), ref_norm)
