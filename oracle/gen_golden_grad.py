"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/grad_hrnet32_b2_128x96.npz: loss, volume_net gradients and the
parameters after one optimiser step, from the UNMODIFIED reference model under autograd + torch.optim.AdamW exactly as
train.py:186-201 / :337-345 drive it (MPJPE criterion, lr = volume_net_lr of human36m.yaml, weight_decay 0.1, no clipping).
Blocks run in eval mode (DropPath = identity, BatchNorm running stats; the backbone is frozen by CA_PF itself).
Per parameter the fixture keeps the gradient norm, its sum and 64 sampled elements (positions derived from the name).
`--train` writes grad_train_hrnet32_b2_128x96.npz instead: the same step with the model in the state train.py:145-148 puts it in
(volume_net.train(): DropPath with rates linspace(0, 0.2, 4) is live; masks fixed by torch.manual_seed(TRAIN_SEED)).
Run in the authoring container:  python oracle/gen_golden_grad.py [--train]"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import protocol  # noqa: E402
import ref_import  # noqa: E402

LR = 0.00064          # experiments/human36m/human36m.yaml:58
CASE = ("hrnet_32", 2, 128, 96, 0, 11)


def positions(name, numel, n=64):
    g = np.random.Generator(np.random.PCG64([0x6AD, zlib.crc32(name.encode()), numel % 9973]))
    return g.integers(0, numel, size=min(n, numel))


def make_target(B, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.randn(B, 1, 17, 3, generator=g) * 0.3
    gt[:, :, 0] = 0
    return gt


TRAIN_SEED = 4321     # torch.manual_seed right before the train-mode forward: fixes the DropPath masks


def main(train_mode=False):
    backbone, B, H, W, wseed, iseed = CASE
    torch.manual_seed(0)
    model = ref_import.build_reference_model(backbone)
    spec = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(protocol.make_weights(spec, wseed), strict=True)
    model.eval()
    if train_mode:          # train.py:145-148: model.train(); backbone.eval(); volume_net.train()  -> DropPath is live
        model.train()
        model.backbone.eval()
        model.volume_net.train()
        torch.manual_seed(TRAIN_SEED)
    images, kp2d, crop = protocol.make_inputs(B, H, W, iseed)
    gt = make_target(B, 99)
    from mvn.models.loss import MPJPE
    params = [(n, p) for n, p in model.volume_net.named_parameters() if p.requires_grad]
    opt = torch.optim.AdamW([{"params": [p for _, p in params], "lr": LR}], weight_decay=0.1)
    with torch.enable_grad():
        pred = model(images, kp2d, crop.clone())
        loss = MPJPE()(pred, gt)
        opt.zero_grad()
        loss.backward()
    out = {"loss": np.array(float(loss)), "lr": np.array(LR), "names": np.array([n for n, _ in params])}
    grads = {n: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for n, p in params}
    opt.step()
    for k, (n, p) in enumerate(params):
        g = grads[n].reshape(-1).double()
        pos = positions(n, g.numel())
        out[f"g{k}_norm"], out[f"g{k}_sum"] = np.array(float(g.norm())), np.array(float(g.sum()))
        out[f"g{k}_samples"] = g[pos].numpy()
        out[f"p{k}_after"] = p.detach().reshape(-1)[pos].double().numpy()
    name = "grad_train_hrnet32_b2_128x96.npz" if train_mode else "grad_hrnet32_b2_128x96.npz"
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", name), **out)
    print("loss", float(loss), "params", len(params), "total grad norm", float(torch.sqrt(sum(g.double().norm() ** 2 for g in grads.values()))))


if __name__ == "__main__":
    main(train_mode="--train" in sys.argv)
