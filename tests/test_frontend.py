"""Row f1 of SURVEY.md section 8: the uint8 pre-processing + flip-test front end (capf_b200.frontend) against the oracle's
restatement of data_prefetcher.preload (mvn/datasets/utils.py:33-82) and of the flip-test merge (train.py:170-181)."""
import numpy as np
import pytest
import torch

import capf_b200
import capf_oracle
import protocol
from capf_b200 import frontend


def _u8(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)


def test_oracle_frontend_restatement_properties():
    """CPU: the restated prefetcher is self-consistent -- mirroring commutes with the per-pixel transform, the keypoint
    mirror is an involution, and un-mirroring a mirrored prediction restores it (so merge(p, mirror(p)) == p)."""
    img = _u8(2, 8, 12, 1)
    kp = torch.rand(2, 17, 2) * 2 - 1
    crop = torch.rand(2, 17, 2) * torch.tensor([191.0, 255.0])
    for bb in ("hrnet_32", "cpn"):
        images, kps, crops = capf_oracle.prefetch_flip_test(img, kp, crop, bb)
        assert images.shape == (2, 2, 8, 12, 3) and images.dtype == torch.float32
        assert torch.equal(images[:, 1], torch.flip(capf_oracle.prefetch_images(img, bb), [2]))
        assert torch.equal(images[:, 1], capf_oracle.prefetch_images(torch.flip(img, [2]), bb))
        k2, c2 = frontend.flip_keypoints(kps[:, 1], crops[:, 1])
        assert torch.equal(k2, kp) and torch.allclose(c2, crop, atol=1e-4)
    assert frontend.JOINTS_LEFT == capf_oracle.JOINTS_LEFT and frontend.JOINTS_RIGHT == capf_oracle.JOINTS_RIGHT
    p = torch.randn(3, 1, 17, 3)
    pm = p.clone()
    pm[..., 0] *= -1
    pm[:, :, capf_oracle.JOINTS_LEFT + capf_oracle.JOINTS_RIGHT] = pm[:, :, capf_oracle.JOINTS_RIGHT + capf_oracle.JOINTS_LEFT]
    assert torch.allclose(capf_oracle.flip_test_merge(p, pm), p, atol=1e-7)
    assert torch.equal(frontend.merge_flip_test(p, pm), capf_oracle.flip_test_merge(p, pm))
    ms = frontend.normalisation_params("cpn", "cpu")
    assert torch.equal(ms[3:], torch.ones(3)) and abs(float(ms[0]) - 122.7717 / 255) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("backbone", ["hrnet_32", "cpn"])
@pytest.mark.parametrize("shape", [(3, 16, 12), (2, 9, 7), (1, 256, 192)], ids=str)
@pytest.mark.parametrize("mirror", [False, True], ids=["plain", "mirrored"])
def test_preprocess_kernel_is_bit_exact(backbone, shape, mirror):
    """CAPF_OP_PREPROCESS_U8 == (flip(images, [-1]) / 255 - mean) / std [+ flip along W], bit for bit (IEEE divisions),
    for widths that are and are not multiples of 4."""
    B, H, W = shape
    img = _u8(B, H, W, 7)
    want = capf_oracle.prefetch_images(img, backbone)
    if mirror:
        want = torch.flip(want, [2])
    got = frontend.preprocess(img.cuda(), backbone, mirror=mirror)
    assert got.dtype == torch.float32 and torch.equal(got.cpu(), want)


@pytest.mark.gpu
def test_flip_test_forward_matches_two_pass_reference_recipe():
    """frontend.flip_test_forward (one forward of 2B frames from uint8 crops) == the reference recipe: prefetcher inputs,
    two separate forwards on cloned crop tensors, un-mirror + mean (train.py:170-181) -- with the oracle's CPU forward as
    the model of the second path."""
    B, H, W = 2, 128, 96
    cfg = capf_b200.make_config("hrnet_32")
    model = capf_b200.CA_PF(cfg, precision="fp32").eval()
    w = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 3)
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    img = _u8(B, H, W, 11)
    g = torch.Generator().manual_seed(5)
    kp = torch.rand(B, 17, 2, generator=g) * 2 - 1
    crop = torch.rand(B, 17, 2, generator=g) * torch.tensor([191.0, 255.0])
    crop_dev = crop.cuda()
    with torch.no_grad():
        got = frontend.flip_test_forward(model, img.cuda(), kp.cuda(), crop_dev)
    assert torch.equal(crop_dev.cpu(), crop), "flip-test path must not mutate the caller's crop tensor (train.py clones it)"
    images, kps, crops = capf_oracle.prefetch_flip_test(img, kp, crop, "hrnet_32")
    preds = []
    for v in (0, 1):
        preds.append(capf_oracle.ca_pf_forward(w, "hrnet_32", cfg.model.backbone, images[:, v].contiguous(), kps[:, v].contiguous(),
                                               crops[:, v].clone()))
    want = capf_oracle.flip_test_merge(preds[0], preds[1])
    rel = float((got.cpu() - want).norm() / want.norm())
    assert got.shape == (B, 1, 17, 3) and rel < 1e-4, rel


@pytest.mark.gpu
def test_flip_test_forward_from_frames_equals_crop_then_flip_test():
    """Camera frames -> prediction in one call == crop_image on every frame (bit-exact with cv2, tests/test_crop.py)
    followed by the flip-test forward from uint8 crops."""
    import numpy as np
    from capf_b200.mvn.utils import img as host
    B = 3
    cfg = capf_b200.make_config("hrnet_32")
    model = capf_b200.CA_PF(cfg, precision="fp16").eval()
    w = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 4)
    model.load_state_dict(w, strict=True)
    model = model.cuda()
    g = torch.Generator().manual_seed(8)
    frames = torch.randint(0, 256, (B, 150, 170, 3), generator=g, dtype=torch.uint8).cuda()
    sizes = torch.tensor([[150, 170], [148, 170], [150, 160]], dtype=torch.int32).cuda()
    trans = np.stack([host.get_affine_transform(np.float32([80 + 5 * b, 70 - 3 * b]), np.float32([0.45, 0.6]), 0, (96, 128)) for b in range(B)])
    kp = (torch.rand(B, 17, 2, generator=g) * 2 - 1).cuda()
    crop = (torch.rand(B, 17, 2, generator=g) * torch.tensor([191.0, 255.0])).cuda()
    with torch.no_grad():
        got = frontend.flip_test_forward_from_frames(model, frames, trans, kp, crop, sizes=sizes, image_shape=(96, 128)).clone()
        crops_u8 = host.crop_images(frames, trans, (96, 128), sizes=sizes)
        want = frontend.flip_test_forward(model, crops_u8, kp, crop)
    assert got.shape == (B, 1, 17, 3) and torch.isfinite(got).all() and torch.equal(got, want)


def test_oracle_prefetch_is_pinned_to_the_reference_prefetcher():
    """tests/golden/prefetch_cases.npz = the reference's own data_prefetcher (utils.py:15-88) run on CPU by
    oracle/gen_golden_prefetch.py (its four CUDA entry points stubbed, arithmetic untouched): the restatement and the host
    mirror reproduce it exactly -- images, flip-test stacks of both keypoint tensors, root-relative targets."""
    import os
    from gen_golden_prefetch import make_batch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prefetch_cases.npz"))
    for k in range(int(g["n"])):
        seed, is_cpn, flip = (int(v) for v in g[f"p{k}_cfg"])
        backbone = "cpn" if is_cpn else "hrnet_32"
        images, gt, kp, crop = make_batch(seed)
        want = [torch.from_numpy(g[f"p{k}_{n}"]) for n in ("images", "gt", "kp", "crop")]
        if flip:
            got_img, got_kp, got_crop = capf_oracle.prefetch_flip_test(images, kp, crop, backbone)
            kf, cf = frontend.flip_keypoints(kp, crop)
            assert torch.equal(torch.stack([kp, kf], 1), want[2]) and torch.equal(torch.stack([crop, cf], 1), want[3])
        else:
            got_img, got_kp, got_crop = capf_oracle.prefetch_images(images, backbone), kp, crop
        assert torch.equal(got_img, want[0]) and torch.equal(got_kp, want[2]) and torch.equal(got_crop, want[3])
        assert torch.equal(capf_oracle.prefetch_targets(gt), want[1]) and torch.equal(frontend.root_relative(gt), want[1])
