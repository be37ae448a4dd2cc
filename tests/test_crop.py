"""f4: crop_image = cv2.warpAffine on uint8 frames (mvn/utils/img.py:51-69).
CPU: the oracle restatement against the fixtures written by the reference's own crop_image, and against cv2 itself.
GPU: CAPF_OP_WARP_AFFINE_U8 against the oracle, bit for bit."""
import os

import numpy as np
import pytest
import torch

import capf_oracle
from gen_golden_crop import frames

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "crop_cases.npz")


def load_cases():
    g = np.load(GOLD)
    out = []
    for k in range(int(g["n"])):
        if f"c{k}_frame" in g:
            img = g[f"c{k}_frame"]
        else:
            h, w, kind = (int(v) for v in g[f"c{k}_frame_spec"])
            img = frames(k, h, w, "noise" if kind == 0 else "smooth")
        out.append(dict(k=k, frame=img, center=g[f"c{k}_center"], scale=g[f"c{k}_scale"], osize=tuple(int(v) for v in g[f"c{k}_osize"]),
                        trans=g[f"c{k}_trans"], crop=g[f"c{k}_crop"]))
    return out


def test_oracle_crop_matches_reference_fixtures():
    for c in load_cases():
        # the warp given the reference's own matrix: every byte
        got = capf_oracle.warp_affine_u8(c["frame"], c["trans"], c["osize"])
        assert got.shape == c["crop"].shape and np.array_equal(got, c["crop"]), c["k"]
        # the matrix (a float64 LU solve in both; LAPACK and OpenCV may differ in the last place)
        t = capf_oracle.affine_transform(c["center"], c["scale"], c["osize"])
        assert np.allclose(t, c["trans"], rtol=0, atol=1e-9 * max(1.0, np.abs(c["trans"]).max()))
        full = capf_oracle.crop_image(c["frame"], c["center"], c["scale"], c["osize"])
        assert (full != c["crop"]).mean() < 1e-4                  # a 1e-16 matrix difference may flip a 1/32-pixel rounding


def test_oracle_warp_matches_cv2_on_random_maps():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for t in range(12):
        h, w = int(rng.integers(40, 400)), int(rng.integers(40, 400))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        s = rng.uniform(0.2, 3.0)
        m = np.array([[s, 0, rng.uniform(-80, 40)], [0, s, rng.uniform(-80, 40)]])
        want = cv2.warpAffine(img, m, (96, 128), flags=cv2.INTER_LINEAR)
        assert np.array_equal(capf_oracle.warp_affine_u8(img, m, (96, 128)), want), t


def test_host_mirror_matrices_match_oracle():
    from capf_b200.mvn.utils import img as host
    for c in load_cases():
        assert np.array_equal(host.get_affine_transform(c["center"], c["scale"], 0, c["osize"]),
                              capf_oracle.affine_transform(c["center"], c["scale"], c["osize"]))
        assert np.array_equal(host.invert_affine(c["trans"]), capf_oracle.invert_affine(c["trans"]))
    with pytest.raises(NotImplementedError):
        host.get_affine_transform((1.0, 2.0), (1.0, 1.0), 30, (48, 64))
    with pytest.raises(Exception):
        host.crop_images(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), np.eye(2, 3)[None], (4, 4))       # CPU tensor: no CPU path


@pytest.mark.gpu
def test_gpu_crop_is_bit_exact_with_oracle_and_fixtures():
    from capf_b200.mvn.utils import img as host
    for c in load_cases():
        f = torch.from_numpy(c["frame"]).cuda()
        got = host.crop_images(f[None], c["trans"][None], c["osize"])[0].cpu().numpy()
        assert np.array_equal(got, c["crop"]), c["k"]
        one = host.crop_image(f, c["center"], c["scale"], c["osize"]).cpu().numpy()
        assert np.array_equal(one, capf_oracle.crop_image(c["frame"], c["center"], c["scale"], c["osize"]))


@pytest.mark.gpu
def test_gpu_crop_batch_with_ragged_frames_rotation_and_borders():
    """A padded batch whose frames have different live sizes (Human3.6M: 1000x1000 and 1002x1000), general affine maps
    (rotation / shear: every matrix entry non-zero) and boxes that leave the frame on every side."""
    from capf_b200.mvn.utils import img as host
    rng = np.random.default_rng(9)
    B, Hs, Ws = 9, 131, 157
    frames_np = rng.integers(0, 256, (B, Hs, Ws, 3), dtype=np.uint8)
    sizes = np.stack([rng.integers(60, Hs + 1, B), rng.integers(60, Ws + 1, B)], 1).astype(np.int32)
    sizes[0] = (Hs, Ws)
    trans = []
    for b in range(B):
        a, s = rng.uniform(-0.6, 0.6), rng.uniform(0.3, 2.5)
        trans.append([[s * np.cos(a), -s * np.sin(a) * 1.1, rng.uniform(-60, 30)], [s * np.sin(a), s * np.cos(a) * 0.9, rng.uniform(-60, 30)]])
    trans = np.array(trans)
    got = host.crop_images(torch.from_numpy(frames_np).cuda(), trans, (48, 64), sizes=torch.from_numpy(sizes).cuda()).cpu().numpy()
    for b in range(B):
        want = capf_oracle.warp_affine_u8(frames_np[b, :sizes[b, 0], :sizes[b, 1]], trans[b], (48, 64))
        assert np.array_equal(got[b], want), b


@pytest.mark.gpu
@pytest.mark.parametrize("backbone", ["hrnet_32", "cpn"])
@pytest.mark.parametrize("mirror", [False, True])
def test_gpu_fused_crop_preprocess_equals_two_steps(backbone, mirror):
    from capf_b200 import frontend
    from capf_b200.mvn.utils import img as host
    rng = np.random.default_rng(3)
    f = torch.from_numpy(rng.integers(0, 256, (5, 90, 110, 3), dtype=np.uint8)).cuda()
    trans = np.array([[[0.8 + 0.1 * b, 0.0, -3.0 * b], [0.0, 0.8 + 0.1 * b, 4.0 - b]] for b in range(5)])
    crop = host.crop_images(f, trans, (48, 64))
    want = frontend.preprocess(crop, backbone, mirror=mirror)
    got = host.crop_images(f, trans, (48, 64), normalise=backbone, mirror=mirror)
    assert got.dtype == torch.float32 and torch.equal(got, want)
