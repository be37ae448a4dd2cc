"""f4: the dataset row -- label table, frame paths, per-rank slicing and the crops, against what the
reference's own Human36MSingleViewDataset returned for tests/golden/h36m_mini (oracle/gen_golden_h36m_mini.py)."""
import os

import numpy as np
import pytest
import torch

import capf_oracle

MINI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h36m_mini")


def open_ds(rank=None, world=None, image_shape=(48, 64)):
    from capf_b200.mvn.datasets.human36m import Human36MSingleViewDataset
    return Human36MSingleViewDataset(os.path.join(MINI, "processed"), os.path.join(MINI, "labels.pkl"), image_shape=image_shape, rank=rank, world_size=world)


@pytest.mark.parametrize("tag,rank,world", [("all", None, None), ("r1of2", 1, 2)])
def test_label_table_paths_and_rank_slices_match_reference(tag, rank, world):
    cv2 = pytest.importorskip("cv2")
    g = np.load(os.path.join(MINI, "reference_items.npz"))
    ds = open_ds(rank, world)
    assert len(ds) == int(g[f"{tag}_len"])
    assert np.array_equal(ds.labels_action_idx, g[f"{tag}_action_idx"]) and np.array_equal(ds.video_idx, g[f"{tag}_video_idx"])
    assert (ds.dist_size or []) == g[f"{tag}_dist_size"].tolist()
    for i in range(len(ds)):
        assert os.path.isfile(ds.image_path(i))
        s = ds.labels[i]
        # decode (library) + the oracle's crop == the crop the reference's __getitem__ returned
        crop = capf_oracle.crop_image(ds.read_frame(i), s["center"], s["scale"], ds.image_shape)
        assert (crop != g[f"{tag}_image"][i]).mean() < 1e-3
        assert np.array_equal(np.expand_dims(s["joints_3d"], 0), g[f"{tag}_gt"][i])
        assert np.array_equal(s["joints_2d_cpn"], g[f"{tag}_kp"][i]) and np.array_equal(s["joints_2d_cpn_crop"], g[f"{tag}_kp_crop"][i])


def test_threaded_frame_reads_keep_order():
    pytest.importorskip("cv2")
    ds = open_ds()
    idx = [3, 0, 6, 6, 1]
    a, b = ds.read_frames(idx, workers=4), [ds.read_frame(i) for i in idx]
    assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    with pytest.raises(FileNotFoundError):
        ds.labels[0]["image_id"] = 999999
        ds.read_frame(0)


@pytest.mark.gpu
def test_gpu_batches_equal_reference_items():
    pytest.importorskip("cv2")
    g = np.load(os.path.join(MINI, "reference_items.npz"))
    for tag, rank, world in (("all", None, None), ("r1of2", 1, 2)):
        ds = open_ds(rank, world)
        b = ds.batch(list(range(len(ds))), "cuda")
        got = b["images"].cpu().numpy()
        assert got.shape == g[f"{tag}_image"].shape and (got != g[f"{tag}_image"]).mean() < 1e-3     # LAPACK vs OpenCV solve, see test_crop
        assert np.array_equal(b["keypoints_3d_gt"].cpu().numpy(), g[f"{tag}_gt"])
        assert np.array_equal(b["keypoints_2d_cpn"].cpu().numpy(), g[f"{tag}_kp"])
        assert np.array_equal(b["keypoints_2d_cpn_crop"].cpu().numpy(), g[f"{tag}_kp_crop"])
        # with the reference's own matrices the crops are the same bytes
        import cv2
        from capf_b200.mvn.utils import img
        for i in range(len(ds)):
            s = ds.labels[i]
            want = cv2.warpAffine(ds.read_frame(i), b["trans"][i], ds.image_shape, flags=cv2.INTER_LINEAR)
            assert np.array_equal(got[i], want), i


@pytest.mark.gpu
def test_gpu_evaluation_epoch_from_frames_runs_end_to_end():
    """labels + JPEG frames -> batch -> crop + normalise + flip test -> CA_PF -> per-action scores, all through the mirrors."""
    pytest.importorskip("cv2")
    import capf_b200
    import protocol
    from capf_b200 import frontend
    ds = open_ds(image_shape=(96, 128))                          # HRNet needs sides divisible by 32
    cfg = capf_b200.make_config("hrnet_32")
    model = capf_b200.CA_PF(cfg, precision="fp16").eval()
    model.load_state_dict(protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 2), strict=True)
    model = model.cuda()
    b = ds.batch([i % len(ds) for i in range(60)], "cuda")
    ds.labels_action_idx = np.repeat(np.arange(30), 2)          # two frames per action trial (the mini table has too few to score)
    with torch.no_grad():
        pred = frontend.flip_test_forward_from_frames(model, b["frames"], b["trans"], b["keypoints_2d_cpn"], b["keypoints_2d_cpn_crop"],
                                                      sizes=b["sizes"], image_shape=ds.image_shape).clone()
    scores = ds.evaluate_using_pred(b["keypoints_3d_gt"], pred)
    want = capf_oracle.evaluate_using_pred(b["keypoints_3d_gt"].cpu(), pred.cpu(), ds.labels_action_idx)
    assert set(scores) == set(want) and len(scores) == 15
    for a in scores:
        for m in ("MPJPE", "P_MPJPE", "MPJVE"):
            x, y = scores[a][m], want[a][m]
            assert (np.isnan(x) and np.isnan(y)) or abs(x - y) <= 2e-5 * abs(y), (a, m, x, y)


@pytest.mark.gpu
def test_gpu_nvjpeg_decode_agrees_with_cv2_and_feeds_the_crop():
    """decode='nvjpeg': frames decoded on the GPU into the padded storage.  nvJPEG and libjpeg-turbo (cv2) are different
    decoders of the same 4:2:0 streams (chroma upsampling, IDCT rounding): on these noisy synthetic frames they differ by
    3.4 grey levels on average, 99.0 % of the bytes within 16 (profiles/r1c_nvjpeg_vs_cv2.txt) -- a tolerance test with
    margin; sizes are exact and the padding stays untouched."""
    pytest.importorskip("cv2")
    from capf_b200 import lib
    if not lib.load().capf_jpeg_available():
        pytest.skip("libnvjpeg not loadable on this machine")
    ds = open_ds()
    idx = list(range(len(ds)))
    frames, sizes = ds.decode_frames(idx, "cuda")
    frames = frames.cpu().numpy().astype(np.int32)
    sizes = sizes.cpu().numpy()
    for k in idx:
        want = ds.read_frame(k).astype(np.int32)
        assert tuple(int(v) for v in sizes[k]) == want.shape[:2]
        d = np.abs(frames[k, :want.shape[0], :want.shape[1]] - want)
        assert d.mean() < 5.0 and (d <= 16).mean() > 0.98, (k, d.mean(), d.max())
        assert not frames[k, want.shape[0]:].any() and not frames[k, :, want.shape[1]:].any()      # padding untouched
    a, b = ds.batch(idx, "cuda", decode="nvjpeg"), ds.batch(idx, "cuda", decode="cv2")
    dc = (a["images"].int() - b["images"].int()).abs().float()
    assert a["images"].shape == b["images"].shape and float(dc.mean()) < 3.5       # measured 2.1 after the 3x down-scaling crop
    assert torch.equal(a["keypoints_2d_cpn"], b["keypoints_2d_cpn"])


_BATCHED_DECODE = r"""
import os, sys, numpy as np
sys.path.insert(0, {root!r})
from capf_b200.mvn.datasets.human36m import Human36MSingleViewDataset
ds = Human36MSingleViewDataset(os.path.join({mini!r}, "processed"), os.path.join({mini!r}, "labels.pkl"), image_shape=(48, 64))
idx = list(range(len(ds)))
frames, sizes = ds.decode_frames(idx, "cuda")
np.save({out!r}, frames.cpu().numpy())
"""


@pytest.mark.gpu
def test_gpu_nvjpeg_batched_backend_decodes_the_same_frames(tmp_path):
    """CAPF_JPEG_BACKEND=gpu_hybrid (nvjpegDecodeBatched, Huffman stage on the SMs; 3.5x the per-frame default on 1000x1000 frames,
    profiles/r3f_jpeg_throughput.txt) is read once per process, so it runs in a child: same sizes, padding untouched, pixels within
    the decoder tolerance of the default backend's; an unknown backend name is an error, not a fallback."""
    pytest.importorskip("cv2")
    import subprocess
    import sys
    from capf_b200 import lib
    if not lib.load().capf_jpeg_available():
        pytest.skip("libnvjpeg not loadable on this machine")
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path / "frames.npy")
    code = _BATCHED_DECODE.format(root=os.path.dirname(here), mini=MINI, out=out)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CAPF_JPEG_BACKEND="gpu_hybrid"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(out).astype(np.int32)
    ds = open_ds()
    idx = list(range(len(ds)))
    base, _ = ds.decode_frames(idx, "cuda")
    base = base.cpu().numpy().astype(np.int32)
    assert got.shape == base.shape
    for k in idx:
        want = ds.read_frame(k).astype(np.int32)
        d = np.abs(got[k, :want.shape[0], :want.shape[1]] - want)
        assert d.mean() < 5.0 and (d <= 16).mean() > 0.98, (k, d.mean(), d.max())
        assert not got[k, want.shape[0]:].any() and not got[k, :, want.shape[1]:].any()
    assert np.abs(got - base).mean() < 1.0                       # the two nvJPEG backends against each other
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CAPF_JPEG_BACKEND="no_such_backend"), capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "CAPF_JPEG_BACKEND" in r.stderr
