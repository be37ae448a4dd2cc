"""f4: the evaluation reducer evaluate_using_pred (mvn/datasets/human36m.py:358-422).
CPU: oracle restatement == fixtures written by the reference's own method.  GPU: CAPF_OP_POSE_ERRORS + host merge."""
import os

import numpy as np
import pytest
import torch

import capf_oracle
from gen_golden_eval import make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_cases.npz")
RTOL = 2e-5          # the reference computes in fp32 (torch.mean / numpy float32 SVD); the kernel in fp64


def cases():
    g = np.load(GOLD)
    for k in range(int(g["n"])):
        seed, n = (int(v) for v in g[f"e{k}_seed_n"])
        yield make_case(seed, n), [str(s) for s in g[f"e{k}_names"]], g[f"e{k}_scores"]


def as_table(res, names):
    return np.array([[res[a]["MPJPE"], res[a]["P_MPJPE"], res[a]["MPJVE"]] for a in names])


def test_oracle_reducer_matches_reference_fixtures():
    for (gt, pred, labels), names, want in cases():
        res = capf_oracle.evaluate_using_pred(torch.from_numpy(gt), torch.from_numpy(pred), labels)
        assert sorted(res) == names
        np.testing.assert_allclose(as_table(res, names), want, rtol=1e-6)


def test_previous_in_action_pairs_are_the_masked_diffs():
    from capf_b200.mvn.datasets import human36m as host
    labels = np.array([3, 3, 1, 3, 1, 1, 0, 3])
    assert host.previous_in_action(labels).tolist() == [-1, 0, -1, 1, 2, 4, -1, 3]
    assert host.retval["action_names"] == capf_oracle.H36M_ACTION_NAMES and len(host.retval["action_names"]) == 30
    with pytest.raises(Exception):
        host.pose_errors(torch.zeros(2, 17, 3), torch.zeros(2, 17, 3))            # CPU tensors: no CPU path


@pytest.mark.gpu
def test_gpu_reducer_matches_reference_fixtures_and_oracle():
    from capf_b200.mvn.datasets import human36m as host
    for (gt, pred, labels), names, want in cases():
        res = host.evaluate_using_pred(torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda(), labels)
        assert sorted(res) == names
        np.testing.assert_allclose(as_table(res, names), want, rtol=RTOL)


@pytest.mark.gpu
def test_gpu_pose_errors_per_frame_including_reflections_and_planar_poses():
    """Per-frame rows against the reference formulas evaluated in float64 one frame at a time: mirrored predictions (the
    det(R) = -1 branch), planar ground truth (third singular value 0) and exact predictions (zero error)."""
    from capf_b200.mvn.datasets import human36m as host
    rng = np.random.default_rng(4)
    n = 700
    gt = rng.normal(0, 0.4, (n, 17, 3))
    pred = gt + rng.normal(0, 1, gt.shape) * rng.uniform(0.001, 0.5, (n, 1, 1))
    pred[:60, :, 0] *= -1
    gt[60:90, :, 2] = 0.25
    pred[90:100] = gt[90:100] * 1.7 + 0.3                    # similarity-related: P-MPJPE ~ 0
    gt32, pred32 = gt.astype(np.float32), pred.astype(np.float32)
    prev = np.r_[-1, np.arange(n - 1)].astype(np.int32)
    prev[350] = -1
    rows = host.pose_errors(torch.from_numpy(pred32).cuda(), torch.from_numpy(gt32).cuda(), torch.from_numpy(prev).cuda()).cpu().numpy()
    g64, p64 = gt32.astype(np.float64), pred32.astype(np.float64)
    want = np.zeros((n, 3))
    for k in range(n):
        want[k, 0] = np.linalg.norm(p64[k] - g64[k], axis=-1).mean()
        want[k, 1] = capf_oracle.p_mpjpe(p64[k:k + 1].copy(), g64[k:k + 1].copy())
        if prev[k] >= 0:
            want[k, 2] = np.linalg.norm((pred32[k] - pred32[prev[k]]) - (gt32[k] - gt32[prev[k]]), axis=-1).astype(np.float64).mean()
    np.testing.assert_allclose(rows[:, 0], want[:, 0], rtol=1e-12)
    np.testing.assert_allclose(rows[:, 1], want[:, 1], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(rows[:, 2], want[:, 2], rtol=1e-6, atol=1e-12)


@pytest.mark.gpu
def test_gpu_reducer_scores_nan_for_an_action_with_an_empty_trial():
    from capf_b200.mvn.datasets import human36m as host
    (gt, pred, labels), _, _ = next(cases())
    keep = labels != 5                                        # drop trial 2 of "Eating": 0 * mean(empty) = nan in the reference
    # (with BOTH trials missing the reference divides by a zero frame count and raises)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = capf_oracle.evaluate_using_pred(torch.from_numpy(gt[keep]), torch.from_numpy(pred[keep]), labels[keep])
    res = host.evaluate_using_pred(torch.from_numpy(gt[keep]).cuda(), torch.from_numpy(pred[keep]).cuda(), labels[keep])
    assert all(np.isnan(res["Eating"][m]) and np.isnan(want["Eating"][m]) for m in ("MPJPE", "P_MPJPE", "MPJVE"))
    for a in res:
        if a != "Eating":
            for m in ("MPJPE", "P_MPJPE", "MPJVE"):
                assert abs(res[a][m] - want[a][m]) <= RTOL * abs(want[a][m])
