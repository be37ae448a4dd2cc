"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/eval_cases.npz: the UNMODIFIED reference's
Human36MSingleViewDataset.evaluate_using_pred (mvn/datasets/human36m.py:358-422) on seeded poses.
Run in the authoring container:  python oracle/gen_golden_eval.py"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def make_case(seed, n_frames):
    """Root-relative poses in metres, predictions with per-frame noise levels, some mirrored (the reflection branch of the
    Procrustes fit), frames of the 30 action trials in shuffled runs (as the per-rank label slices deliver them)."""
    rng = np.random.default_rng(seed)
    gt = rng.normal(0, 0.35, (n_frames, 1, 17, 3)).astype(np.float32)
    gt[:, :, 0] = 0
    walk = np.cumsum(rng.normal(0, 0.01, (n_frames, 1, 17, 3)), axis=0).astype(np.float32)
    gt = gt * 0.2 + walk + rng.normal(0, 0.3, (1, 1, 17, 3)).astype(np.float32)
    pred = gt + (rng.normal(0, 1, gt.shape) * rng.uniform(0.005, 0.12, (n_frames, 1, 1, 1))).astype(np.float32)
    pred[::41, ..., 0] *= -1
    runs = rng.permutation(np.repeat(np.arange(30), 3))
    bounds = np.sort(rng.choice(np.arange(1, n_frames), len(runs) - 1, replace=False))
    labels = np.zeros(n_frames, dtype=np.int64)
    for r, (a, b) in zip(runs, zip(np.r_[0, bounds], np.r_[bounds, n_frames])):
        labels[a:b] = r
    return gt, pred.astype(np.float32), labels


def main():
    import ref_import
    ref_import._install_shims()
    if ref_import.REF_PKG not in sys.path:
        sys.path.insert(0, ref_import.REF_PKG)
    import importlib
    ref = importlib.import_module("mvn.datasets.human36m")
    out = {}
    for k, (seed, n) in enumerate([(1, 1500), (2, 4000)]):
        gt, pred, labels = make_case(seed, n)
        fake = types.SimpleNamespace(labels_action_idx=labels)
        res = ref.Human36MSingleViewDataset.evaluate_using_pred(fake, torch.from_numpy(gt), torch.from_numpy(pred))
        names = sorted(res)
        out[f"e{k}_seed_n"] = np.array([seed, n])
        out[f"e{k}_names"] = np.array(names)
        out[f"e{k}_scores"] = np.array([[res[a]["MPJPE"], res[a]["P_MPJPE"], res[a]["MPJVE"]] for a in names], dtype=np.float64)
    out["n"] = np.array(2)
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "eval_cases.npz"), **out)
    print("wrote eval_cases.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
