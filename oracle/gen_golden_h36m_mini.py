"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/h36m_mini/: a tiny dataset in the reference's on-disk formats (pickled
label list + JPEG frames under root/s_XX_act_XX_subact_XX_ca_XX/) and what the UNMODIFIED reference's
Human36MSingleViewDataset (mvn/datasets/human36m.py:482-584) returns for it, whole and sliced for rank 1 of 2.
Run in the authoring container:  python oracle/gen_golden_h36m_mini.py"""
import importlib
import os
import pickle
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "..", "tests", "golden", "h36m_mini")


def make_labels(rng):
    shots = []
    spec = [(9, 2, 1, 0, 11), (9, 2, 2, 1, 12), (9, 16, 1, 3, 7), (11, 5, 2, 2, 31), (11, 16, 2, 0, 5), (11, 3, 1, 1, 64), (9, 3, 1, 2, 9)]
    for vid, (subject, action, subaction, cam, image_id) in enumerate(spec):
        h, w = (150, 170) if cam % 2 == 0 else (152, 170)                 # two frame sizes, like the 1000/1002-pixel cameras
        box_h = float(rng.uniform(70, 130))
        shots.append({
            "image": None, "joints_3d": rng.normal(0, 0.4, (17, 3)).astype(np.float32),
            "joints_2d_cpn": rng.uniform(-1, 1, (17, 2)).astype(np.float32),
            "joints_2d_cpn_crop": (rng.uniform(0, 1, (17, 2)) * [191, 255]).astype(np.float32),
            "center": np.array([rng.uniform(50, 120), rng.uniform(40, 110)], dtype=np.float32),
            "scale": np.array([box_h * 0.75 / 200, box_h / 200], dtype=np.float32),
            "subject": subject, "action": action, "subaction": subaction, "camera_id": cam, "image_id": image_id, "video_id": vid,
            "_hw": (h, w)})
    return shots


def main():
    import ref_import
    ref_import._install_shims()
    if ref_import.REF_PKG not in sys.path:
        sys.path.insert(0, ref_import.REF_PKG)
    ref = importlib.import_module("mvn.datasets.human36m")
    rng = np.random.default_rng(77)
    root = os.path.join(OUT, "processed")
    shots = make_labels(rng)
    for s in shots:
        h, w = s.pop("_hw")
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([127 + 100 * np.sin(xx / 9.0 + s["video_id"]), 127 + 100 * np.cos(yy / 7.0), (xx + 2 * yy) % 256], -1)
        img = np.clip(img + rng.normal(0, 10, img.shape), 0, 255).astype(np.uint8)
        sub = "s_{:02d}_act_{:02d}_subact_{:02d}_ca_{:02d}".format(s["subject"], s["action"], s["subaction"], s["camera_id"] + 1)
        os.makedirs(os.path.join(root, sub), exist_ok=True)
        cv2.imwrite(os.path.join(root, sub, "{}_{:06d}.jpg".format(sub, s["image_id"])), img, [cv2.IMWRITE_JPEG_QUALITY, 90])
    labels_path = os.path.join(OUT, "labels.pkl")
    with open(labels_path, "wb") as f:
        pickle.dump(shots, f, protocol=4)
    out = {}
    for tag, (rank, world) in {"all": (None, None), "r1of2": (1, 2)}.items():
        ds = ref.Human36MSingleViewDataset(root=root, labels_path=labels_path, image_shape=(48, 64), test=True, rank=rank, world_size=world)
        out[f"{tag}_len"] = np.array(len(ds))
        out[f"{tag}_action_idx"] = np.asarray(ds.labels_action_idx)
        out[f"{tag}_video_idx"] = np.asarray(ds.video_idx)
        out[f"{tag}_dist_size"] = np.array(ds.dist_size if ds.dist_size is not None else [])
        items = [ds[i] for i in range(len(ds))]
        out[f"{tag}_image"] = np.stack([it[0] for it in items])
        out[f"{tag}_gt"] = np.stack([it[1] for it in items])
        out[f"{tag}_kp"] = np.stack([it[2] for it in items])
        out[f"{tag}_kp_crop"] = np.stack([it[3] for it in items])
    np.savez_compressed(os.path.join(OUT, "reference_items.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
