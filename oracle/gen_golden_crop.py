"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/crop_cases.npz from the UNMODIFIED reference crop path
(/root/reference/ContextPose/mvn/utils/img.py: get_affine_transform + crop_image, i.e. cv2.getAffineTransform +
cv2.warpAffine) on seeded synthetic frames.  Run in the authoring container:  python oracle/gen_golden_crop.py"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_IMG = "/root/reference/ContextPose/mvn/utils/img.py"


def frames(k, h, w, kind):
    """Synthetic frame of case k (own seed, so tests re-derive the two 1000-pixel frames instead of storing them)."""
    rng = np.random.default_rng(1000 + k)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([127 + 120 * np.sin(xx / 7.0 + yy / 11.0), 127 + 120 * np.cos(xx / 5.0), (xx * 3 + yy * 2) % 256], -1)
    return np.clip(base + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)


def main():
    spec = importlib.util.spec_from_file_location("ref_img", REF_IMG)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    # (frame h, w, kind, center, scale, output_size (W, H)) -- boxes inside, across every border, larger than the frame,
    # tiny (magnifying) and the Human3.6M geometry (1000/1002-pixel frames, 192x256 crop, scale = box / 200)
    cases = [
        (120, 140, "noise", (70.0, 60.0), (0.45, 0.6), (48, 64)),
        (120, 140, "smooth", (10.5, 100.25), (0.5, 0.5 * 4 / 3), (48, 64)),
        (90, 70, "noise", (35.0, 45.0), (0.9, 1.2), (48, 64)),
        (64, 64, "smooth", (31.7, 30.2), (0.06, 0.08), (24, 32)),
        (150, 100, "noise", (120.0, -20.0), (0.7, 0.93), (36, 48)),
        (1002, 1000, "smooth", (512.3, 488.9), (2.618, 3.4906), (192, 256)),
        (1000, 1000, "noise", (880.0, 300.0), (1.9, 2.5333), (192, 256)),
    ]
    keep_frame = []
    for k, (h, w, kind, center, scale, osize) in enumerate(cases):
        img = frames(k, h, w, kind)
        center, scale = np.array(center, dtype=np.float32), np.array(scale, dtype=np.float32)
        trans = ref.get_affine_transform(center, scale, 0, osize)
        crop = ref.crop_image(img, center, scale, osize)
        out[f"c{k}_center"], out[f"c{k}_scale"], out[f"c{k}_osize"] = center, scale, np.array(osize)
        out[f"c{k}_trans"], out[f"c{k}_crop"] = trans, crop
        if h * w <= 20000:
            out[f"c{k}_frame"] = img
        else:                                  # large frames are regenerated from the seed by the test (frames())
            out[f"c{k}_frame_spec"] = np.array([h, w, 0 if kind == "noise" else 1])
            keep_frame.append(k)
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "crop_cases.npz"), **out)
    print("wrote", len(cases), "cases; large frames re-derived in tests:", keep_frame)


if __name__ == "__main__":
    main()
