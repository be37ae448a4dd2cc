"""Seeded synthetic weights and inputs shared by the golden generator, the parity
tests, ``smoke()`` and ``bench.py``.  Pure data synthesis: no model code, no dependency on the oracle.

The reference ships no checkpoints we can reach (no network) and its default init zeroes the
tensors that drive the deformable sampler (pose_dformer.py:103-113,184), so parity runs use
*seeded random* tensors for every ``state_dict`` entry, generated from (name, shape, seed) only.
That makes weights reproducible on the GPU box (where /root/reference is absent) without shipping
100 MB blobs: the same function feeds the reference model here (-> tests/golden) and our model there.

Pure numpy + torch-CPU; no dependency on the reference, the oracle or the CUDA library.
"""
import zlib

import numpy as np
import torch


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([zlib.crc32(name.encode()), seed]))


def make_weights(spec, seed: int = 0):
    """spec: iterable of (key, shape) in state_dict order -> {key: fp32/int64 torch tensor}.

    Rules (chosen so BN folding, both grid_sample padding modes and every residual path are
    exercised while activations stay O(1..10), i.e. representable in fp16):
      BatchNorm (has sibling running_mean): weight U[0.5,1.0] bias N(0,.1) mean U[-.2,.2] var U[.5,1.5]
      LayerNorm (norm1/norm2/head.0)      : weight U[0.8,1.2] bias N(0,.05)
      conv weight (4-D)                   : N(0, 1/sqrt(fan_in))
      linear weight (2-D)                 : U(+-1/sqrt(fan_in));  sampling_offsets.weight N(0,.05);
                                            attention_weights.weight N(0,.1)
      other 1-D (linear bias)             : N(0,.02); sampling_offsets.bias N(0,.15)
      Spatial_pos_embed                   : N(0,.02)
    """
    spec = [(k, tuple(s)) for k, s in spec]
    keys = {k for k, _ in spec}
    out = {}
    for k, shape in spec:
        g = _rng(k, seed)
        stem, _, leaf = k.rpartition(".")
        is_bn = (stem + ".running_mean") in keys
        n = int(np.prod(shape)) if len(shape) else 1
        if leaf == "num_batches_tracked":
            out[k] = torch.zeros((), dtype=torch.int64)
            continue
        if is_bn:
            if leaf == "weight":
                a = g.uniform(0.5, 1.0, n)
            elif leaf == "bias":
                a = g.normal(0.0, 0.1, n)
            elif leaf == "running_mean":
                a = g.uniform(-0.2, 0.2, n)
            else:
                a = g.uniform(0.5, 1.5, n)
        elif k.endswith("Spatial_pos_embed"):
            a = g.normal(0.0, 0.02, n)
        elif len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            a = g.normal(0.0, 1.0 / np.sqrt(fan_in), n)
        elif len(shape) == 2:
            if "sampling_offsets" in k:
                a = g.normal(0.0, 0.05, n)
            elif "attention_weights" in k:
                a = g.normal(0.0, 0.1, n)
            else:
                b = 1.0 / np.sqrt(shape[1])
                a = g.uniform(-b, b, n)
        else:  # 1-D, not BN
            is_ln = stem.endswith(("norm1", "norm2", "head.0"))
            if is_ln and leaf == "weight":
                a = g.uniform(0.8, 1.2, n)
            elif is_ln:
                a = g.normal(0.0, 0.05, n)
            elif "sampling_offsets" in k:
                a = g.normal(0.0, 0.15, n)
            else:
                a = g.normal(0.0, 0.02, n)
        out[k] = torch.from_numpy(a.astype(np.float32).reshape(shape))
    return out


def make_inputs(batch: int, height: int, width: int, seed: int = 1234, outlier_frac: float = 0.05):
    """Synthetic inputs of SURVEY.md section 8d.

    images  [B,H,W,3] f32 ~ N(0,1)           (mean/std-normalised RGB, datasets/utils.py:45-50)
    kp2d    [B,17,2]  f32 in [-1,1]x[-H/W,H/W] (normalize_screen_coordinates, transform.py:92-96)
    crop    [B,17,2]  f32 in [0,191]x[0,255] crop pixels (conpose.py:34-35 hard-codes 192x256),
            with ``outlier_frac`` of the joints pushed outside the map to hit both padding modes.
    """
    g = np.random.Generator(np.random.PCG64([0xCA9F, seed]))
    images = g.standard_normal((batch, height, width, 3), dtype=np.float32)
    kp2d = (g.random((batch, 17, 2), dtype=np.float32) * 2.0 - 1.0)
    kp2d[..., 1] *= np.float32(height / width)
    crop = g.random((batch, 17, 2), dtype=np.float32) * np.array([191.0, 255.0], np.float32)
    n_out = int(round(outlier_frac * batch * 17))
    if n_out:
        idx = g.choice(batch * 17, size=n_out, replace=False)
        lo = g.random((n_out, 2), dtype=np.float32)
        far = np.where(lo < 0.5, -20.0 * lo * 2, np.array([192.0, 256.0], np.float32) * (1.0 + 0.2 * (lo - 0.5) * 2))
        crop.reshape(-1, 2)[idx] = far.astype(np.float32)
    # a few exact-boundary reference points (ref == -1 / +1 after normalisation)
    if batch * 17 >= 8:
        flat = crop.reshape(-1, 2)
        flat[0] = (0.0, 0.0)
        flat[1] = (192.0, 256.0)
        flat[2] = (96.0, 128.0)
    return (torch.from_numpy(images), torch.from_numpy(kp2d), torch.from_numpy(np.ascontiguousarray(crop)))
