"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/prefetch_cases.npz: outputs of the UNMODIFIED reference
data_prefetcher (mvn/datasets/utils.py:15-88) on seeded uint8 batches, for both normalisations, with and without the
flip test.  The class is written for CUDA (`.cuda()`, torch.cuda.Stream); it runs here on the CPU with those four
entry points replaced by no-ops for the duration of the call -- the arithmetic (torch ops on fp32) is untouched.
Run in the authoring container:  python oracle/gen_golden_prefetch.py"""
import contextlib
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def make_batch(seed, b=3, h=32, w=24):
    g = torch.Generator().manual_seed(seed)
    images = torch.randint(0, 256, (b, h, w, 3), generator=g, dtype=torch.uint8)
    gt = torch.randn(b, 1, 17, 3, generator=g)
    kp = torch.rand(b, 17, 2, generator=g) * 2 - 1
    crop = torch.rand(b, 17, 2, generator=g) * torch.tensor([191.0, 255.0])
    return images, gt, kp, crop


@contextlib.contextmanager
def cuda_free_torch():
    class _Stream:
        def wait_stream(self, other):
            pass
    saved = (torch.cuda.Stream, torch.cuda.stream, torch.cuda.current_stream, torch.Tensor.cuda)
    torch.cuda.Stream = _Stream
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.cuda.Stream, torch.cuda.stream, torch.cuda.current_stream, torch.Tensor.cuda = saved


def run_reference(batch, backbone, flip_test):
    import ref_import
    ref_import._install_shims()
    if ref_import.REF_PKG not in sys.path:
        sys.path.insert(0, ref_import.REF_PKG)
    ref = importlib.import_module("mvn.datasets.utils")
    with cuda_free_torch():
        pf = ref.data_prefetcher([[t.clone() for t in batch]], "cpu", False, flip_test, backbone)
        return pf.next()


def main():
    out = {}
    k = 0
    for backbone in ("hrnet_32", "cpn"):
        for flip_test in (False, True):
            batch = make_batch(40 + k)
            images, gt, kp, crop = run_reference(batch, backbone, flip_test)
            out[f"p{k}_cfg"] = np.array([40 + k, int(backbone == "cpn"), int(flip_test)])
            out[f"p{k}_images"], out[f"p{k}_gt"], out[f"p{k}_kp"], out[f"p{k}_crop"] = images.numpy(), gt.numpy(), kp.numpy(), crop.numpy()
            k += 1
    out["n"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "prefetch_cases.npz"), **out)
    print({a: b.shape for a, b in out.items()})


if __name__ == "__main__":
    main()
