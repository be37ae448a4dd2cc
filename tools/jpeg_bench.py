"""Frame-decode throughput: capf_jpeg_decode_batch (nvJPEG, a library call) against cv2.imdecode on the host (run on the GPU box).

Synthetic Human3.6M-sized frames (1000x1000, smooth background + noise, quality 90, 4:2:0) are encoded once with OpenCV; the timed
part is streams-in-host-memory -> decoded BGR frames in the padded [B,Hs,Ws,3] storage the crop kernel reads (wall clock around the
call + stream synchronize: the Huffman stage of nvjpegDecode runs on the host, so device events would miss most of it).

  python tools/jpeg_bench.py [frames=64] [reps=3]            every backend, one subprocess each (CAPF_JPEG_BACKEND is read once)
  python tools/jpeg_bench.py [frames] [reps] <backend>        one backend: default | hybrid | gpu_hybrid | hardware
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from capf_b200 import lib  # noqa: E402


def main():
    import cv2
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if len(sys.argv) <= 3:
        import subprocess
        for be in ("default", "gpu_hybrid", "hardware", "hybrid"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(n), str(reps), be], env=dict(os.environ, CAPF_JPEG_BACKEND=be),
                               capture_output=True, text=True, timeout=90)
            tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else "FAILED: " + (r.stderr.strip().splitlines() or ["?"])[-1]
            print(f"[{be}] {tail}", flush=True)
        return
    backend = sys.argv[3]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    cpu_legs = backend == "default"
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:1000, 0:1000].astype(np.float32)
    blobs = []
    for k in range(n):
        base = 110 + 60 * np.sin(xx / (40.0 + k)) * np.cos(yy / (55.0 + 2 * k))
        img = np.stack([base, base * 0.9 + 10, base * 0.8 + 25], axis=-1) + rng.normal(0, 6, (1000, 1000, 3))
        ok, enc = cv2.imencode(".jpg", np.clip(img, 0, 255).astype(np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 90])
        assert ok
        blobs.append(enc.tobytes())
    mb = sum(len(b) for b in blobs) / 1e6
    L = lib.load()
    if not L.capf_jpeg_available():
        print("libnvjpeg not loadable on this machine")
        return
    dev = torch.device("cuda:0")
    frames = torch.zeros(n, 1000, 1000, 3, dtype=torch.uint8, device=dev)
    data = (ctypes.c_char_p * n)(*blobs)
    lens = (ctypes.c_size_t * n)(*[len(b) for b in blobs])
    st = torch.cuda.current_stream(dev)

    def gpu_pass():
        lib.check(L.capf_jpeg_decode_batch(data, lens, n, frames.data_ptr(), 1000, 1000, None, 0, st.cuda_stream), "capf_jpeg_decode_batch")
        st.synchronize()

    gpu_pass()                                   # handle creation, lazy loads
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        gpu_pass()
        t.append(time.perf_counter() - t0)
    g = n / min(t)
    ref = cv2.imdecode(np.frombuffer(blobs[0], np.uint8), cv2.IMREAD_COLOR).astype(np.int32)
    diff = np.abs(frames[0].cpu().numpy().astype(np.int32) - ref)
    d16 = float((diff <= 16).mean())
    if not cpu_legs:
        msg = (f"{n} frames of 1000x1000: capf_jpeg_decode_batch {g:.0f} frames/s (nvjpegDecodeBatched, backend {backend}); "
               f"frame 0 vs cv2: mean |d| {diff.mean():.2f}, max {int(diff.max())}, within 16: {d16:.4f}")
        print(msg)
        with open(os.path.join(ROOT, "gpurun_out", "jpeg_bench.txt"), "a") as f:
            f.write(msg + "\n")
        return
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for b in blobs:
            cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR)
        t.append(time.perf_counter() - t0)
    c = n / min(t)
    from concurrent.futures import ThreadPoolExecutor
    workers = 14                                  # the reference's num_workers (experiments/human36m/train/*.yaml)
    with ThreadPoolExecutor(workers) as pool:
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            list(pool.map(lambda b: cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR), blobs))
            t.append(time.perf_counter() - t0)
    cw = n / min(t)
    msg = (f"{n} frames of 1000x1000 ({mb / n * 1e3:.0f} KB per stream): capf_jpeg_decode_batch {g:.0f} frames/s "
           f"(one nvJPEG state, one host thread); cv2.imdecode {c:.0f} frames/s on one host thread, {cw:.0f} frames/s on {workers} threads "
           f"({os.cpu_count()} host cores); frame 0 vs cv2: mean |d| {diff.mean():.2f}, max {int(diff.max())}")
    print(msg)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "jpeg_bench.txt"), "a") as f:
        f.write(msg + "\n")


if __name__ == "__main__":
    main()
