#!/usr/bin/env python
"""bench.py -- frames/sec of the CA_PF lifting path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps 20 --warmup 5                 # this arm (libcapf_b200), BASELINE configs[1]
  python bench.py --config 2                                      # BASELINE configs[2]: HRNet-48, bs=512, 384x288, bf16
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W                  # N GPUs, one rank per GPU, NCCL
  python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference's CPU forward (oracle port)

A step is one CA_PF.forward over one batch of synthetic input.  `--config k` selects BASELINE.json configs[k]
(default 1: HRNet-32 + 4-level PoseFormer, 17 joints, bs=256, 256x256, fp16 storage / fp32 accumulate).  With N GPUs every
rank runs its own shard of that many frames (frames are independent, weak scaling) and the step ends with the NCCL
all-gather of the [B,1,17,3] outputs -- the only collective of the path (reference train.py:216-226) -- captured inside
the step's CUDA graph.

One JSON line on stdout (rank 0).
  value         device-resident inputs, CUDA-graph replay, CUDA events, max over ranks.
  e2e           the same metric through the public API from pinned HOST memory: per step the uint8 crops the reference's
                loader delivers + keypoints -> H2D -> `frontend.preprocess` (the image half of data_prefetcher.preload,
                mvn/datasets/utils.py:45-50) -> CA_PF.forward -> D2H of the result, all inside the timed region.
  roofline      the dominant kernel (largest summed device time; per-op CUDA events on the launching stream, one in-order
                pass): algorithmic bytes or FLOPs per launch / average launch duration against MEASURED_PEAKS.json.
                roofline.qkv_gemm: the joint-block QKV GEMM (the path north_star quotes), in-step figure first.
  cpu_baseline  the oracle (CPU restatement of the reference, kind "port") on this box's host cores, bounded sample.
  gpu_yardstick the reference forward in PyTorch-eager CUDA on the same GPU (the oracle's functions on cuda:0 = the same
                ATen/cuDNN/cuBLAS calls conpose.py:30-42 makes): fp32 with TF32 off, fp32 with PyTorch's default TF32 convs,
                and torch.autocast in the config's 16-bit dtype (channels-last) -- frames/s and rel-L2 against the fp32 CPU
                oracle ("PyTorch-autocast's own deviation alongside").
  other_configs the other single-node BASELINE configs timed in the same run (configs[2] at N=1; configs[3], [4] at N=8).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

_T0 = time.time()


def log(msg):
    """Phase timestamps on stderr (rank-tagged): what a slow run spent its time on."""
    print(f"[bench +{time.time() - _T0:6.1f}s rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}

# BASELINE.json configs[k] -> per-GPU workload (configs[3], [4] are 8-GPU totals: 2048/8 and 4096/8 frames per GPU)
CONFIGS = {
    1: dict(backbone="hrnet_32", batch=256, height=256, width=256, precision="fp16", gpus=1,
            text="HRNet-32 backbone, 17 joints, bs=256, 256x256, fp16, 1xB200"),
    2: dict(backbone="hrnet_48", batch=512, height=384, width=288, precision="bf16", gpus=1,
            text="HRNet-48 backbone, 17 joints, bs=512, 384x288, bf16, 1xB200"),
    3: dict(backbone="cpn", batch=256, height=256, width=256, precision="bf16", gpus=8,
            text="CPN backbone, 17 joints, bs=2048, 256x256, bf16, 8xB200 frame-sharded"),
    4: dict(backbone="hrnet_48", batch=512, height=384, width=288, precision="bf16", gpus=8,
            text="HRNet-48, 17 joints, bs=4096, 384x288, bf16, 8xB200 with NCCL output all-gather"),
}
DT_NAME = {"fp16": "f16", "bf16": "bf16", "fp32": "f32", "bf16x3": "bf16x3"}
DT_TEXT = {"fp16": "f16 storage, f32 accumulate", "bf16": "bf16 storage, f32 accumulate", "fp32": "f32 (CUDA-core kernels)",
           "bf16x3": "f32 storage, tensor-core products on split bf16 operands (hi*Wh + lo*Wh + hi*Wl), f32 accumulate"}


def resolve_config(args):
    """Fill backbone/batch/size/precision from --config unless given explicitly; returns the workload label."""
    base = CONFIGS[args.config]
    custom = []
    for k in ("backbone", "batch", "height", "width", "precision"):
        if getattr(args, k) is None:
            setattr(args, k, base[k])
        elif getattr(args, k) != base[k]:
            custom.append(k)
    return workload_label(args, args.config if not custom else None)


def workload_label(a, cfg_idx):
    s = f"{a.backbone} + 4-level PoseFormer, 17 joints, bs={a.batch} per GPU, {a.height}x{a.width}, {a.precision}"
    if cfg_idx is None:
        return s + " (custom: not a BASELINE config)"
    return s + f" (BASELINE configs[{cfg_idx}]: \"{CONFIGS[cfg_idx]['text']}\")"


def metric_name(a):
    return f"frames/sec (17-joint, {a.height}x{a.width}, bs={a.batch})"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi needs a
    few hundred ms to produce its first line, longer than a short timed region, so the sampler is started early (before the
    warm-up) and `stop(t_begin, t_end)` keeps the samples whose arrival time falls inside the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, windows):
        """windows: list of (t_begin, t_end) wall-clock spans of the timed regions."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = self.rows
        window = "timed regions (device-resident + end-to-end loops)"
        inside = [r for r in rows if any(b <= r[0] <= e + 0.06 for b, e in windows)]
        if inside:
            rows = inside
        elif windows:   # regions shorter than the sampling period: the samples closest to the first one
            mid = 0.5 * (windows[0][0] + windows[0][1])
            rows = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            window = "nearest samples (timed region shorter than the sampling period)"
        sm, mx, reasons = [], None, set()
        for _, r in rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx = float(c[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def build_model(backbone, precision, device, graph):
    import capf_b200
    cfg = capf_b200.make_config(backbone)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = capf_b200.CA_PF(cfg, precision=precision, use_cuda_graph=graph).eval()
    w = capf_b200.synth.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 0)
    model.load_state_dict(w, strict=True)
    return model.to(device), w, cfg


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU forward (oracle port) on this box's host cores
# ------------------------------------------------------------------------------------------------------
def time_cpu_reference(backbone, H, W, sample_frames, steps, warmup, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import capf_b200
    import capf_oracle
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    cfg = capf_b200.make_config(backbone)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = capf_b200.CA_PF(cfg)
    w = capf_b200.synth.make_weights([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 0)
    images, kp2d, crop = capf_b200.synth.make_inputs(sample_frames, H, W, 1234)
    times = []
    for it in range(warmup + steps):
        c = crop.clone()
        t0 = time.perf_counter()
        capf_oracle.ca_pf_forward(w, backbone, cfg.model.backbone, images, kp2d, c)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, threads


def run_reference(args, label):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_sample
    times, threads = time_cpu_reference(args.backbone, args.height, args.width, sample, args.steps, args.warmup)
    total = sum(times)
    fps = sample * len(times) / total
    line = {
        "impl": "reference", "metric": metric_name(args), "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded randn images, random-init weights of the named architecture)",
        "config": {"workload": label},
        "sample": f"each step = one forward over {sample} frames of the workload (bounded CPU sample of the {args.batch}-frame batch)",
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"oracle/capf_oracle.py (CPU restatement of the reference forward, torch-CPU fp32), "
                                   f"{sample}-frame batches of the same workload, {len(times)} timed steps, all host threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU yardstick: the reference forward in PyTorch-eager CUDA (library kernels: cuDNN / cuBLAS / ATen)
# ------------------------------------------------------------------------------------------------------
def gpu_yardstick(backbone, bb_cfg, weights, images, kp2d, crop, precision, dev, parity_frames, want_cpu, steps=3):
    """Times oracle.ca_pf_forward -- the reference's own sequence of torch calls -- on `dev`.  Modes: fp32 with TF32 off
    (the reference's numerics), fp32 with PyTorch's default TF32 convolutions, autocast(16-bit) with channels-last tensors.
    Returns frames/s and the rel-L2 of each mode against the fp32 CPU oracle on the first `parity_frames` frames."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import capf_oracle
    B = images.shape[0]
    sd = {k: v.to(dev) for k, v in weights.items()}
    sd_cl = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
    img_d, kp_d, crop_d = images.to(dev), kp2d.to(dev), crop.to(dev)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    adt = {"fp16": torch.float16, "bf16": torch.bfloat16}.get(precision)

    def fwd(mode, n=None):
        im, kp, cr = (img_d, kp_d, crop_d.clone()) if n is None else (img_d[:n], kp_d[:n], crop_d[:n].clone())
        if mode == "autocast":
            with torch.autocast("cuda", dtype=adt):
                x = im.permute(0, 3, 1, 2)                       # NHWC storage viewed as NCHW == channels_last, no copy
                ref = capf_oracle.normalize_crop_(cr)
                feats = capf_oracle.cpn_forward(sd_cl, x) if backbone == "cpn" else capf_oracle.hrnet_forward(sd_cl, x, bb_cfg)
                return capf_oracle.lifter_forward(sd_cl, kp, ref, feats).float()
        return capf_oracle.ca_pf_forward(sd, backbone, bb_cfg, im, kp, cr)

    out = {"what": "oracle.ca_pf_forward (the reference's torch calls, conpose.py:30-42) in PyTorch-eager CUDA on this GPU; "
                   "cudnn.benchmark on; CUDA events; library kernels only (none of this repo's)", "batch": B, "modes": {}}
    modes = [("fp32_tf32_off", False), ("fp32_tf32_convs_default", True)] + ([("autocast", True)] if adt is not None else [])
    try:
        with torch.no_grad():
            for name, tf32 in modes:
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = False
                mode = "autocast" if name == "autocast" else "fp32"
                for _ in range(2):
                    fwd(mode)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    fwd(mode)
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / steps
                got = fwd(mode, parity_frames).cpu()
                key = name if name != "autocast" else f"autocast_{DT_NAME[precision]}_channels_last"
                out["modes"][key] = {"value": B / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                                     "rel_l2_vs_fp32_cpu_oracle": rel_l2(got, want_cpu),
                                     "mpjpe_vs_ref_mm": float((got - want_cpu).norm(dim=-1).mean()) * 1000.0}
    except Exception as e:  # noqa: BLE001  -- a yardstick failure (e.g. out of memory) must not lose the bench line
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
        del sd, sd_cl
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------
# this arm
# ------------------------------------------------------------------------------------------------------
class Workload:
    """One (backbone, batch, size, precision) on this rank: model, plan, inputs, optional in-graph output gather."""

    def __init__(self, a, dev, rank, world, graph=True):
        import capf_b200
        self.a, self.dev, self.rank, self.world = a, dev, rank, world
        B, H, W = a.batch, a.height, a.width
        self.model, self.weights, self.cfg = build_model(a.backbone, a.precision, dev, graph=graph)
        self.images, self.kp2d, self.crop = capf_b200.synth.make_inputs(B, H, W, 1234 + rank)
        self.static = self.model.static_inputs(B, H, W, dev)
        self.plan = self.model.plan_for(B, H, W, dev)
        self.gather = None
        if world > 1:
            self.gather = capf_b200.dist.OutputGatherer([B] * world, (1, 17, 3), dev)
            if graph and os.environ.get("CAPF_GRAPH_GATHER", "1") != "0":
                self.gather.attach(self.model, B, H, W)          # the all-gather becomes the last node of the step's graph
        self.static["images"].copy_(self.images.to(dev))
        self.kp_d, self.crop_pristine = self.kp2d.to(dev), self.crop.to(dev)
        self.crop_work = self.crop_pristine.clone()

    def step_resident(self):
        self.crop_work.copy_(self.crop_pristine)                 # forward normalises it in place (conpose.py:34-35)
        out = self.model(self.static["images"], self.kp_d, self.crop_work)
        if self.gather is None:
            return out
        return self.gather.result() if self.gather.attached else self.gather(out)

    def local(self, full):
        B = self.a.batch
        return full if self.gather is None else full[self.rank * B:(self.rank + 1) * B]


def timed_resident(wl, steps, warmup, barrier):
    with torch.no_grad():
        for _ in range(max(warmup, 3)):
            out = wl.step_resident()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        wall0 = time.time()
        e0.record()
        for _ in range(steps):
            out = wl.step_resident()
        e1.record()
        barrier()
        wall1 = time.time()
    return e0.elapsed_time(e1), out, (wall0, wall1)


def timed_e2e(wl, steps, barrier):
    """Every step: H2D of its inputs from pinned host memory -- the uint8 BGR crops a loader delivers plus the keypoints --
    on a copy stream (double-buffered, one step ahead of the compute), frontend.preprocess (uint8 -> normalised fp32 RGB,
    mvn/datasets/utils.py:45-50) straight into the plan's image buffer, CA_PF.forward through the public API, async D2H of
    the [B,1,17,3] result into pinned memory.  The host never blocks inside the loop: slot reuse is ordered by events."""
    import capf_b200
    a, dev, B = wl.a, wl.dev, wl.a.batch
    g = torch.Generator().manual_seed(99 + wl.rank)
    h_u8 = torch.randint(0, 256, (B, a.height, a.width, 3), dtype=torch.uint8, generator=g).pin_memory()
    h_kp, h_crop = wl.kp2d.pin_memory(), wl.crop.pin_memory()
    h_out = [torch.empty(B, 1, 17, 3).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    stage = [dict(kp=torch.empty_like(wl.kp_d), crop=torch.empty_like(wl.crop_work), u8=torch.empty(h_u8.shape, dtype=torch.uint8, device=dev),
                  ev=torch.cuda.Event(), done=torch.cuda.Event()) for _ in range(2)]
    static_img = wl.static["images"]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            s = stage[slot]
            s["u8"].copy_(h_u8, non_blocking=True)
            s["kp"].copy_(h_kp, non_blocking=True)
            s["crop"].copy_(h_crop, non_blocking=True)
            s["ev"].record(copy_stream)

    def loop(n):
        cur = torch.cuda.current_stream(dev)
        upload(0)
        for i in range(n):
            s = stage[i & 1]
            if i + 1 < n:
                if i >= 1:
                    copy_stream.wait_event(stage[(i + 1) & 1]["done"])    # step i-1 has consumed that slot
                upload((i + 1) & 1)
            cur.wait_event(s["ev"])
            capf_b200.frontend.preprocess(s["u8"], a.backbone, out=static_img)
            o = wl.model(static_img, s["kp"], s["crop"])
            s["done"].record(cur)
            if wl.gather is not None:
                o = wl.local(wl.gather.result() if wl.gather.attached else wl.gather(o))
            h_out[i & 1].copy_(o, non_blocking=True)                      # D2H of the step's result

    with torch.no_grad():
        loop(2)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall0 = time.time()
        t0.record()
        loop(steps)
        t1.record()
        barrier()
        wall1 = time.time()
    h2d = h_u8.numel() + h_kp.numel() * 4 + h_crop.numel() * 4
    d2h = h_out[0].numel() * 4
    wl.static["images"].copy_(wl.images.to(dev))               # restore the fp32 synthetic images for the later legs
    return t0.elapsed_time(t1), h2d, d2h, (wall0, wall1)


def qkv_in_graph_marginal(plan, qkv_idx, reps=12):
    """In-graph cost of the QKV GEMMs: the lifter section of the forward captured as a CUDA graph WITH and WITHOUT the QKV ops (their
    consumers then read stale buffers: timing is data-independent), replayed alternately after an L2 flush (so the weights are cold as
    in the real step while the LayerNorm output that feeds the GEMM is hot), CUDA events around each replay.  Returns the median
    difference per QKV launch in ms, or None.  Unlike events around a single launch this keeps the programmatic-dependent-launch
    overlap the real step has."""
    try:
        dev = plan.device
        nb, n = plan.prog.n_backbone_ops, len(plan.prog.ops)
        if not qkv_idx or min(qkv_idx) < nb:
            return None
        side = torch.cuda.Stream(dev)

        def ranges(skip):
            out, k = [], nb
            for q in sorted(skip) + [n]:
                if q > k:
                    out.append((k, q - k))
                k = q + 1
            return out

        def capture(skip):
            rs = ranges(skip)
            with torch.cuda.stream(side):
                for a, c in rs:
                    plan.run(a, c, stream=side)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for a, c in rs:
                    plan.run(a, c, stream=side)
            return g

        g_full, g_skip = capture([]), capture(qkv_idx)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        diffs = []
        for _ in range(reps):
            t = []
            for g in (g_full, g_skip):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize(dev)
                t.append(e0.elapsed_time(e1))
            diffs.append(t[0] - t[1])
        return statistics.median(diffs) / len(qkv_idx)
    except Exception:  # noqa: BLE001 -- an auxiliary number must never lose the bench line
        return None


def roofline_of(wl, ms_total, steps, peaks, ops_csv=None):
    """Dominant-kernel roofline from live per-op timing (see module docstring)."""
    import capf_b200
    plan = wl.plan
    B = wl.a.batch
    with torch.no_grad():
        op_ms = plan.time_ops(passes=3)
    kern = [plan.op_kernel(k) for k in range(len(plan.prog.ops))]
    groups, fam = {}, {}

    def shape_of(op):
        return "x".join(str(v) for v in (op.i[:11] if op.kind == capf_b200.lib.OP_CONV2D else op.i[:6]))

    for k, (op, ms) in enumerate(zip(plan.prog.ops, op_ms)):
        g = groups.setdefault((kern[k], shape_of(op)), {"ms": 0.0, "flops": 0, "bytes": 0, "launches": 0, "tag": op.tag})
        g["ms"] += ms; g["flops"] += op.flops; g["bytes"] += op.nbytes; g["launches"] += 1
        name = kern[k].split("[")[0].split("<")[0]
        d = fam.setdefault(name, {"ms": 0.0, "flops": 0, "launches": 0})
        d["ms"] += ms; d["flops"] += op.flops; d["launches"] += 1
    if ops_csv:
        os.makedirs(os.path.dirname(os.path.abspath(ops_csv)), exist_ok=True)
        with open(ops_csv, "w") as f:
            f.write("idx,kind,lane,kernel,tag,shape,ms,gflop,mbytes,tflops,gbps\n")
            for k, (op, ms) in enumerate(zip(plan.prog.ops, op_ms)):
                f.write(f"{k},{op.kind},{op.lane},\"{kern[k]}\",{op.tag},{shape_of(op)},{ms:.5f},{op.flops / 1e9:.4f},{op.nbytes / 1e6:.3f},"
                        f"{op.flops / max(ms, 1e-9) / 1e9:.2f},{op.nbytes / max(ms, 1e-9) / 1e6:.1f}\n")
    step_ms_sum = sum(op_ms)
    (dom_kernel, dom_shape), dom = max(groups.items(), key=lambda kv: kv[1]["ms"])
    peak_tf_sus = peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])
    peak_tf = peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"])
    peak_bw = peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"])
    intensity = dom["flops"] / max(dom["bytes"], 1)
    ridge = peak_tf_sus * 1e12 / (peak_bw * 1e9)
    dom_tflops = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
    dom_gbps = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
    traffic = traffic_detail = None          # measured DRAM bytes per launch (one ncu --set full capture), or null
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic_detail = tj.get(dom_kernel.split("[")[0]) or tj.get(dom_kernel)
        if traffic_detail:
            traffic = traffic_detail.get("traffic_bytes_per_launch")
    conv = [g for (kn, _), g in groups.items() if kn.startswith("tc_") or kn.startswith("stem_tc") or kn.startswith("lifter_")]
    conv_ms, conv_fl = sum(g["ms"] for g in conv), sum(g["flops"] for g in conv)
    # joint-block QKV GEMM (the path north_star quotes).  Headline = its in-step figure (per-op CUDA events of the in-order
    # pass above, operands in their in-step cache state) against the SUSTAINED bf16 peak; the figure from 20 back-to-back
    # launches per op against the burst peak is kept beside it.
    qkv_idx = [k for k, op in enumerate(plan.prog.ops) if "joint_blocks" in op.tag and op.tag.endswith("attn.qkv")]
    qkv = None
    if qkv_idx:
        qkv_fl = sum(plan.prog.ops[k].flops for k in qkv_idx)
        tf_step = qkv_fl / (sum(op_ms[k] for k in qkv_idx) * 1e-3) / 1e12
        tf_b2b = None
        try:
            with torch.no_grad():
                b2b_ms = [plan.time_op_repeated(k, 20) for k in qkv_idx]
            tf_b2b = qkv_fl / (sum(b2b_ms) * 1e-3) / 1e12
        except ValueError:
            pass
        with torch.no_grad():
            marg_ms = qkv_in_graph_marginal(plan, qkv_idx)
        tf_marg = (qkv_fl / len(qkv_idx)) / (marg_ms * 1e-3) / 1e12 if marg_ms and marg_ms > 0 else None
        qkv = {"achieved": tf_step, "peak": peak_tf_sus, "unit": "TFLOP/s", "frac": tf_step / peak_tf_sus,
               "in_graph_marginal": {"achieved": tf_marg, "peak": peak_tf_sus, "frac": (tf_marg / peak_tf_sus) if tf_marg else None,
                                     "us_per_launch": marg_ms * 1e3 if marg_ms else None,
                                     "how": "lifter section as a CUDA graph with vs without the 4 QKV launches, L2 flushed before every replay, median of 12 "
                                            "alternating pairs: the launch's cost inside the graph, programmatic-dependent-launch overlap included"},
               "how": "in-step: per-op CUDA events of one in-order pass of the whole forward", "peak_source": f"{peaks['_source']} bf16_tflops_sustained",
               "shape": f"M={B * 17} K=640 N=1920 x{len(qkv_idx)} blocks", "kernel": plan.op_kernel(qkv_idx[0]),
               "back_to_back": {"achieved": tf_b2b, "peak": peak_tf, "frac": (tf_b2b / peak_tf) if tf_b2b else None,
                                "how": "20 back-to-back launches per op, CUDA events; burst bf16 peak"}}
    hbm_bound = intensity < ridge
    return {
        "bound": "hbm" if hbm_bound else "tensor", "kernel": dom_kernel, "op_shape": dom_shape, "example_op": dom["tag"],
        "achieved": dom_gbps if hbm_bound else dom_tflops, "peak": peak_bw if hbm_bound else peak_tf_sus,
        "unit": "GB/s" if hbm_bound else "TFLOP/s",
        "frac": (dom_gbps / peak_bw) if hbm_bound else (dom_tflops / peak_tf_sus),
        "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)", "traffic_detail": traffic_detail,
        "peak_source": f"{peaks['_source']} " + ("hbm_gbs (copy, read+write bytes)" if hbm_bound else "bf16_tflops_sustained"),
        "arithmetic_intensity_flop_per_byte": intensity, "ridge_flop_per_byte": ridge,
        "launches_per_step": dom["launches"], "algorithmic_bytes_per_launch": dom["bytes"] / dom["launches"],
        "flops_per_launch": dom["flops"] / dom["launches"], "avg_launch_us": 1e3 * dom["ms"] / dom["launches"],
        "kernel_ms_per_step": dom["ms"], "share_of_step": dom["ms"] / step_ms_sum,
        "also_tflops": dom_tflops, "also_frac_of_tensor_peak": dom_tflops / peak_tf_sus,
        "tcgen05_kernels": {"achieved": conv_fl / (conv_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "peak": peak_tf_sus,
                            "frac": conv_fl / (conv_ms * 1e-3) / 1e12 / peak_tf_sus, "ms_per_step": conv_ms,
                            "share_of_step": conv_ms / step_ms_sum},
        "qkv_gemm": qkv,
        "whole_step": {"achieved": plan.prog.flops() / (ms_total / steps * 1e-3) / 1e12, "unit": "TFLOP/s",
                       "frac_of_tensor_peak": plan.prog.flops() / (ms_total / steps * 1e-3) / 1e12 / peak_tf_sus,
                       "flops_per_frame": plan.prog.flops() / B},
        "kernels_ms": {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
        "top_groups": [{"kernel": kn, "shape": sh, "ms": round(g["ms"], 4), "launches": g["launches"],
                        "tflops": round(g["flops"] / (g["ms"] * 1e-3) / 1e12, 1), "gbps": round(g["bytes"] / (g["ms"] * 1e-3) / 1e9, 1)}
                       for (kn, sh), g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])[:8]],
    }


def training_step_record(args, dev, steps=8):
    """one optimisation step of volume_net (train.py:186-201) on the configs[1] workload -- frozen backbone
    in the bench precision, lifter forward + backward in fp32 through capf_b200.train, fused AdamW -- next to the same step in
    PyTorch-eager CUDA (autograd over the oracle's lifter, autocast backbone, torch.optim.AdamW)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import capf_b200
    import capf_oracle
    from capf_b200 import train
    B, H, W = args.batch, args.height, args.width
    model, sd, cfg = build_model(args.backbone, args.precision, dev, graph=False)
    model.train(); model.backbone.eval(); model.volume_net.train()
    images, kp2d, crop = [t.to(dev) for t in capf_b200.synth.make_inputs(B, H, W, 1234)]
    gt = (torch.randn(B, 1, 17, 3, generator=torch.Generator().manual_seed(9)) * 0.3).to(dev)
    opt = train.FusedAdamW(model.volume_net.parameters(), lr=6.4e-4, weight_decay=0.1)

    def step():
        pred = model(images, kp2d, crop.clone())
        loss = torch.mean(torch.norm(pred - gt, dim=3))
        opt.zero_grad()
        loss.backward()
        opt.step()

    def timed(fn, n, warm=4):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    ms = timed(step, steps)
    rec = {"what": f"volume_net training step on the configs[1] workload: {args.backbone} frozen in {args.precision}, bs={B}, {H}x{W}; lifter forward + "
                   "backward (fp32, csrc/capf_train.cu) + fused AdamW; DropPath live", "value": B / ms * 1e3, "unit": "frames/s", "ms_per_step": ms, "steps": steps}
    del model, opt
    torch.cuda.empty_cache()
    if not args.no_yardstick:
        sdd = {k: v.to(dev) for k, v in sd.items()}
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sdd.items() if k.startswith("volume_net.") and v.is_floating_point()}
        sd2 = dict(sdd)
        sd2.update(leaves)
        topt = torch.optim.AdamW([{"params": list(leaves.values()), "lr": 6.4e-4}], weight_decay=0.1)
        adt = {"fp16": torch.float16, "bf16": torch.bfloat16}.get(args.precision, torch.float16)
        bench_flag = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = True

        def ystep():
            with torch.no_grad(), torch.autocast("cuda", dtype=adt):
                ref = capf_oracle.normalize_crop_(crop.clone())
                feats = [f.float() for f in (capf_oracle.cpn_forward(sdd, images.permute(0, 3, 1, 2)) if args.backbone == "cpn"
                                             else capf_oracle.hrnet_forward(sdd, images.permute(0, 3, 1, 2), cfg.model.backbone))]
            loss = torch.mean(torch.norm(capf_oracle.lifter_forward(sd2, kp2d, ref, feats) - gt, dim=3))
            topt.zero_grad()
            loss.backward()
            topt.step()

        try:
            yms = timed(ystep, max(3, steps // 2), warm=3)
            rec["pytorch_eager_yardstick"] = {"value": B / yms * 1e3, "unit": "frames/s", "ms_per_step": yms,
                                              "what": f"same step in PyTorch-eager CUDA: autocast({args.precision}) channels-last backbone under no_grad, "
                                                      "autograd over the lifter (eval-mode blocks), torch.optim.AdamW; library kernels only"}
        except Exception as e:  # noqa: BLE001
            rec["pytorch_eager_yardstick"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.backends.cudnn.benchmark = bench_flag
    return rec


def parity_of(wl, out, n):
    """rel-L2 / MPJPE of the first n frames of this rank against the fp32 CPU oracle (+ the library's fp32 mode)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import capf_oracle
    a = wl.a
    c = wl.crop[:n].clone()
    want = capf_oracle.ca_pf_forward(wl.weights, a.backbone, wl.cfg.model.backbone, wl.images[:n], wl.kp2d[:n], c)
    got = wl.local(out)[:n].cpu()
    per = [rel_l2(got[i], want[i]) for i in range(n)]
    parity = {"frames": n, "precision": a.precision, "rel_l2": rel_l2(got, want), "rel_l2_per_frame": per,
              "mpjpe_vs_ref_mm": float((got - want).norm(dim=-1).mean()) * 1000.0,
              "tolerance": "north_star: 1e-3 rel (fp32); 16-bit storage modes are reported against the fp32 oracle with "
                           "PyTorch-autocast's own deviation alongside (gpu_yardstick), BASELINE.md section 3"}
    return parity, want


def run_native(args, label):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libcapf_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:       # torchrun pins OMP_NUM_THREADS=1: weight folding / the parity oracle would crawl on one host thread
        torch.set_num_threads(max(1, (os.cpu_count() or world) // world))
    saved_stdout = None
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL writes its version banner to fd 1 at its first collective, so fd 1
        # points at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce_max(*vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    log("process group up, building the workload")
    wl = Workload(args, dev, rank, world, graph=not args.no_graph)
    B = args.batch
    log("plan built")
    ms_total, out, win1 = timed_resident(wl, args.steps, args.warmup, barrier)
    log(f"device-resident loop done: {ms_total / args.steps:.3f} ms/step")
    ms_e2e, h2d, d2h, win2 = timed_e2e(wl, args.steps, barrier)
    log(f"end-to-end loop done: {ms_e2e / args.steps:.3f} ms/step")
    clocks = sampler.stop([win1, win2]) if rank == 0 else None
    ms_total, ms_e2e = reduce_max(ms_total, ms_e2e)
    frames = B * world * args.steps
    with torch.no_grad():
        out = wl.step_resident().clone()             # result of the fp32 synthetic images (the e2e leg used uint8 crops)
    torch.cuda.synchronize(dev)

    line = None
    if rank == 0:
        peaks = load_peaks()
        roofline = roofline_of(wl, ms_total, args.steps, peaks, args.ops_csv)
        log("per-op timing done")
        cpu_base = parity = yard = None
        n = args.cpu_sample
        if not args.no_cpu:
            parity, want = parity_of(wl, out, n)
            log("parity vs the CPU oracle done")
            if world == 1:
                times, threads = time_cpu_reference(args.backbone, args.height, args.width, n, args.cpu_steps, 1)
                fps_cpu = n * len(times) / sum(times)
                cpu_base = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
                            "sample": f"oracle/capf_oracle.py (CPU restatement of the reference forward, fp32): {args.backbone}, bs={n}, "
                                      f"{args.height}x{args.width} (BASELINE configs[0] shape for config 1); {len(times)} timed forwards after 1 warm-up, all host threads",
                            "best_fps": n / min(times)}
                if args.precision not in ("fp32", "bf16x3"):
                    # the same frames through the fp32 parity mode of the library (CUDA-core kernels, the mode held to 1e-3)
                    with torch.no_grad():
                        m32, _, _ = build_model(args.backbone, "fp32", dev, graph=False)
                        got32 = m32(wl.images[:n].to(dev), wl.kp2d[:n].to(dev), wl.crop[:n].clone().to(dev)).cpu()
                    parity["fp32_mode_rel_l2"] = rel_l2(got32, want)
                    parity["fp32_mode_mpjpe_vs_ref_mm"] = float((got32 - want).norm(dim=-1).mean()) * 1000.0
                    del m32
                log("cpu baseline done")
                if not args.no_yardstick:
                    yard = gpu_yardstick(args.backbone, wl.cfg.model.backbone, wl.weights, wl.images, wl.kp2d, wl.crop,
                                         args.precision, dev, n, want)
        traffic_note = "profiles/traffic.json: constant from one ncu --set full capture (commit stamped in the file), not re-measured by this run"
        roofline["traffic_source"] = traffic_note
        line = {
            "metric": metric_name(args), "value": frames / (ms_total / 1e3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DT_TEXT[args.precision],
            "data": "synthetic (seeded randn images, random-init weights of the named architecture)",
            "config": {"workload": label, "baseline_config_index": args.config,
                       "global_batch": B * world, "parallelism": f"frame-sharded x{world}, NCCL all-gather of outputs"
                       + (" inside the CUDA graph" if (wl.gather is not None and wl.gather.attached) else "") if world > 1 else "single GPU",
                       "l2": f"inputs larger than L2: {wl.images.numel() * 4 / 1e6:.0f} MB of images per step (L2 = 126 MB); activations ~{wl.plan.workspace_bytes / 1e9:.1f} GB",
                       "cuda_graph": not args.no_graph, "precision": args.precision},
            "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "per step: pinned host uint8 BGR crops + f32 keypoints -> H2D on a copy stream (double-buffered, one step ahead) -> "
                            "frontend.preprocess (data_prefetcher.preload's image transform, one kernel into the plan's input) -> CA_PF.forward -> "
                            "async D2H of [B,1,17,3] into pinned memory; one synchronise at the end of the timed region"},
            "gpu_launches": args.steps * (wl.plan.num_launches + 1 + (1 if wl.gather is not None else 0)),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_base, "parity": parity, "gpu_yardstick": yard,
        }
    del wl
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs of this node size, shorter runs, same code path ----------------------------------
    others = []
    train_rec = None

    def emit(note=None):
        """Rank 0 prints THE line (the headline above is complete; the extras are whatever finished)."""
        if rank != 0:
            return
        line["other_configs"] = others
        line["train_step"] = train_rec
        if note:
            line["extras_note"] = note
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)

    # The extra records must never cost the headline: if they stall (a rank that failed inside one of them leaves its peers waiting in
    # the next collective), every rank's watchdog fires after the same limit, rank 0 prints the line without them and all exit 0.
    def watchdog_fire():
        log("extra records exceeded their time limit: printing the headline line without them")
        try:
            emit("the extra records (other_configs / train_step) did not finish within their time limit and were dropped")
        finally:
            os._exit(0)

    import threading
    watchdog = threading.Timer(float(os.environ.get("CAPF_BENCH_EXTRAS_LIMIT_S", "240")), watchdog_fire)
    watchdog.daemon = True
    watchdog.start()
    if args.config == 1 and not args.no_other_configs and label.find("custom") < 0:
        # (config index, precision override): configs[2] and the configs[1] workload in the fp32-class tensor-core mode at N = 1;
        # the two 8-GPU configs at N = 8
        todo = [(1, "bf16x3"), (2, None)] if world == 1 else ([(3, None), (4, None)] if world == 8 else [])
        if os.environ.get("CAPF_BENCH_FORCE_OTHERS"):          # testing: e.g. "3,4" exercises the 8-GPU list on 2 GPUs
            todo = [(int(v), None) for v in os.environ["CAPF_BENCH_FORCE_OTHERS"].split(",")]
        for k, prec in todo:
            a2 = argparse.Namespace(**vars(args))
            for key in ("backbone", "batch", "height", "width", "precision"):
                setattr(a2, key, CONFIGS[k][key])
            a2.config = k
            if prec:
                a2.precision = prec
            rec = {"baseline_config_index": k, "workload": workload_label(a2, k if not prec else None).replace(
                       "(custom: not a BASELINE config)", f"(BASELINE configs[{k}] workload in precision {prec}: the tensor-core mode held to 1e-3)"),
                   "metric": metric_name(a2)}
            log(f"other config {k} {prec or ''}")
            try:
                w2 = Workload(a2, dev, rank, world, graph=not args.no_graph)
                st = max(5, args.steps // 2)
                ms2, out2, _ = timed_resident(w2, st, 3, barrier)
                ms2e, h2d2, d2h2, _ = timed_e2e(w2, st, barrier)
                ms2, ms2e = reduce_max(ms2, ms2e)
                with torch.no_grad():
                    out2 = w2.step_resident().clone()
                torch.cuda.synchronize(dev)
                rec.update({"value": a2.batch * world * st / (ms2 / 1e3), "unit": "frames/s", "n_gpus": world, "steps": st, "warmup": 3,
                            "ms_per_step": ms2 / st, "dtype": DT_TEXT[a2.precision],
                            "e2e": {"value": a2.batch * world * st / (ms2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d2, "d2h_bytes_per_step": d2h2},
                            "whole_step_tflops": w2.plan.prog.flops() / (ms2 / st * 1e-3) / 1e12})
                if rank == 0 and not args.no_cpu:
                    rec["parity"], _ = parity_of(w2, out2, 4 if prec else 2)
                del w2
            except Exception as e:  # noqa: BLE001
                rec["error"] = f"{type(e).__name__}: {e}"[:300]
            torch.cuda.empty_cache()
            others.append(rec)
    if world == 1 and rank == 0 and args.config == 1 and not args.no_other_configs and label.find("custom") < 0:
        log("training step")
        try:
            train_rec = training_step_record(args, dev)
        except Exception as e:  # noqa: BLE001
            train_rec = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    watchdog.cancel()
    log("done")
    emit()
    if world > 1:
        if rank == 0 and saved_stdout is not None:
            os.dup2(2, 1)
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="BASELINE.json configs[k] (per-GPU shard for k = 3, 4)")
    ap.add_argument("--backbone", default=None, choices=["hrnet_32", "hrnet_48", "cpn"])
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--precision", default=None, choices=["fp16", "bf16", "fp32", "bf16x3"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity / yardstick legs")
    ap.add_argument("--no-yardstick", action="store_true", help="skip the PyTorch-eager-CUDA yardstick")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the other BASELINE configs of this node size")
    ap.add_argument("--cpu-sample", type=int, default=4, help="frames per CPU-baseline forward (BASELINE configs[0] uses 4)")
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--ops-csv", default=None, help="write the per-op device-time table (one in-order pass) to this CSV")
    ap.add_argument("--ref-sample", type=int, default=32, help="frames per step of the --impl reference arm")
    args = ap.parse_args()
    label = resolve_config(args)
    if args.impl == "reference":
        run_reference(args, label)
    else:
        run_native(args, label)


if __name__ == "__main__":
    main()
