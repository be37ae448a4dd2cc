#!/usr/bin/env python
"""bench.py -- frames/sec of the CA_PF lifting path (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps 20 --warmup 5                 # this arm (libcapf_b200)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W                  # N GPUs, one rank per GPU, NCCL
  python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 # the reference's CPU forward (oracle port)

A step is one CA_PF.forward over one batch of synthetic input.  Workload at N=1 = BASELINE.json configs[1]:
HRNet-32 + 4-level PoseFormer, 17 joints, bs=256, 256x256, fp16 storage / fp32 accumulate.  With N GPUs every rank
runs its own 256-frame shard (frames are independent, weak scaling) and the step ends with the NCCL all-gather of
the [256,1,17,3] outputs -- the only collective of the path (reference train.py:216-226).

One JSON line on stdout (rank 0).  `value`: device-resident inputs, CUDA-graph replay, CUDA events, max over ranks.
`e2e`: same metric through the public API with pinned HOST inputs: H2D copies + D2H of the result inside the timed
region.  `roofline`: the dominant kernel (largest summed device time, per-op CUDA events on the launching stream, one
in-order pass) -- its algorithmic bytes or FLOPs per launch divided by its average launch duration, against the measured
HBM / bf16 tensor peak of MEASURED_PEAKS.json, whichever bounds it at its arithmetic intensity.  `cpu_baseline`: the oracle (CPU restatement of the reference,
kind "port") timed on this box's host cores on a bounded sample.
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

METRIC = "frames/sec (17-joint, 256x256, bs=256)"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


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

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = self.rows
        window = "timed region"
        if t_begin is not None:
            inside = [r for r in rows if t_begin <= r[0] <= t_end + 0.06]
            if inside:
                rows = inside
            else:       # region shorter than the sampling period: the samples closest to it (under load: the warm-up / e2e loop)
                rows = sorted(rows, key=lambda r: abs(r[0] - 0.5 * (t_begin + t_end)))[:3]
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_sample
    times, threads = time_cpu_reference(args.backbone, args.height, args.width, sample, args.steps, args.warmup)
    total = sum(times)
    fps = sample * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.backbone} + 4-level PoseFormer, 17 joints, bs={args.batch}, {args.height}x{args.width}",
                   "sample": f"{sample} frames per step"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"oracle/capf_oracle.py (CPU restatement of the reference forward, torch-CPU fp32), "
                                   f"{sample}-frame batches of the same workload, {len(times)} timed steps, all host threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# this arm
# ------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist
    import capf_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libcapf_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
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
    B, H, W = args.batch, args.height, args.width
    model, weights, cfg = build_model(args.backbone, args.precision, dev, graph=not args.no_graph)
    images, kp2d, crop = capf_b200.synth.make_inputs(B, H, W, 1234 + rank)
    static = model.static_inputs(B, H, W, dev)
    plan = model.plan_for(B, H, W, dev)
    gather = capf_b200.dist.OutputGatherer([B] * world, (1, 17, 3), dev) if world > 1 else None

    static["images"].copy_(images.to(dev))
    kp_d, crop_pristine = kp2d.to(dev), crop.to(dev)
    crop_work = crop_pristine.clone()

    def step_resident():
        crop_work.copy_(crop_pristine)                 # forward normalises it in place (conpose.py:34-35)
        out = model(static["images"], kp_d, crop_work)
        return gather(out) if gather is not None else out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = step_resident()
        barrier()
        # ---- timed region 1: device-resident inputs ------------------------------------------------------
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        wall0 = time.time()
        e0.record()
        for _ in range(args.steps):
            out = step_resident()
        e1.record()
        barrier()
        wall1 = time.time()
        ms_total = e0.elapsed_time(e1)

        # ---- timed region 2: end to end from pinned host memory ------------------------------------------
        h_img, h_kp, h_crop = images.pin_memory(), kp2d.pin_memory(), crop.pin_memory()
        h_out = [torch.empty(B, 1, 17, 3).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(dev)
        stage = [dict(kp=torch.empty_like(kp_d), crop=torch.empty_like(crop_work), img=torch.empty_like(static["images"]),
                      ev=torch.cuda.Event(), done=torch.cuda.Event()) for _ in range(2)]

        def upload(slot):
            with torch.cuda.stream(copy_stream):
                s = stage[slot]
                s["img"].copy_(h_img, non_blocking=True)
                s["kp"].copy_(h_kp, non_blocking=True)
                s["crop"].copy_(h_crop, non_blocking=True)
                s["ev"].record(copy_stream)

        def e2e_loop(n):
            """Every step: H2D of its inputs from pinned host memory (copy stream, one step ahead of the compute),
            CA_PF.forward through the public API, D2H of its [B,1,17,3] result into pinned host memory.  The host never
            blocks inside the loop: slot reuse is ordered by events, the results are complete at the final synchronise."""
            cur = torch.cuda.current_stream(dev)
            upload(0)
            for i in range(n):
                s = stage[i & 1]
                if i + 1 < n:
                    if i >= 1:
                        copy_stream.wait_event(stage[(i + 1) & 1]["done"])    # step i-1 has consumed that slot
                    upload((i + 1) & 1)
                cur.wait_event(s["ev"])
                o = model(s["img"], s["kp"], s["crop"])
                s["done"].record(cur)
                if gather is not None:
                    o = gather(o)[rank * B:(rank + 1) * B]
                h_out[i & 1].copy_(o, non_blocking=True)                      # D2H of the step's result
            return h_out

        e2e_loop(2)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(args.steps)
        t1.record()
        barrier()
        ms_e2e = t0.elapsed_time(t1)
        clocks = sampler.stop(wall0, wall1) if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = t.tolist()

    frames = B * world * args.steps
    value = frames / (ms_total / 1e3)
    e2e_value = frames / (ms_e2e / 1e3)
    h2d = h_img.numel() * 4 + h_kp.numel() * 4 + h_crop.numel() * 4
    d2h = h_out[0].numel() * 4

    line = None
    if rank == 0:
        peaks = load_peaks()
        # ---- roofline of the dominant kernel (live per-op timing) ------------------------------------------------
        # Per-op device time: CUDA events around every op of one in-order pass on the launching stream.  Ops are
        # grouped by the kernel the library reports for them (capf_plan_op_kernel) and their operator shape; the group
        # with the largest summed time is "the dominant kernel".  Its bound follows from its arithmetic intensity
        # against the measured ridge (bf16 peak / HBM peak); `achieved` = algorithmic bytes (or FLOPs) of its launches
        # / their summed duration.
        with torch.no_grad():
            op_ms = plan.time_ops(passes=2)
        kern = [plan.op_kernel(k) for k in range(len(plan.prog.ops))]
        groups, fam = {}, {}
        for k, (op, ms) in enumerate(zip(plan.prog.ops, op_ms)):
            shape = "x".join(str(v) for v in op.i[:11]) if op.kind == capf_b200.lib.OP_CONV2D else "x".join(str(v) for v in op.i[:6])
            g = groups.setdefault((kern[k], shape), {"ms": 0.0, "flops": 0, "bytes": 0, "launches": 0, "tag": op.tag})
            g["ms"] += ms; g["flops"] += op.flops; g["bytes"] += op.nbytes; g["launches"] += 1
            name = kern[k].split("[")[0].split("<")[0]
            d = fam.setdefault(name, {"ms": 0.0, "flops": 0, "launches": 0})
            d["ms"] += ms; d["flops"] += op.flops; d["launches"] += 1
        if args.ops_csv:
            os.makedirs(os.path.dirname(os.path.abspath(args.ops_csv)), exist_ok=True)
            with open(args.ops_csv, "w") as f:
                f.write("idx,kind,lane,kernel,tag,shape,ms,gflop,mbytes,tflops,gbps\n")
                for k, (op, ms) in enumerate(zip(plan.prog.ops, op_ms)):
                    shape = "x".join(str(v) for v in op.i[:11]) if op.kind == capf_b200.lib.OP_CONV2D else "x".join(str(v) for v in op.i[:6])
                    f.write(f"{k},{op.kind},{op.lane},\"{kern[k]}\",{op.tag},{shape},{ms:.5f},{op.flops / 1e9:.4f},{op.nbytes / 1e6:.3f},"
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
                traffic_detail = json.load(f).get(dom_kernel.split("[")[0])
            if traffic_detail:
                traffic = traffic_detail.get("traffic_bytes_per_launch")
        conv = [g for (kn, _), g in groups.items() if kn.startswith("tc_") or kn.startswith("stem_tc")]
        conv_ms, conv_fl = sum(g["ms"] for g in conv), sum(g["flops"] for g in conv)
        # joint-block QKV GEMM (the path north_star quotes): 20 back-to-back launches of each of the 4 ops between two
        # events (its operands are L2-resident in the step as well); the in-step per-op figure is kept beside it
        qkv_idx = [k for k, op in enumerate(plan.prog.ops) if "joint_blocks" in op.tag and op.tag.endswith("attn.qkv")]
        qkv_tf = qkv_tf_step = None
        if qkv_idx:
            with torch.no_grad():
                qkv_ms = [plan.time_op_repeated(k, 20) for k in qkv_idx]
            qkv_fl = sum(plan.prog.ops[k].flops for k in qkv_idx)
            qkv_tf = qkv_fl / (sum(qkv_ms) * 1e-3) / 1e12
            qkv_tf_step = qkv_fl / (sum(op_ms[k] for k in qkv_idx) * 1e-3) / 1e12
        hbm_bound = intensity < ridge
        roofline = {
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
            "qkv_gemm": {"achieved": qkv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": (qkv_tf / peak_tf) if qkv_tf else None,
                         "shape": f"M={B * 17} K=640 N=1920 x4 blocks", "peak_source": f"{peaks['_source']} bf16_tflops (burst)",
                         "kernel": plan.op_kernel(qkv_idx[0]) if qkv_idx else None, "how": "20 back-to-back launches per op, CUDA events",
                         "in_step_per_op_events": qkv_tf_step},
            "whole_step": {"achieved": plan.prog.flops() / (ms_total / args.steps * 1e-3) / 1e12, "unit": "TFLOP/s",
                           "flops_per_frame": plan.prog.flops() / B},
            "kernels_ms": {k: round(v["ms"], 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])},
            "top_groups": [{"kernel": kn, "shape": sh, "ms": round(g["ms"], 4), "launches": g["launches"],
                            "tflops": round(g["flops"] / (g["ms"] * 1e-3) / 1e12, 1), "gbps": round(g["bytes"] / (g["ms"] * 1e-3) / 1e9, 1)}
                           for (kn, sh), g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])[:8]],
        }
        # ---- MPJPE vs reference on a small slice (parity carried with the number) ---------------------------
        cpu_base, parity = None, None
        if world == 1 and not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import capf_oracle
            n = args.cpu_sample
            times, threads = time_cpu_reference(args.backbone, H, W, n, args.cpu_steps, 1)
            fps_cpu = n * len(times) / sum(times)
            cpu_base = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": "port",
                        "sample": f"oracle/capf_oracle.py (CPU restatement of the reference forward, fp32) on BASELINE configs[0]: "
                                  f"{args.backbone}, bs={n}, {H}x{W}; {len(times)} timed forwards after 1 warm-up, all host threads",
                        "best_fps": n / min(times)}
            c = crop[:n].clone()
            want = capf_oracle.ca_pf_forward(weights, args.backbone, cfg.model.backbone, images[:n], kp2d[:n], c)
            got = out[:n].cpu() if gather is None else out[:n].cpu()
            parity = {"frames": n, "precision": args.precision, "rel_l2": float((got - want).norm() / want.norm()),
                      "mpjpe_vs_ref_mm": float((got - want).norm(dim=-1).mean()) * 1000.0}
            if args.precision != "fp32":
                # the same frames through the fp32 parity mode of the library (the mode that carries north_star's 1e-3 bar;
                # 16-bit storage of ~100 layers of a random-init network costs ~1.5e-3 -- PyTorch autocast on the reference
                # itself is at ~1e-2)
                with torch.no_grad():
                    m32, _, _ = build_model(args.backbone, "fp32", dev, graph=False)
                    got32 = m32(images[:n].to(dev), kp2d[:n].to(dev), crop[:n].clone().to(dev)).cpu()
                parity["fp32_mode_rel_l2"] = float((got32 - want).norm() / want.norm())
                parity["fp32_mode_mpjpe_vs_ref_mm"] = float((got32 - want).norm(dim=-1).mean()) * 1000.0
                del m32
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision] + " storage, f32 accumulate",
            "data": "synthetic (seeded randn images, random-init weights of the named architecture)",
            "config": {"workload": f"{args.backbone} + 4-level PoseFormer, 17 joints, bs={B} per GPU, {H}x{W} (BASELINE configs[1])",
                       "global_batch": B * world, "parallelism": f"frame-sharded x{world}, NCCL all-gather of outputs" if world > 1 else "single GPU",
                       "l2": f"inputs larger than L2: {h_img.numel() * 4 / 1e6:.0f} MB of images per step (L2 = 126 MB); activations ~{plan.workspace_bytes / 1e9:.1f} GB",
                       "cuda_graph": not args.no_graph, "precision": args.precision},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "note": "per step: pinned host f32 images/keypoints -> H2D on a copy stream (double-buffered, one step ahead) -> CA_PF.forward -> async D2H of [B,1,17,3] into pinned memory; one synchronise at the end of the timed region"},
            "gpu_launches": args.steps * (plan.num_launches + 1),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_base, "parity": parity,
        }
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
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
    ap.add_argument("--backbone", default="hrnet_32")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=256)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-sample", type=int, default=4, help="frames per CPU-baseline forward (BASELINE configs[0] uses 4)")
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--ops-csv", default=None, help="write the per-op device-time table (one in-order pass) to this CSV")
    ap.add_argument("--ref-sample", type=int, default=32, help="frames per step of the --impl reference arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
