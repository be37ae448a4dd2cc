"""GPU, 2 ranks over NCCL: the gathered [N*B,1,17,3] of the frame-sharded forward equals the single-GPU forward of the
same frames -- with the all-gather as a separate launch and as the last node of the forward's CUDA graph.
Skipped on a one-GPU box (the driver's 1-GPU test tier); `gpurun --gpus 2` runs it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _leave(q):
    """Flush the result queue and exit the worker process at once: tearing NCCL communicators and captured graphs down at
    interpreter exit was observed to stall for minutes on the 2-GPU box (the parent then waits for its non-daemon children)."""
    try:
        q.close()
        q.join_thread()
        torch.cuda.synchronize()
    finally:
        os._exit(0)


def _join(procs):
    for p in procs:
        p.join(30)
    for p in procs:
        if p.is_alive():
            p.terminate()


def _worker(rank, world, port, q):
    import capf_b200
    import protocol
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, H, W = 6, 64, 64
        cfg = capf_b200.make_config("hrnet_32")
        res = []
        for graph in (False, True):
            model = capf_b200.CA_PF(cfg, precision="fp16", use_cuda_graph=graph).eval()
            w = protocol.make_weights([(k, tuple(v.shape)) for k, v in model.state_dict().items()], 0)
            model.load_state_dict(w, strict=True)
            model = model.to(dev)
            images, kp2d, crop = protocol.make_inputs(world * B, H, W, 21)
            with torch.no_grad():
                # (a) sharded_forward: eager all_gather_into_tensor after the local forward
                full_a = capf_b200.dist.sharded_forward(model, images.to(dev), kp2d.to(dev), crop.clone().to(dev)).clone()
                # (b) the gather attached to the plan (inside the CUDA graph when graph=True)
                gat = capf_b200.dist.OutputGatherer([B] * world, (1, 17, 3), dev).attach(model, B, H, W)
                s, e = capf_b200.dist.shard_bounds(world * B, rank, world)
                for _ in range(2):          # second call replays the captured graph
                    model(images[s:e].to(dev), kp2d[s:e].to(dev), crop[s:e].clone().to(dev))
                    full_b = gat.result().clone()
                torch.cuda.synchronize(dev)
                # single-GPU forward of ALL frames on this rank
                solo = capf_b200.CA_PF(cfg, precision="fp16").eval()
                solo.load_state_dict(w, strict=True)
                want = solo.to(dev)(images.to(dev), kp2d.to(dev), crop.clone().to(dev))
            res.append((graph, bool(torch.equal(full_a, want)), bool(torch.equal(full_b, want)), tuple(full_b.shape)))
        q.put((rank, res))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"error: {type(e).__name__}: {e}"))
    finally:
        _leave(q)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_nccl_gather_equals_single_gpu_forward():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    _join(procs)
    for rank, rows in res:
        assert not isinstance(rows, str), rows
        for graph, eq_a, eq_b, shape in rows:
            assert eq_a and eq_b and shape == (12, 1, 17, 3), (rank, graph, eq_a, eq_b, shape)


def _ddp_worker(rank, world, port, q):
    """train.py:362 wraps the model in DistributedDataParallel: the hand-written backward must compose with DDP's gradient
    all-reduce (hooks on the leaf parameters), i.e. every rank ends with the MEAN of the per-shard gradients."""
    import capf_b200
    import protocol
    from torch.nn.parallel import DistributedDataParallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, H, W = 2, 64, 64
        cfg = capf_b200.make_config("hrnet_32")

        def make():
            m = capf_b200.CA_PF(cfg, precision="fp32")
            w = protocol.make_weights([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 0)
            m.load_state_dict(w, strict=True)
            m = m.to(dev)
            m.train()
            m.backbone.eval()
            m.volume_net.train()
            m.volume_net.drop_path_rate = 0.0
            return m

        images, kp2d, crop = protocol.make_inputs(world * B, H, W, 31)
        gt = torch.randn(world * B, 1, 17, 3, generator=torch.Generator().manual_seed(5)) * 0.3

        def loss_of(model, s, e):
            pred = model(images[s:e].to(dev), kp2d[s:e].to(dev), crop[s:e].clone().to(dev))
            return torch.mean(torch.norm(pred - gt[s:e].to(dev), dim=3))

        ddp = DistributedDataParallel(make(), device_ids=[rank])
        loss_of(ddp, rank * B, (rank + 1) * B).backward()
        got = {n: p.grad.detach().clone() for n, p in ddp.module.volume_net.named_parameters()}
        solo = make()
        for r in range(world):                       # same shards, one process: accumulate and average by hand
            (loss_of(solo, r * B, (r + 1) * B) / world).backward()
        worst = 0.0
        for n, p in solo.volume_net.named_parameters():
            worst = max(worst, float((got[n] - p.grad).norm() / p.grad.norm().clamp_min(1e-30)))
        q.put((rank, worst))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"error: {type(e).__name__}: {e}"))
    finally:
        _leave(q)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_training_step_under_ddp_averages_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    _join(procs)
    assert all(not isinstance(w, str) and w < 1e-5 for _, w in res), res
