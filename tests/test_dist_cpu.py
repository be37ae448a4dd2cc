"""CPU: the N>1 path (frame sharding + output all-gather) with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from capf_b200 import dist as cdist


def test_shard_rule_matches_reference():
    # human36m.py:536-552: n // world per rank, remainder on the last
    assert cdist.shard_sizes(10, 4) == [2, 2, 2, 4]
    assert [cdist.shard_bounds(10, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 10)]
    assert cdist.shard_sizes(2048, 8) == [256] * 8
    assert cdist.shard_sizes(3, 4) == [0, 0, 0, 3]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        images = torch.randn(n, 4, 4, 3, generator=g)
        kp = torch.randn(n, 17, 2, generator=g)
        crop = torch.rand(n, 17, 2, generator=g) * 100

        def fake_model(im, k, c):            # stands in for CA_PF (GPU-only): deterministic per-frame function
            c /= 2.0                         # in-place on the caller's slice, like conpose.py:34-35
            return (k.sum(-1, keepdim=True) + im.mean((1, 2, 3)).view(-1, 1, 1)).unsqueeze(1).expand(-1, 1, 17, 3).contiguous()

        full = cdist.sharded_forward(fake_model, images, kp, crop.clone())
        want = fake_model(images, kp, crop.clone())
        q.put((rank, bool(torch.equal(full, want)), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])      # equal shards (all_gather_into_tensor) and ragged (pad / trim, train.py:216-226)
def test_sharded_forward_gathers_full_batch(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True, (n, 1, 17, 3)), (1, True, (n, 1, 17, 3))]
