"""Frame sharding across GPUs + the single collective of the path: the all-gather of the 3D outputs.

Frames are independent in eval mode, so the batch is split contiguously over ranks -- the same rule
the reference uses for its validation labels (mvn/datasets/human36m.py:536-552: ``n // world`` rows per rank, the
remainder on the last rank) -- and the only exchange is the gather of ``[N_r,1,17,3]`` predictions
(train.py:216-226: zero-pad every shard to the largest one, ``all_gather``, trim).  One process per GPU,
``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in the CPU tests); no other data-path collective.
"""
import torch
import torch.distributed as dist


def shard_sizes(n: int, world_size: int):
    """Rows per rank, reference rule (human36m.py:538-541)."""
    per = n // world_size
    return [per if r < world_size - 1 else n - per * (world_size - 1) for r in range(world_size)]


def shard_bounds(n: int, rank: int, world_size: int):
    """[start, end) of `rank`'s contiguous slice (human36m.py:542-543)."""
    per = n // world_size
    start = per * rank
    return start, (n if rank == world_size - 1 else start + per)


class OutputGatherer:
    """Reusable all-gather of per-rank predictions into one ``[sum(sizes), ...]`` tensor on every rank.

    Equal shards (the benchmark's case) use a single ``all_gather_into_tensor`` on a preallocated buffer;
    ragged shards follow the reference's pad/trim scheme."""

    def __init__(self, sizes, tail_shape=(1, 17, 3), device="cpu", dtype=torch.float32, group=None):
        self.sizes = list(sizes)
        self.group = group
        self.world = len(self.sizes)
        self.equal = len(set(self.sizes)) == 1
        self.max_rows = max(self.sizes)
        self.tail = tuple(tail_shape)
        self.buf = torch.zeros((self.world * self.max_rows,) + self.tail, device=device, dtype=dtype)
        self.pad = None if self.equal else torch.zeros((self.max_rows,) + self.tail, device=device, dtype=dtype)
        self.attached = False

    def attach(self, model, B, H, W):
        """Make the all-gather part of `model`'s forward for the (B, H, W) geometry: it is enqueued on the forward's
        stream right after the last kernel, reading the plan's output buffer in place -- and therefore becomes the last
        node of the forward's CUDA graph when the model replays one (no separate launch per step).  Equal shards only.
        After ``model(...)`` returns, ``result()`` is the gathered ``[world * B, ...]`` tensor (valid in stream order)."""
        if not self.equal or self.sizes[0] != B:
            raise ValueError("attach() needs equal shards of the plan's batch size")
        dev = self.buf.device
        plan = model.plan_for(B, H, W, dev)
        src = plan.tensor(plan.prog.outputs["out"]).view((B,) + self.tail)
        dist.all_gather_into_tensor(self.buf, src, group=self.group)      # communicator set-up happens here, outside any capture
        torch.cuda.synchronize(dev)

        def post(stream):
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(self.buf, src, group=self.group)

        plan.post = post
        plan._graph = None          # re-capture with the collective inside
        self.attached = True
        return self

    def result(self) -> torch.Tensor:
        return self.buf

    def __call__(self, local: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return local
        if self.equal:
            dist.all_gather_into_tensor(self.buf, local.contiguous(), group=self.group)
            return self.buf
        self.pad.zero_()
        self.pad[: local.shape[0]] = local
        chunks = list(self.buf.view((self.world, self.max_rows) + self.tail).unbind(0))
        dist.all_gather(chunks, self.pad, group=self.group)
        return torch.cat([c[:n] for c, n in zip(chunks, self.sizes)], dim=0)


def sharded_forward(model, images, kp2d, crop, gatherer=None):
    """Run `model` on this rank's contiguous slice of a *global* batch and gather the full prediction.
    All ranks pass the same global tensors (or at least their own slice at the right offsets)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n = images.shape[0]
    s, e = shard_bounds(n, rank, world)
    local = model(images[s:e].contiguous(), kp2d[s:e].contiguous(), crop[s:e])
    if world == 1:
        return local
    if gatherer is None:
        gatherer = OutputGatherer(shard_sizes(n, world), tuple(local.shape[1:]), local.device, local.dtype)
    return gatherer(local)
