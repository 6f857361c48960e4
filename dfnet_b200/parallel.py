"""Multi-GPU plumbing for the render path: one process per GPU, images (or ray blocks) sharded
across ranks, no data-path collective (rays are independent; weights are replicated).
torch.distributed is used only for rendezvous, barriers and gathering small results."""
import os

import torch
import torch.distributed as dist


def dist_info():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_items, rank, world):
    """Contiguous, balanced partition of range(n_items): the first n_items % world ranks get one
    extra item.  Returns (start, stop)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def render_images_sharded(render_one, poses, rank=None, world=None, gather=True):
    """Render `poses[i]` for the images of this rank with `render_one(i, pose) -> Tensor[H,W,C]`
    and (optionally) all-gather the results so that every rank returns the full stack in image
    order (the reference's render_path returns all images, rendering.py:403-458).
    Ranks may hold different numbers of images; padding keeps the collective shapes equal."""
    if rank is None or world is None:
        rank, world, _ = dist_info()
    start, stop = shard_range(len(poses), rank, world)
    mine = [render_one(i, poses[i]) for i in range(start, stop)]
    if world == 1 or not gather:
        return torch.stack(mine) if mine else None
    counts = [shard_range(len(poses), r, world) for r in range(world)]
    max_n = max(b - a for a, b in counts)
    # rank 0 always owns image 0: it tells ranks without images (len(poses) < world) the image shape
    meta = [(list(mine[0].shape), mine[0].dtype) if rank == 0 else None]
    dist.broadcast_object_list(meta, src=0)
    shape, dt = meta[0]
    dev = mine[0].device if mine else (poses[0].device if isinstance(poses[0], torch.Tensor) else torch.device("cpu"))
    pad = torch.zeros([max_n] + shape, device=dev, dtype=dt)
    for k, m in enumerate(mine):
        pad[k] = m
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([out[r][: b - a] for r, (a, b) in enumerate(counts)], 0)


def allreduce_gradients_(params, group=None):
    """Data-parallel training step: average the gradients of `params` over the ranks with ONE
    all-reduce of a flat fp32 bucket (the pose regressor's 15.4 M parameters = 61.6 MB; NeRF and feature net are
    frozen).  In place; parameters without gradient are skipped identically on every rank."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()


# ---------------------------------------------------------------------------------------------------------------------
# Gradient all-reduce as the tail of the pose regressor's backward
# ---------------------------------------------------------------------------------------------------------------------
_STATS = {"calls": 0, "bytes": 0, "ms": 0.0, "exposed_ms": 0.0, "timed_calls": 0, "pending": []}


def allreduce_stats(reset=False):
    """Accumulated all-reduce measurements of GradSync (bench.py): device time of the collectives and the time the
    compute stream actually waited for them, from CUDA events (resolved lazily here, never inside a step)."""
    for ev in _STATS["pending"]:
        a0, a1, w0, w1 = ev
        a1.synchronize(), w1.synchronize()
        _STATS["ms"] += a0.elapsed_time(a1)
        _STATS["exposed_ms"] += w0.elapsed_time(w1)
        _STATS["timed_calls"] += 1
    _STATS["pending"] = []
    out = {k: v for k, v in _STATS.items() if k != "pending"}
    out["bytes_per_call"] = _STATS["bytes"] / max(_STATS["calls"], 1)
    if reset:
        _STATS.update(calls=0, bytes=0, ms=0.0, exposed_ms=0.0, timed_calls=0)
    return out


class GradSync:
    """Averages ONE flat fp32 gradient bucket over the ranks while the backward that fills it is still running.

    dfb_dfnet_bwd writes every parameter gradient of the pose regressor into one flat buffer and records an event when
    the tail of that buffer (fc_pose and the deep, parameter-heavy encoder layers, differentiated first) is complete.
    `launch` enqueues the collective for that tail on a side stream behind the event and the collective for the head
    behind the end of the backward; `finish` makes the compute stream wait for both.  The result is the average in
    place, so the optimizer reads it without another copy.  Counted as one logical all-reduce of the bucket per step
    (two NCCL launches).  CPU tensors (gloo tests) take the same path without streams."""

    def __init__(self, group=None):
        self.group = group
        self.side = None
        self._pending = None

    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def launch(self, flat, split, tail_ready=None):
        """flat: the bucket; flat[split:] is complete when `tail_ready` (torch.cuda.Event) fires, flat[:split] when the
        work enqueued so far on the current stream is done."""
        w = self.world()
        if w == 1:
            return
        _STATS["calls"] += 1
        _STATS["bytes"] += flat.numel() * 4
        if not flat.is_cuda:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat /= w
            return
        main = torch.cuda.current_stream(flat.device)
        if self.side is None:
            self.side = torch.cuda.Stream(flat.device)
        done_all = torch.cuda.Event()
        done_all.record(main)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.side):
            self.side.wait_event(tail_ready if tail_ready is not None else done_all)
            a0.record(self.side)
            if split < flat.numel():
                dist.all_reduce(flat[split:], op=dist.ReduceOp.AVG, group=self.group)
            self.side.wait_event(done_all)
            if split > 0:
                dist.all_reduce(flat[:split], op=dist.ReduceOp.AVG, group=self.group)
            a1.record(self.side)
        self._pending = (a0, a1, flat)

    def finish(self):
        """The current stream waits for the collectives launched by `launch`."""
        if self._pending is None:
            return
        a0, a1, flat = self._pending
        self._pending = None
        main = torch.cuda.current_stream(flat.device)
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(main)
        main.wait_stream(self.side)
        w1.record(main)
        flat.record_stream(self.side)
        if len(_STATS["pending"]) < 64:
            _STATS["pending"].append((a0, a1, w0, w1))
