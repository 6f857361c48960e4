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
