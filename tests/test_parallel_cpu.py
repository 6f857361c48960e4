"""World-size-2 (gloo, CPU) tests of the multi-GPU plumbing: sharding is a partition, the
gathered stack is in image order, uneven shards work."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dfnet_b200.parallel import render_images_sharded, shard_range


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (5, 2), (8, 8), (7, 4), (640 * 480, 8)])
def test_shard_range_is_a_partition(n, world):
    parts = [shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (a0, b0), (a1, b1) in zip(parts, parts[1:]):
        assert b0 == a1 and b0 >= a0
    sizes = [b - a for a, b in parts]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(n, world, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_img, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    poses = [torch.full((3, 4), float(i)) for i in range(n_img)]
    calls = []

    def render_one(i, pose):
        calls.append(i)
        return pose.sum() * torch.ones(2, 3, 3) + i  # stands in for an [H,W,3] image

    out = render_images_sharded(render_one, poses, rank, world)
    q.put((rank, calls, out.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_img", [4, 5, 1])
def test_sharded_render_gloo_world2(n_img):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_img, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    all_calls = sorted(res[0][1] + res[1][1])
    assert all_calls == list(range(n_img))            # every image rendered exactly once
    assert not set(res[0][1]) & set(res[1][1])
    want = torch.stack([torch.full((2, 3, 3), 12.0 * i + i) for i in range(n_img)])
    for _, _, out in res:
        assert torch.equal(out, want)                 # gathered in image order on every rank


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dfnet_b200.parallel import allreduce_gradients_
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Linear(5, 2))
    net[1].bias.requires_grad_(False)  # a frozen parameter is skipped on every rank alike
    for i, p in enumerate(net.parameters()):
        if p.requires_grad:
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = allreduce_gradients_(net.parameters())
    q.put((rank, n, [None if p.grad is None else p.grad.clone() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2():
    """train_on_batch data-parallel: one flat all-reduce averages the pose regressor's gradients."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in ps)
    for rank, n, grads in res:
        assert n == 4 * 3 * 9 + 4 + 2 * 5
        assert grads[3] is None
        for i, g in enumerate(grads[:3]):
            assert torch.allclose(g, torch.full_like(g, 1.5 * (i + 1)))  # mean of (1, 2) * (i + 1)


def _sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dfnet_b200 import parallel
    parallel.allreduce_stats(reset=True)
    sync = parallel.GradSync()
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
    views = [flat[:4].view(2, 2), flat[4:]]
    sync.launch(flat, 4)     # two buckets: head [0,4), tail [4,10)
    sync.finish()
    st = parallel.allreduce_stats()
    q.put((rank, flat.clone(), views[0].clone(), st["calls"], st["bytes_per_call"]))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_sync_flat_bucket_gloo_world2():
    """GradSync (the bucketed all-reduce train_on_batch overlaps with the pose regressor's backward) averages the flat
    bucket in place on every rank, so parameter views into it see the averaged values."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_sync_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in ps)
    want = torch.arange(10, dtype=torch.float32) * 1.5
    for rank, flat, v0, calls, nbytes in res:
        assert torch.equal(flat, want) and torch.equal(v0, want[:4].view(2, 2))
        assert calls == 1 and nbytes == 40
