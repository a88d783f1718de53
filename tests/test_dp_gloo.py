"""Data-parallel host logic on CPU: world_size 2 over gloo.  The per-rank arithmetic is the CUDA path (tested with
-m gpu); here: the gradient all-reduce + 1/world scaling reproduces the single-process gradient of the global
batch for a loss that is a mean over images, shards are disjoint and cover the batch, parameters broadcast."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from footprints_b200.dp import GradSync
    sync = GradSync(dist.group.WORLD)
    assert sync.world == world and abs(sync.grad_scale - 1.0 / world) < 1e-12
    g = torch.Generator().manual_seed(0)
    gb = 8
    x = torch.randn(gb, 5, generator=g)              # "global batch", identical on every rank
    w = torch.randn(5, generator=g)
    # rank 1 starts from different weights; broadcast makes them rank 0's
    w_local = w.clone() if rank == 0 else torch.zeros(5)
    sync.broadcast_([w_local])
    assert torch.equal(w_local, w)
    b, e = sync.shard(gb)
    assert (e - b) == gb // world
    w_local.requires_grad_(True)
    loss = ((x[b:e] @ w_local) ** 2).mean()          # per-rank mean loss over its shard
    loss.backward()
    flat = w_local.grad.clone()
    sync.all_reduce_(flat)
    flat *= sync.grad_scale
    w_ref = w.clone().requires_grad_(True)
    ((x @ w_ref) ** 2).mean().backward()             # single-process gradient of the global mean loss
    assert torch.allclose(flat, w_ref.grad, rtol=1e-6, atol=1e-7)
    with pytest.raises(ValueError):
        sync.shard(7)
    torch.save({'range': (b, e)}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


def test_gradient_allreduce_equals_global_batch_gradient(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ranges = sorted(torch.load(tmp_path / f'r{r}.pt')['range'] for r in range(world))
    assert ranges == [(0, 4), (4, 8)]


def test_single_process_is_identity():
    from footprints_b200.dp import GradSync
    s = GradSync(None)
    t = torch.ones(3)
    assert s.world == 1 and s.grad_scale == 1.0 and torch.equal(s.all_reduce_(t), torch.ones(3)) and s.shard(6) == (0, 6)
