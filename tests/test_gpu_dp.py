"""Data parallelism at the NETWORK level (SURVEY §8e's equivalence test): k ranks x b images, one gradient all-reduce per
step, must equal a single process that runs the k shards one after the other (BatchNorm statistics per shard, as with
the reference's plain BatchNorm2d under DDP), averages the k gradients and takes one Adam step.

Two processes share the one GPU of the test box; the collective runs over gloo on CUDA tensors (NCCL needs one GPU per
rank -- the multi-GPU NCCL path is exercised by `bench.py --gpus N`).  Both the bucketed eager path (the decoders' bucket
is all-reduced while the encoder's backward still runs) and the CUDA-graph path are covered."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
B, H, W = 2, 64, 96


def _worker(rank, world, port, out_dir, use_graph):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from footprints_b200 import FootprintNetwork
    from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict
    from footprints_b200.trainer import Trainer
    net = FootprintNetwork(pretrained=False)
    sd = synthetic_state_dict('resnet34', 0 if rank == 0 else 7)          # rank 1 starts elsewhere: broadcast fixes it
    net.load_state_dict(sd)
    net = net.cuda().train()
    tr = Trainer(net, lr=1e-3, process_group=dist.group.WORLD)
    if use_graph:
        tr.enable_cuda_graph()
    tr.broadcast_parameters()
    losses = []
    for step in range(4):
        d = synthetic_batch(world * B, H, W, 50 + step)                    # the global batch; this rank takes its shard
        b0, b1 = tr.sync.shard(world * B)
        img = d['image'][b0:b1].cuda()
        tg = {k: v[b0:b1].cuda() for k, v in d.items() if k != 'image'}
        losses.append(tr.step(img, tg)[20].item())
    torch.cuda.synchronize()
    torch.save({'flat': tr.flat.cpu(), 'losses': losses}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.destroy_process_group()


@pytest.mark.parametrize('use_graph', [False, True])
def test_two_ranks_equal_chunked_single_process(tmp_path, use_graph):
    from footprints_b200 import FootprintNetwork
    from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict
    from footprints_b200.trainer import Trainer
    world = 2
    port = 29500 + (os.getpid() % 2000) + (7 if use_graph else 0)
    mp.spawn(_worker, args=(world, port, str(tmp_path), use_graph), nprocs=world, join=True)
    got = [torch.load(tmp_path / f'r{r}.pt') for r in range(world)]
    assert torch.equal(got[0]['flat'], got[1]['flat'])                    # replicas stay bit-identical

    # single process: shards in sequence, gradients summed, one Adam step with 1/world
    net = FootprintNetwork(pretrained=False)
    net.load_state_dict(synthetic_state_dict('resnet34', 0))
    net = net.cuda().train()
    tr = Trainer(net, lr=1e-3)
    ref_losses = [[], []]
    for step in range(4):
        d = synthetic_batch(world * B, H, W, 50 + step)
        total = torch.zeros_like(tr.grad)
        for r in range(world):
            img = d['image'][r * B:(r + 1) * B].cuda()
            tg = {k: v[r * B:(r + 1) * B].cuda() for k, v in d.items() if k != 'image'}
            ref_losses[r].append(tr._step_body(img, tg, None)[20].item())
            total += tr.grad
        tr.grad.copy_(total)
        tr.step_count += 1
        h = tr._hyper_host()
        h[6] = (1.0 / world) / tr.eng.grad_scale
        tr._hyper_host = lambda h=h: h
        tr._adam(False)
        del tr._hyper_host
    torch.cuda.synchronize()
    for r in range(world):
        for a, b in zip(got[r]['losses'], ref_losses[r]):
            assert abs(a - b) <= 1e-6 * abs(b), (r, got[r]['losses'], ref_losses[r])
    ref = tr.flat.cpu()
    diff = (got[0]['flat'] - ref).abs().max().item()
    assert diff <= 2e-6, diff                # same kernels, same order of the two-term sum: agreement to fp32 rounding
