"""B200-native training step without autograd in the loop (what ``bench.py`` times):

    forward (train-mode BN) -> fused loss+gradient -> hand-written backward -> [NCCL all-reduce of the flat
    fp32 gradient buffer over NVLink] -> fused Adam on the flat parameter buffer

It performs the arithmetic of the reference's hot loop (footprints/training/train.py:147-159:
``process_batch`` -> ``zero_grad`` -> ``backward`` -> ``optimiser.step``) with Adam's defaults from
footprints/model_manager.py:27.  Data parallelism shards the batch across ranks (one process per GPU); the only
collective is the gradient all-reduce (BatchNorm statistics stay per-GPU, as with the reference's plain
BatchNorm2d).  Parameters are re-pointed into one flat fp32 buffer so that the all-reduce and Adam are single
launches; ``net.state_dict()`` keeps working (the parameters are views).
"""
import ctypes

import torch

from . import _lib
from .dp import GradSync
from .engine import SCALES
from .losses import fused_loss


def trainable_parameters(net):
    """Parameters that receive a gradient: everything except the decoders' never-used BatchNorms
    (footprints/network.py:110,113 create them, use_bn=False skips them)."""
    eng = net.train_engine()
    out, seen = [], set()
    for L in eng.layers.values():
        for p in (L.weight, L.bias, L.bn.weight if L.bn is not None else None, L.bn.bias if L.bn is not None else None):
            if p is not None and id(p) not in seen:
                seen.add(id(p))
                out.append(p)
    return out


class Trainer:

    def __init__(self, net, lr=1e-4, depth_range=(0.1, 100.0), footprint_prior=0.25, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None):
        self.net = net
        self.eng = net.train_engine()
        self.lr, self.betas, self.eps = lr, betas, eps
        self.depth_range, self.prior = depth_range, footprint_prior
        self.sync = GradSync(process_group)
        self.pg, self.world = self.sync.pg, self.sync.world
        self.params = trainable_parameters(net)
        dev = self.params[0].device
        # 16-byte aligned slots in one flat buffer
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        views = {}
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                v = self.flat[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                views[p] = self.grad[o:o + p.numel()].view(p.shape)
        self.eng.grad_views = views
        self.step_count = 0
        self.lib = _lib.load()

    def broadcast_parameters(self):
        """Make every rank start from rank 0's weights (and BatchNorm buffers)."""
        self.sync.broadcast_([self.flat] + list(self.net.buffers()))

    def step(self, image, targets):
        """One training step on this rank's shard.  Returns the 21 losses as a device tensor (no host sync)."""
        outs = self.eng.forward_train(image)
        losses21, grads = fused_loss([outs[k] for k in SCALES], targets, self.depth_range, self.prior, True)
        self.eng.launches += 2
        self.eng.backward(dict(zip(SCALES, grads)))
        self.sync.all_reduce_(self.grad)                                  # sum over ranks; Adam applies 1/world
        self.step_count += 1
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.fp_adam_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                         self.exp_avg_sq.data_ptr(), self.flat.numel(), self.lr, self.betas[0],
                                         self.betas[1], self.eps, self.step_count, self.sync.grad_scale, stream))
        self.eng.launches += 1
        return losses21
