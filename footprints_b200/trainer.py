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
import os

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
                 process_group=None, optimiser=None):
        """``optimiser``: an (otherwise unused) ``torch.optim.Adam`` over ``net.parameters()`` -- e.g.
        ``ModelManager.optimiser`` -- whose first param group supplies lr / betas / eps EVERY step, so the reference's
        ``StepLR`` scheduler (footprints/model_manager.py:28, stepped in train.py:191) drives this path unmodified."""
        self.net = net
        self.eng = net.train_engine()
        self._lr, self._betas, self._eps = lr, betas, eps
        self.optimiser = optimiser
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
        self._offs = offs
        views = {}
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                v = self.flat[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                views[p] = self.grad[o:o + p.numel()].view(p.shape)
        self.eng.grad_views = views
        # gradient buckets in the order the backward finishes them: decoders (the flat buffer's tail), encoder layer4,
        # layer3, then the rest (layer2, layer1, stem).  bounds[tag] = first flat element of that bucket.
        first = {}
        for L, o in self._layer_offsets(offs):
            tag = 4 if not L.name.startswith(('stem', 'enc')) else (3 if L.name.startswith('enc4') else
                                                                  (2 if L.name.startswith('enc3') else 0))
            first.setdefault(tag, o)
        self._bucket = {4: (first[4], total), 3: (first[3], first[4]), 2: (first[2], first[3])}
        self._rest_end = first[2]
        self._comm_stream = None
        self.step_count = 0
        self.lib = _lib.load()

    def _layer_offsets(self, offs):
        """(layer, flat offset of its first parameter) in the flat buffer's order."""
        i = 0
        for L in self.eng.layers.values():
            n = sum(1 for p in (L.weight, L.bias, L.bn.weight if L.bn is not None else None,
                                L.bn.bias if L.bn is not None else None) if p is not None)
            if i < len(offs):
                yield L, offs[i]
            i += n

    def _hyper(self, name, default):
        return self.optimiser.param_groups[0][name] if self.optimiser is not None else default

    lr = property(lambda self: self._hyper('lr', self._lr), lambda self, v: setattr(self, '_lr', v))
    betas = property(lambda self: self._hyper('betas', self._betas))
    eps = property(lambda self: self._hyper('eps', self._eps))

    # ---- checkpoint: torch.optim.Adam's state_dict format, so optimiser.pth is interchangeable with the reference's
    # (footprints/model_manager.py:40-56) and training resumes with the moments and bias correction intact
    def _param_indices(self):
        order = {id(p): i for i, p in enumerate(self.net.parameters())}
        return [order[id(p)] for p in self.params]

    def state_dict(self):
        state = {}
        if self.step_count > 0:
            for p, o, i in zip(self.params, self._offs, self._param_indices()):
                n = p.numel()
                state[i] = {'step': torch.tensor(float(self.step_count)),
                            'exp_avg': self.exp_avg[o:o + n].view(p.shape).clone(),
                            'exp_avg_sq': self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {'lr': self.lr, 'betas': tuple(self.betas), 'eps': self.eps, 'weight_decay': 0, 'amsgrad': False,
                 'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
                 'decoupled_weight_decay': False, 'params': list(range(len(list(self.net.parameters()))))}
        if self.optimiser is not None:
            group.update({k: v for k, v in self.optimiser.param_groups[0].items() if k != 'params'})
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        steps = set()
        with torch.no_grad():
            for p, o, i in zip(self.params, self._offs, self._param_indices()):
                st = sd['state'].get(i)
                if st is None:
                    continue
                n = p.numel()
                self.exp_avg[o:o + n].view(p.shape).copy_(st['exp_avg'])
                self.exp_avg_sq[o:o + n].view(p.shape).copy_(st['exp_avg_sq'])
                steps.add(int(st['step']))
        if len(steps) > 1:
            raise RuntimeError(f'per-parameter step counts differ ({sorted(steps)}); the flat Adam keeps one counter')
        self.step_count = steps.pop() if steps else 0
        g = sd['param_groups'][0]
        if self.optimiser is not None:
            self.optimiser.param_groups[0].update({k: g[k] for k in ('lr', 'betas', 'eps') if k in g})
        else:
            self._lr, self._betas, self._eps = g.get('lr', self._lr), tuple(g.get('betas', self._betas)), g.get('eps', self._eps)

    def broadcast_parameters(self):
        """Make every rank start from rank 0's weights (and BatchNorm buffers)."""
        self.sync.broadcast_([self.flat] + list(self.net.buffers()))

    # ------------------------------------------------------------------------------ the step
    def _hyper_host(self):
        b1, b2 = self.betas
        t = self.step_count
        return [self.lr, b1, b2, self.eps, 1.0 - b1 ** t, (1.0 - b2 ** t) ** 0.5, self.sync.grad_scale / self.eng.grad_scale]

    def _step_body(self, image, targets, adam_from_device):
        """forward -> fused loss+grad -> backward [-> all-reduce] -> Adam.  Pure stream work (no host sync), so it can be
        executed eagerly or captured into a CUDA graph."""
        outs, saved = self.eng.forward_train(image)
        losses21, grads = fused_loss([outs[k] for k in SCALES], targets, self.depth_range, self.prior, True)
        self.eng.launches += 2
        self._pending = []
        hook = self._reduce_decoder_bucket if (self.world > 1 and self.bucketed and not torch.cuda.is_current_stream_capturing()) else None
        self.eng.backward(dict(zip(SCALES, grads)), saved, on_decoders_done=hook)
        del saved, grads, outs
        if adam_from_device is None:
            return losses21
        self._adam(adam_from_device)
        return losses21

    # world > 1: all-reduce gradient buckets while the rest of the backward runs (FP_DP_BUCKETS=0: one all-reduce at the end)
    bucketed = os.environ.get('FP_DP_BUCKETS', '1') != '0' 

    def _reduce_decoder_bucket(self, side_stream, tag=4):
        """Called by the backward when a bucket's gradients have all been enqueued (tag 4: both decoders, 3: encoder layer4,
        2: layer3): all-reduce that slice of the flat gradient on a communication stream that waits for the main stream
        and the weight-gradient side stream -- NOT the other way round, so the rest of the backward runs underneath."""
        import torch.distributed as dist
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream()
        cs = self._comm_stream
        cs.wait_stream(torch.cuda.current_stream())
        if side_stream is not None:
            cs.wait_stream(side_stream)
        with torch.cuda.stream(cs):
            lo, hi = self._bucket[tag]
            self._pending.append(dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def _finish_all_reduce(self):
        """Sum the rest of the flat gradient over ranks (everything when no bucket went ahead) and join."""
        import torch.distributed as dist
        if self.world <= 1:
            return
        if self._pending:
            dist.all_reduce(self.grad[:self._rest_end], op=dist.ReduceOp.SUM, group=self.pg)
            for w in self._pending:
                w.wait()                      # the current stream waits for the bucket's collective
            torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._pending = []
        else:
            self.sync.all_reduce_(self.grad)

    def _push_hyper(self):
        """Adam's per-step scalars -> device (pinned staging, async copy ordered before the kernel on this stream).  Both the
        eager and the graph-replay path read them from device memory, so the two follow the same trajectory bit for bit."""
        if getattr(self, '_hyper_dev', None) is None:
            self._hyper_dev = torch.zeros(8, dtype=torch.float32, device=self.flat.device)
            # ring of pinned staging rows: the host runs ahead of the device, a single row would be overwritten before
            # the asynchronous copy of the previous step has read it
            self._hyper_pin = torch.zeros((256, 8), dtype=torch.float32).pin_memory()
        row = self._hyper_pin[self.step_count % 256]
        row[:7] = torch.tensor(self._hyper_host(), dtype=torch.float32)
        self._hyper_dev.copy_(row, non_blocking=True)

    def _adam(self, from_device):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if not from_device:          # eager call: refresh the scalars first (a graph replay did it before replaying)
            self._push_hyper()
        _lib.check(self.lib.fp_adam_step_dev(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                             self.exp_avg_sq.data_ptr(), self.flat.numel(), self._hyper_dev.data_ptr(),
                                             stream))
        self.eng.launches += 1

    def enable_cuda_graph(self, enabled=True):
        """Capture the step into a CUDA graph on the third call with a given input shape (the first two run eagerly: lazy
        allocations, function attributes) and replay it afterwards: ~450 kernel launches become one graph launch, so the
        host (ctypes calls, Python) is off the critical path.  Inputs are copied into static buffers before each replay;
        Adam's per-step scalars live in device memory.  With world > 1 the graph ends after the backward and the NCCL
        all-reduce + Adam follow eagerly (every rank must then take the same path on a given step: an eager step issues
        four bucketed collectives, a replayed one a single collective)."""
        self._graph_on = enabled
        self._graph = None
        self._graph_key = None
        self._eager_calls = 0
        return self

    def _capture(self, image, targets):
        dev = image.device
        self._static = {'image': image.clone()}
        self._static.update({k: v.clone() for k, v in targets.items()})
        self._push_hyper()
        in_graph_adam = self.world == 1
        l0 = self.eng.launches
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        # world > 1: NCCL's watchdog thread polls CUDA events while this thread captures -> thread-local capture mode
        with torch.cuda.graph(g, capture_error_mode='thread_local' if self.world > 1 else 'global'):
            out = self._step_body(self._static['image'], {k: self._static[k] for k in targets},
                                  True if in_graph_adam else None)
        self._graph, self._graph_out, self._graph_adam = g, out, in_graph_adam
        self._graph_launches = self.eng.launches - l0
        self.eng.launches = l0                     # capturing executed nothing

    def step(self, image, targets):
        """One training step on this rank's shard.  Returns the 21 losses as a device tensor (no host sync)."""
        if getattr(self, '_graph_on', False) and self.eng.profile is None:
            key = (tuple(image.shape), image.dtype, tuple(sorted(targets)))
            if self._graph is None or self._graph_key != key:
                if self._graph_key != key:
                    self._graph, self._graph_key, self._eager_calls = None, key, 0
                if self._eager_calls >= 2:
                    self._capture(image, targets)
            if self._graph is not None:
                for k, v in (('image', image), *targets.items()):
                    if v.data_ptr() != self._static[k].data_ptr():
                        self._static[k].copy_(v, non_blocking=True)
                self.step_count += 1
                if self._graph_adam:
                    self._push_hyper()
                self._graph.replay()
                self.eng.launches += self._graph_launches
                if not self._graph_adam:
                    self._finish_all_reduce()
                    self._adam(False)
                self.net._weights_epoch = getattr(self.net, '_weights_epoch', 0) + 1
                return self._graph_out
            self._eager_calls += 1
        losses21 = self._step_body(image, targets, None)
        self._finish_all_reduce()                                         # sum over ranks; Adam applies 1/world
        self.step_count += 1
        self._adam(False)
        self.net._weights_epoch = getattr(self.net, '_weights_epoch', 0) + 1     # eval engine must re-pack / re-fold
        return losses21
