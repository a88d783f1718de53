"""``LossManager`` — drop-in for ``footprints/training/losses.py:14-92`` of the reference.

Same constructor (``depth_range``, ``footprint_prior_weight``), same call contract
(``__call__(predictions, targets) -> dict`` with the reference's 21 keys in the reference's insertion
order, ``predictions`` updated in place with the visualisation tensors of losses.py:56-78,90), but the
~60 elementwise/reduce kernels of the reference are ONE fused loss+gradient kernel
(``fp_loss_fwd_bwd``); ``losses['loss'].backward()`` feeds the saved gradient into the network's
backward.  ``side_outputs=False`` skips the visualisation tensors (they are only read by the
tensorboard logger, footprints/training/logger.py:38-66).
"""
import ctypes

import torch

from . import _lib
from .engine import SCALES

_TARGETS = ('depth', 'visible_ground', 'all_ground', 'ground_depth', 'moving_object_mask', 'depth_mask')
_TERMS = ('visible_ground', 'all_ground', 'depth', 'ground_depth', 'loss')


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * 4)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def fused_loss(preds, targets, depth_range, prior_weight, want_grad):
    """preds: 4 fp32 [B,4,H,W] CUDA tensors ('1/8'..'1/1').  Returns (losses21, grads or None)."""
    lib = _lib.load()
    p0 = preds[0]
    b, _, h, w = p0.shape
    preds = [p.contiguous() for p in preds]
    tg = [targets[k].contiguous().float() for k in _TARGETS]
    for t in tg:
        if tuple(t.shape) != (b, h, w) or not t.is_cuda:
            raise RuntimeError(f'target planes must be CUDA [B,H,W]={b, h, w}, got {tuple(t.shape)}')
    out = torch.empty(21, dtype=torch.float32, device=p0.device)
    ws = torch.empty(1184 * 16, dtype=torch.float32, device=p0.device)
    grads = [torch.empty_like(p) for p in preds] if want_grad else None
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.fp_loss_fwd_bwd(_ptr_array(preds), _ptr_array(grads) if want_grad else None,
                                   *[ctypes.c_void_p(t.data_ptr()) for t in tg], b, h, w,
                                   float(depth_range[0]), float(depth_range[1]), float(prior_weight),
                                   ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(ws.data_ptr()), stream))
    return out, grads


class _LossFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, depth_range, prior_weight, targets, *preds):
        want = any(p.requires_grad for p in preds)
        out, grads = fused_loss(preds, targets, depth_range, prior_weight, want)
        ctx.grads = grads
        return out

    @staticmethod
    def backward(ctx, gvec):
        if ctx.grads is None:
            return (None, None, None, None, None, None, None)
        # the kernel stored d(losses[20])/d(pred); route an arbitrary upstream vector through the linear
        # structure losses[5s+4] = sum_c losses[5s+c], losses[20] = sum_s losses[5s+4] / 4
        g = gvec.view(-1)
        outs = []
        for s in range(4):
            coef = 4.0 * (g[5 * s:5 * s + 4] + g[5 * s + 4] + g[20] / 4.0)
            outs.append(ctx.grads[s] * coef.view(1, 4, 1, 1))
        ctx.grads = None
        return (None, None, None) + tuple(outs)


class LossManager:
    """ Class to compute losses for AllInOne data - i.e. data with ground truth labels for
    visible ground, visible depth, hidden ground and hidden depth (footprints/training/losses.py:14)"""

    def __init__(self, depth_range, footprint_prior_weight, side_outputs=True):
        self.min_depth, self.max_depth = depth_range
        self.footprint_prior_weight = footprint_prior_weight
        self.side_outputs = side_outputs

    def __call__(self, predictions, targets):
        keys = [k for k in predictions.keys() if isinstance(k, str)]
        if keys != list(SCALES):
            raise RuntimeError(f"predictions must hold the four scales {SCALES}, got {keys}")
        preds = [predictions[k] for k in keys]
        vec = _LossFn.apply((self.min_depth, self.max_depth), self.footprint_prior_weight, targets, *preds)
        losses = {}
        for s, k in enumerate(keys):
            for c, term in enumerate(_TERMS):
                losses[(term, k)] = vec[5 * s + c]
        losses['loss'] = vec[20]
        if self.side_outputs:
            # visualisation tensors the reference adds to `predictions` (losses.py:56,63,69,76-78); not on
            # the timed path — plain elementwise torch on detached outputs
            outputs = {}
            min_disp, max_disp = 1 / self.max_depth, 1 / self.min_depth
            with torch.no_grad():
                for k in keys:
                    o = predictions[k].detach()
                    outputs[('visible_ground', k)] = torch.sigmoid(o[:, 0])
                    outputs[('all_ground', k)] = torch.sigmoid(o[:, 1])
                    outputs[('depth', k)] = 1 / (min_disp + (max_disp - min_disp) * o[:, 2])
                    gd = 1 / (min_disp + (max_disp - min_disp) * o[:, 3])
                    outputs[('ground_depth', k)] = gd
                    outputs[('ground_depth_masked', k)] = gd * (outputs[('all_ground', k)] > 0.5).float()
            predictions.update(outputs)
        return losses
