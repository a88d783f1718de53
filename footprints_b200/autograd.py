"""autograd glue: makes the hand-written backward look like one differentiable op so that the reference's
training loop (``losses['loss'].backward(); optimiser.step()``, footprints/training/train.py:153-156) runs
unmodified on top of ``FootprintNetwork`` in ``.train()`` mode."""
import torch

from .engine import SCALES


class NetFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, net, image, *params):
        eng = net.train_engine()
        outs = eng.forward_train(image)
        ctx.net = net
        ctx.params = params
        return tuple(outs[k] for k in SCALES)

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.net.train_engine()
        douts = {}
        for k, g in zip(SCALES, gouts):
            if g is None:
                raise RuntimeError(f"no gradient reached output '{k}'; the backward needs all four scales")
            douts[k] = g.contiguous().float()
        grads = eng.backward(douts)
        return (None, None) + tuple(grads.get(p) for p in ctx.params)
