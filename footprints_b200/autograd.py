"""autograd glue: makes the hand-written backward look like one differentiable op so that the reference's
training loop (``losses['loss'].backward(); optimiser.step()``, footprints/training/train.py:153-156) runs
unmodified on top of ``FootprintNetwork`` in ``.train()`` mode."""
import torch

from .engine import SCALES


class NetFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, net, image, *params):
        eng = net.train_engine()
        # the saved activations / batch statistics live on THIS node: two forwards followed by two backwards, gradient
        # accumulation over several forwards and train-mode forwards under no_grad all behave as with the reference
        need = any(p.requires_grad for p in params) or image.requires_grad
        outs, saved = eng.forward_train(image, save=need)
        ctx.net = net
        ctx.params = params
        ctx.saved = saved
        return tuple(outs[k] for k in SCALES)

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.net.train_engine()
        douts = {}
        for k, g in zip(SCALES, gouts):
            if g is None:
                raise RuntimeError(f"no gradient reached output '{k}'; the backward needs all four scales")
            douts[k] = g.contiguous().float()
        if ctx.saved is None:
            raise RuntimeError('backward through a FootprintNetwork forward that saved nothing (no_grad) or was already '
                               'back-propagated (retain_graph is not supported)')
        saved, ctx.saved = ctx.saved, None
        grads = eng.backward(douts, saved)
        outs = [grads.get(p) for p in ctx.params]
        if eng.grad_scale != 1.0:              # divide the (power-of-two) loss scale out: one multi-tensor launch
            torch._foreach_mul_([g for g in outs if g is not None], 1.0 / eng.grad_scale)
        return (None, None) + tuple(outs)
