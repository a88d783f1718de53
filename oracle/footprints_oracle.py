"""CPU oracle for the Footprints encoder-decoder hot path.   *** TEST INFRASTRUCTURE ONLY ***

This file is a functional fp32 restatement (``torch.nn.functional`` on the CPU) of the reference
algorithm.  It is the checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``footprints_b200/`` imports it and the product has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against the *reference itself executed in the build container*:
``oracle/gen_golden.py`` imports ``/root/reference/footprints`` (network.py, training/losses.py),
loads the deterministic weights of ``footprints_b200.synthetic`` into it, and commits sampled
outputs / losses / gradient norms under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks
this restatement against those fixtures.

Each function cites the reference lines it restates (paths relative to /root/reference).  The
ResNet BasicBlock arithmetic is torchvision's (pinned 0.4.2 by ``environment.yml:8-9``; installed
0.26 has the same topology, ``torchvision/models/resnet.py:59-106,197-283``).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

RESNET_BLOCKS = {'resnet18': (2, 2, 2, 2), 'resnet34': (3, 4, 6, 3)}
SCALES = ('1/8', '1/4', '1/2', '1/1')


def _q(x, qdtype):
    """Optional storage-precision emulation (NOT part of the reference): round to the CUDA path's 16-bit
    storage type with a straight-through gradient.  With ``qdtype=None`` (default) the oracle is the fp32
    reference arithmetic.  Used by the gradient-parity tests so that ReLU/ELU masks are evaluated at the same
    point as on the GPU; a ReLU network's gradient is piecewise constant, so comparing against the fp32
    forward would measure forward rounding (mask flips), not the backward kernels."""
    if qdtype is None:
        return x
    return x + (x.to(qdtype).to(x.dtype) - x).detach()


def _bn(sd, prefix, x, training, new_stats):
    """torch BatchNorm2d, eps 1e-5, momentum 0.1 (torchvision resnet.py:82 default norm_layer)."""
    w, b = sd[f'{prefix}.weight'], sd[f'{prefix}.bias']
    rm, rv = sd[f'{prefix}.running_mean'], sd[f'{prefix}.running_var']
    if training:
        rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm, rv, w, b, True, 0.1, 1e-5)
        if new_stats is not None:
            new_stats[f'{prefix}.running_mean'] = rm
            new_stats[f'{prefix}.running_var'] = rv
        return y
    return F.batch_norm(x, rm, rv, w, b, False, 0.1, 1e-5)


def encoder_forward(sd, image, arch='resnet34', training=False, new_stats=None, taps=None, qdtype=None):
    """ResnetEncoder.forward — footprints/network.py:48-59; BasicBlock torchvision resnet.py:59-106."""
    blocks = RESNET_BLOCKS[arch]
    q = lambda t: _q(t, qdtype)                                              # identity unless emulating storage
    x = q((image - 0.45) / 0.225)                                            # network.py:50
    x = q(F.conv2d(x, q(sd['encoder.layer0.0.weight']), None, stride=2, padding=3))  # :52 (conv1)
    x = q(F.relu(_bn(sd, 'encoder.layer0.1', x, training, new_stats)))
    feats = [x]
    x = F.max_pool2d(x, 3, 2, 1)                                             # layer1[0] = maxpool
    for li, nb in enumerate(blocks):
        base = 'encoder.layer1.1' if li == 0 else f'encoder.layer{li + 1}'
        for b in range(nb):
            p = f'{base}.{b}'
            stride = 2 if (b == 0 and li > 0) else 1
            idt = x
            out = q(F.conv2d(x, q(sd[f'{p}.conv1.weight']), None, stride=stride, padding=1))
            out = q(F.relu(_bn(sd, f'{p}.bn1', out, training, new_stats)))
            out = q(F.conv2d(out, q(sd[f'{p}.conv2.weight']), None, stride=1, padding=1))
            out = _bn(sd, f'{p}.bn2', out, training, new_stats)
            if f'{p}.downsample.0.weight' in sd:
                idt = q(F.conv2d(x, q(sd[f'{p}.downsample.0.weight']), None, stride=2))
                idt = q(_bn(sd, f'{p}.downsample.1', idt, training, new_stats))
            x = q(F.relu(out + idt))
            if taps is not None:
                taps[p] = x
        feats.append(x)
    return feats


def _conv_block(sd, p, x, qdtype=None):
    """ConvBlock.forward with use_bn=False, ELU — footprints/network.py:124-138."""
    for i in (1, 2):
        x = F.pad(x, (1, 1, 1, 1), mode='reflect')
        x = _q(F.elu(F.conv2d(x, _q(sd[f'{p}.conv{i}.weight'], qdtype), sd[f'{p}.conv{i}.bias'])), qdtype)
    return x


def _up_block(sd, p, x, skip, qdtype=None):
    """ConvUpsampleAndConcatBlock.forward — footprints/network.py:151-158."""
    x = _conv_block(sd, f'{p}.pre_concat_conv', x, qdtype)
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    x = torch.cat([x, skip], 1)
    return _conv_block(sd, f'{p}.post_concat_conv', x, qdtype)


def _out_conv(sd, p, x, scale, apply_sigmoid, qdtype=None):
    """OutConvBlock.forward — footprints/network.py:174-183."""
    x = F.pad(x, (1, 1, 1, 1), mode='reflect')
    x = F.conv2d(x, _q(sd[f'{p}.conv1.weight'], qdtype), sd[f'{p}.conv1.bias'])
    if apply_sigmoid:
        x = torch.sigmoid(x)
    if scale != 1:
        x = F.interpolate(x, scale_factor=scale, mode='bilinear', align_corners=False)
    return x


def decoder_forward(sd, dec, feats, apply_sigmoid, taps=None, qdtype=None):
    """SkipDecoder.forward — footprints/network.py:82-101."""
    out = OrderedDict()
    x = _up_block(sd, f'{dec}.block1', feats[-1], feats[-2], qdtype)
    x = _up_block(sd, f'{dec}.block2', x, feats[-3], qdtype)
    out['1/8'] = _out_conv(sd, f'{dec}.outconv1', x, 8, apply_sigmoid, qdtype)
    x = _up_block(sd, f'{dec}.block3', x, feats[-4], qdtype)
    out['1/4'] = _out_conv(sd, f'{dec}.outconv2', x, 4, apply_sigmoid, qdtype)
    x = _up_block(sd, f'{dec}.block4', x, feats[-5], qdtype)
    if taps is not None:
        taps[f'{dec}.block4'] = x
    out['1/2'] = _out_conv(sd, f'{dec}.outconv3', x, 2, apply_sigmoid, qdtype)
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    x = _conv_block(sd, f'{dec}.outconv4.0', x, qdtype)
    out['1/1'] = _out_conv(sd, f'{dec}.outconv4.1', x, 1, apply_sigmoid, qdtype)
    return out


def network_forward(sd, image, arch='resnet34', training=False, new_stats=None, taps=None, qdtype=None):
    """FootprintNetwork.forward — footprints/network.py:21-30.  Returns OrderedDict of four
    [B,4,H,W] tensors keyed '1/8','1/4','1/2','1/1' (that insertion order)."""
    feats = encoder_forward(sd, image, arch, training, new_stats, taps, qdtype)
    if taps is not None:
        for i, f in enumerate(feats):
            taps[f'f{i}'] = f
    mask = decoder_forward(sd, 'mask_decoder', feats, False, taps, qdtype)
    depth = decoder_forward(sd, 'depth_decoder', feats, True, taps, qdtype)
    return OrderedDict((k, torch.cat([mask[k], depth[k]], 1)) for k in mask)


def psp_forward(sd, x, prefix='decoder.PSP'):
    """PSP.forward -- footprints/preprocessing/segmentation/network.py:174-207: four PSPBlocks (adaptive average pool to
    1/2/4/6 cells -> 1x1 conv 512->128 without bias -> bilinear upsample back, align_corners=True), concatenated behind x
    in the order [x, x6, x4, x2, x1]."""
    h, w = x.shape[2:]
    levels = []
    for i, s in ((1, 1), (2, 2), (3, 4), (4, 6)):
        y = F.conv2d(F.adaptive_avg_pool2d(x, (s, s)), sd[f'{prefix}.block{i}.reduce.weight'])
        levels.append(F.interpolate(y, size=(h, w), mode='bilinear', align_corners=True))
    x1, x2, x4, x6 = levels
    return torch.cat([x, x6, x4, x2, x1], 1)


def segmentor_forward(sd, image, arch='resnet34', training=False):
    """Segmentor.forward -- footprints/preprocessing/segmentation/network.py:20-24,79-101: the same
    encoder, one skip decoder, 1-channel heads (OutConvBlock :158-170: reflection pad + 3x3 conv, NO sigmoid, NO
    upsampling); the pyramid-pooling block in front of the decoder when the state dict holds its weights (use_PSP=True,
    :63-64,82-83).  Returns the list of four [B,1,h,w] outputs at 1/8, 1/4, 1/2 and full resolution."""
    feats = encoder_forward(sd, image, arch, training)
    if 'decoder.PSP.block1.reduce.weight' in sd:
        feats = list(feats)
        feats[-1] = psp_forward(sd, feats[-1])

    def head(p, x):
        return F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), sd[f'{p}.conv1.weight'], sd[f'{p}.conv1.bias'])

    outs = []
    x = _up_block(sd, 'decoder.block1', feats[-1], feats[-2])
    x = _up_block(sd, 'decoder.block2', x, feats[-3])
    outs.append(head('decoder.outconv1', x))
    x = _up_block(sd, 'decoder.block3', x, feats[-4])
    outs.append(head('decoder.outconv2', x))
    x = _up_block(sd, 'decoder.block4', x, feats[-5])
    outs.append(head('decoder.outconv3', x))
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    x = _conv_block(sd, 'decoder.outconv4.0', x)
    outs.append(head('decoder.outconv4.1', x))
    return outs


def sigmoid_to_depth(s, min_depth=0.1, max_depth=100.0):
    """footprints/utils.py:36-42."""
    min_disp, max_disp = 1 / max_depth, 1 / min_depth
    return 1 / (min_disp + (max_disp - min_disp) * s)


def _bce_logits(z, t):
    """BCEWithLogitsLoss(reduction='none') — footprints/training/losses.py:114,121."""
    return F.binary_cross_entropy_with_logits(z, t, reduction='none')


def loss_forward(predictions, targets, depth_range=(0.1, 100.0), prior_weight=0.25):
    """LossManager.__call__ — footprints/training/losses.py:31-92 (closed form in SURVEY §8 'L').
    Returns the same loss dict (21 entries) without mutating ``predictions``."""
    lo, hi = depth_range
    t_d, t_vis, t_all = targets['depth'], targets['visible_ground'], targets['all_ground']
    t_gd, dmask = targets['ground_depth'], targets['depth_mask']
    mv = 1 - targets['moving_object_mask']                                   # losses.py:44
    valid_d = (t_d > 0).float()
    valid_gd = (t_gd > 0).float()
    losses = OrderedDict()
    total = 0
    for k, o in predictions.items():
        if not isinstance(k, str):
            continue
        losses[('visible_ground', k)] = _bce_logits(o[:, 0], t_vis).mean()      # :53-55
        m = ((t_all + dmask) > 0).float()                                      # :140
        l3 = _bce_logits(o[:, 1], t_all) * m * mv                               # :141-145
        l3 = l3 + prior_weight * _bce_logits(o[:, 1], torch.zeros_like(t_all)) * (1 - m)  # :148-150
        losses[('all_ground', k)] = l3.mean()
        d = sigmoid_to_depth(o[:, 2], lo, hi)
        losses[('depth', k)] = (torch.log(torch.abs(d - t_d) + 1) * valid_d).mean()       # :105
        gd = sigmoid_to_depth(o[:, 3], lo, hi)
        losses[('ground_depth', k)] = (torch.log(torch.abs(gd - t_gd) + 1) * valid_gd).mean()
        losses[('loss', k)] = (losses[('depth', k)] + losses[('visible_ground', k)] +
                               losses[('all_ground', k)] + losses[('ground_depth', k)])   # :80-83
        total = total + losses[('loss', k)]
    losses['loss'] = total / 4                                                # :87-88
    return losses


def train_step_reference(sd, batch, arch='resnet34', depth_range=(0.1, 100.0), prior_weight=0.25, qdtype=None):
    """One forward(train-mode BN) + loss + backward (autograd over this restatement), i.e. the
    arithmetic of footprints/training/train.py:153-155,224-225.  Returns (outputs, losses, grads,
    new_running_stats).  Parameters that never receive a gradient (the unused decoder BNs,
    network.py:110,113) are absent from ``grads``."""
    params = OrderedDict()
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var')):
            params[k] = v.clone().requires_grad_(True)
        else:
            params[k] = v
    new_stats = {}
    outs = network_forward(params, batch['image'], arch, True, new_stats, qdtype=qdtype)
    losses = loss_forward(outs, batch, depth_range, prior_weight)
    losses['loss'].backward()
    grads = OrderedDict((k, v.grad) for k, v in params.items()
                        if v.is_floating_point() and v.requires_grad and v.grad is not None)
    return outs, losses, grads, new_stats
