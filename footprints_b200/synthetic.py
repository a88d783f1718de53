"""Deterministic synthetic inputs, targets and weights for the Footprints hot path.

The input/target generators follow SURVEY.md §8(d) / BASELINE.md §3 (image in [0,1] like
``ToTensor`` — reference ``footprints/datasets/footprint_dataset.py:55-65``; target planes mimic
``footprints/datasets/kitti_dataset.py:63-119``).  Everything is generated on the CPU with an
explicit ``torch.Generator`` so that the same tensors are obtained on the build container and on
the GPU box, independent of any module constructor's RNG consumption.
"""
from collections import OrderedDict
import math

import torch

TARGET_KEYS = ('depth', 'visible_ground', 'all_ground', 'ground_depth',
               'moving_object_mask', 'depth_mask')

RESNET_BLOCKS = {'resnet18': (2, 2, 2, 2), 'resnet34': (3, 4, 6, 3)}


def synthetic_batch(batch, height, width, seed=1):
    """Image + the six target planes the loss reads (all float32, CPU)."""
    g = torch.Generator().manual_seed(seed)

    def u(*shape):
        return torch.rand(*shape, generator=g)

    def bern(p, *shape):
        return (torch.rand(*shape, generator=g) < p).float()

    s = (batch, height, width)
    out = {}
    out['image'] = u(batch, 3, height, width)
    out['depth'] = u(*s) * 80.0 * bern(0.9, *s)
    out['visible_ground'] = bern(0.4, *s)
    out['ground_depth'] = u(*s) * 40.0 * bern(0.5, *s)
    out['moving_object_mask'] = bern(0.05, *s)
    out['depth_mask'] = bern(0.1, *s)
    # footprints/datasets/footprint_dataset.py:64
    out['all_ground'] = ((out['ground_depth'] + out['visible_ground']) > 0).float()
    return out


def _conv_block_spec(prefix, cin, cout, spec):
    # footprints/network.py:104-122 (ConvBlock: conv1,bn1,conv2,bn2 — bn unused but present)
    for i, ci in ((1, cin), (2, cout)):
        spec[f'{prefix}.conv{i}.weight'] = (cout, ci, 3, 3)
        spec[f'{prefix}.conv{i}.bias'] = (cout,)
        _bn_spec(f'{prefix}.bn{i}', cout, spec)


def _bn_spec(prefix, c, spec):
    spec[f'{prefix}.weight'] = (c,)
    spec[f'{prefix}.bias'] = (c,)
    spec[f'{prefix}.running_mean'] = (c,)
    spec[f'{prefix}.running_var'] = (c,)
    spec[f'{prefix}.num_batches_tracked'] = ()


def state_spec(arch='resnet34'):
    """Ordered {key: shape} of ``FootprintNetwork.state_dict()`` (footprints/network.py:13-101,
    torchvision BasicBlock ResNet).  ``resnet34`` reproduces the reference's 484 keys."""
    blocks = RESNET_BLOCKS[arch]
    spec = OrderedDict()
    spec['encoder.layer0.0.weight'] = (64, 3, 7, 7)
    _bn_spec('encoder.layer0.1', 64, spec)
    widths = (64, 128, 256, 512)
    cin = 64
    for li, (nb, w) in enumerate(zip(blocks, widths)):
        # layer1 = Sequential(maxpool, layer1) → 'encoder.layer1.1.<b>'; others 'encoder.layerN.<b>'
        base = 'encoder.layer1.1' if li == 0 else f'encoder.layer{li + 1}'
        for b in range(nb):
            p = f'{base}.{b}'
            spec[f'{p}.conv1.weight'] = (w, cin if b == 0 else w, 3, 3)
            _bn_spec(f'{p}.bn1', w, spec)
            spec[f'{p}.conv2.weight'] = (w, w, 3, 3)
            _bn_spec(f'{p}.bn2', w, spec)
            if b == 0 and li > 0:
                spec[f'{p}.downsample.0.weight'] = (w, cin, 1, 1)
                _bn_spec(f'{p}.downsample.1', w, spec)
        cin = w
    for dec in ('mask_decoder', 'depth_decoder'):
        for bi, (ci, co) in enumerate(((512, 256), (256, 128), (128, 64), (64, 64)), start=1):
            _conv_block_spec(f'{dec}.block{bi}.pre_concat_conv', ci, co, spec)
            _conv_block_spec(f'{dec}.block{bi}.post_concat_conv', 2 * co, co, spec)
        for oi, ci in ((1, 128), (2, 64), (3, 64)):
            spec[f'{dec}.outconv{oi}.conv1.weight'] = (2, ci, 3, 3)
            spec[f'{dec}.outconv{oi}.conv1.bias'] = (2,)
        _conv_block_spec(f'{dec}.outconv4.0', 64, 32, spec)
        spec[f'{dec}.outconv4.1.conv1.weight'] = (2, 32, 3, 3)
        spec[f'{dec}.outconv4.1.conv1.bias'] = (2,)
    return spec


def segmentor_state_spec(arch='resnet34', use_PSP=False):
    """Ordered {key: shape} of the ground-segmentation network's ``state_dict``:
    footprints/preprocessing/segmentation/network.py:13-104 -- the same encoder, ONE skip decoder, 1-channel heads; with
    ``use_PSP`` the four 1x1 reduce convolutions of the pyramid-pooling block (:174-207) in front of a 1024-channel block1."""
    spec = OrderedDict((k, v) for k, v in state_spec(arch).items() if k.startswith('encoder.'))
    dec = 'decoder'
    if use_PSP:
        for i in (1, 2, 3, 4):
            spec[f'{dec}.PSP.block{i}.reduce.weight'] = (128, 512, 1, 1)
    for bi, (ci, co) in enumerate(((1024 if use_PSP else 512, 256), (256, 128), (128, 64), (64, 64)), start=1):
        _conv_block_spec(f'{dec}.block{bi}.pre_concat_conv', ci, co, spec)
        _conv_block_spec(f'{dec}.block{bi}.post_concat_conv', 2 * co, co, spec)
    for oi, ci in ((1, 128), (2, 64), (3, 64)):
        spec[f'{dec}.outconv{oi}.conv1.weight'] = (1, ci, 3, 3)
        spec[f'{dec}.outconv{oi}.conv1.bias'] = (1,)
    _conv_block_spec(f'{dec}.outconv4.0', 64, 32, spec)
    spec[f'{dec}.outconv4.1.conv1.weight'] = (1, 32, 3, 3)
    spec[f'{dec}.outconv4.1.conv1.bias'] = (1,)
    return spec


def synthetic_state_dict(arch='resnet34', seed=0, trained_like=True, spec=None):
    """Deterministic weights with non-trivial BatchNorm statistics.

    Encoder convs ~ N(0, 2/fan_out) (torchvision resnet.py kaiming fan_out init), decoder convs
    ~ U(±1/sqrt(fan_in)) (torch Conv2d default).  With ``trained_like`` the BN affine/running
    stats are randomised so that BN folding is actually exercised; otherwise they are the
    constructor defaults (1, 0, 0, 1)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape in (spec if spec is not None else state_spec(arch)).items():
        leaf = key.rsplit('.', 1)[1]
        if leaf == 'num_batches_tracked':
            sd[key] = torch.zeros((), dtype=torch.long)
        elif len(shape) == 4:
            cout, cin, kh, kw = shape
            if key.startswith('encoder'):
                std = math.sqrt(2.0 / (cout * kh * kw))
                sd[key] = torch.randn(shape, generator=g) * std
            else:
                b = 1.0 / math.sqrt(cin * kh * kw)
                sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif leaf == 'bias' and '.conv' in key:
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif not trained_like:
            one = leaf in ('weight', 'running_var')
            sd[key] = torch.ones(shape) if one else torch.zeros(shape)
        elif leaf == 'weight':
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif leaf == 'bias':
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif leaf == 'running_mean':
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif leaf == 'running_var':
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        else:
            raise KeyError(key)
    return sd
