"""B200-native ``FootprintNetwork`` — drop-in for ``footprints/network.py`` of the reference.

Same class names, constructor arguments, children (``encoder``, ``mask_decoder``,
``depth_decoder``), 484-key ``state_dict`` and output contract as the reference
(footprints/network.py:13-183): ``forward(input_image[B,3,H,W] fp32 in [0,1])`` returns a dict with
keys ``'1/8','1/4','1/2','1/1'`` (that order), each a mutable fp32 NCHW ``[B,4,H,W]`` tensor.

The modules below are *parameter containers only*: they carry the reference's parameters/buffers
under the reference's names and initialise them the same way, but none of them runs a PyTorch
operator in ``forward``.  All arithmetic happens in the sm_100a kernels of
``lib/libfootprints_b200.so`` driven by ``footprints_b200.engine``.  There is no CPU fallback.
"""
import torch
import torch.nn as nn

from .engine import Engine, SCALES


class _Placeholder(nn.Module):
    """Parameter-free slot (ReLU / MaxPool in the reference's nn.Sequential index layout)."""

    def forward(self, x):  # pragma: no cover - never called
        raise RuntimeError('footprints_b200 modules are parameter containers; call FootprintNetwork')


def _conv(cin, cout, k, stride=1, padding=0, bias=True):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=padding, bias=bias)


class BasicBlock(nn.Module):
    """torchvision BasicBlock parameter layout (torchvision/models/resnet.py:59-106)."""

    def __init__(self, cin, cout, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _conv(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = _Placeholder()
        self.conv2 = _conv(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = downsample
        self.stride = stride


def _make_layer(cin, cout, blocks, stride):
    downsample = None
    if stride != 1 or cin != cout:      # torchvision resnet.py:239-243
        downsample = nn.Sequential(_conv(cin, cout, 1, stride, 0, bias=False), nn.BatchNorm2d(cout))
    layers = [BasicBlock(cin, cout, stride, downsample)]
    for _ in range(1, blocks):
        layers.append(BasicBlock(cout, cout))
    return nn.Sequential(*layers)


class ResnetEncoder(nn.Module):
    """footprints/network.py:33-59.  ``blocks=(3,4,6,3)`` is the reference's ResNet-34;
    ``(2,2,2,2)`` is the ResNet-18 variant of BASELINE.json's configs."""

    def __init__(self, pretrained=True, blocks=(3, 4, 6, 3)):
        super().__init__()
        self.blocks = tuple(blocks)
        conv1 = _conv(3, 64, 7, 2, 3, bias=False)
        bn1 = nn.BatchNorm2d(64)
        layer1 = _make_layer(64, 64, blocks[0], 1)
        layer2 = _make_layer(64, 128, blocks[1], 2)
        layer3 = _make_layer(128, 256, blocks[2], 2)
        layer4 = _make_layer(256, 512, blocks[3], 2)
        # torchvision builds its (unused, deleted at network.py:46) fc = Linear(512, 1000) here; draw the same
        # random numbers so that a given torch seed yields the reference's random-init weights.
        nn.Linear(512, 1000)
        # torchvision resnet.py:208-213 init
        for m in [conv1, bn1, *layer1.modules(), *layer2.modules(), *layer3.modules(), *layer4.modules()]:
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.layer0 = nn.Sequential(conv1, bn1, _Placeholder())        # network.py:40
        self.layer1 = nn.Sequential(_Placeholder(), layer1)            # network.py:41 (maxpool, layer1)
        self.layer2 = layer2
        self.layer3 = layer3
        self.layer4 = layer4
        if pretrained:
            self._load_imagenet()

    def _load_imagenet(self):
        """The reference calls ``torchvision.models.resnet34(pretrained=True)`` (network.py:38), which
        downloads ImageNet weights; same behaviour (and the same failure when offline) here."""
        import torchvision
        name = {(3, 4, 6, 3): 'resnet34', (2, 2, 2, 2): 'resnet18'}[self.blocks]
        tv = getattr(torchvision.models, name)(weights='IMAGENET1K_V1')
        sd = tv.state_dict()
        mapped = {}
        for k, v in sd.items():
            if k.startswith('conv1.'):
                mapped['layer0.0.' + k[6:]] = v
            elif k.startswith('bn1.'):
                mapped['layer0.1.' + k[4:]] = v
            elif k.startswith('layer1.'):
                mapped['layer1.1.' + k[7:]] = v
            elif k.startswith(('layer2.', 'layer3.', 'layer4.')):
                mapped[k] = v
        self.load_state_dict(mapped)


class ConvBlock(nn.Module):
    """footprints/network.py:104-138 (bn1/bn2 exist but are unused because use_bn=False)."""

    def __init__(self, in_ch, out_ch, use_elu=True, use_bn=False):
        super().__init__()
        if use_bn or not use_elu:
            raise NotImplementedError('the reference only instantiates ConvBlock(use_elu=True, use_bn=False)')
        self.conv1 = _conv(in_ch, out_ch, 3)
        self.bn1 = nn.BatchNorm2d(out_ch)
        self.conv2 = _conv(out_ch, out_ch, 3)
        self.bn2 = nn.BatchNorm2d(out_ch)
        self.pad = _Placeholder()
        self.non_lin = _Placeholder()
        self.use_bn = use_bn


class ConvUpsampleAndConcatBlock(nn.Module):
    """footprints/network.py:141-158."""

    def __init__(self, in_ch, out_ch, use_elu=True, use_bn=False):
        super().__init__()
        self.pre_concat_conv = ConvBlock(in_ch, out_ch, use_elu, use_bn)
        self.post_concat_conv = ConvBlock(out_ch * 2, out_ch, use_elu, use_bn)


class OutConvBlock(nn.Module):
    """footprints/network.py:161-183."""

    def __init__(self, in_ch, out_ch, scale, apply_sigmoid=True):
        super().__init__()
        self.apply_sigmoid = apply_sigmoid
        self.pad = _Placeholder()
        self.scale = scale
        self.conv1 = _conv(in_ch, out_ch, 3)
        if apply_sigmoid:
            self.sigmoid = _Placeholder()


class SkipDecoder(nn.Module):
    """footprints/network.py:62-101."""

    def __init__(self, apply_sigmoid=True):
        super().__init__()
        self.apply_sigmoid = apply_sigmoid
        self.block1 = ConvUpsampleAndConcatBlock(512, 256)
        self.block2 = ConvUpsampleAndConcatBlock(256, 128)
        self.block3 = ConvUpsampleAndConcatBlock(128, 64)
        self.block4 = ConvUpsampleAndConcatBlock(64, 64)
        self.outconv1 = OutConvBlock(128, 2, 8, apply_sigmoid)
        self.outconv2 = OutConvBlock(64, 2, 4, apply_sigmoid)
        self.outconv3 = OutConvBlock(64, 2, 2, apply_sigmoid)
        self.outconv4 = nn.Sequential(ConvBlock(64, 32), OutConvBlock(32, 2, 1, apply_sigmoid))


class FootprintNetwork(nn.Module):
    """footprints/network.py:13-30.

    Extra keyword arguments (not in the reference) select the encoder depth and the 16-bit storage
    type of activations/weights inside the kernels: ``compute_dtype`` is ``'fp16'`` (default, 11-bit
    significand, closest to the reference's fp32) or ``'bf16'``; ``precise=True`` runs the eval forward's
    layers upstream of the low-resolution mask heads in split precision (hi + lo 16-bit pairs, three-term
    tensor-core products) so that ALL four output scales land within 1e-3 of the fp32 reference."""

    def __init__(self, pretrained=True, blocks=(3, 4, 6, 3), compute_dtype='fp16', precise=False):
        super().__init__()
        self.precise = precise          # split-precision layers upstream of the low-resolution mask heads (engine.py), eval and train forward
        self.encoder = ResnetEncoder(pretrained=pretrained, blocks=blocks)
        self.mask_decoder = SkipDecoder(apply_sigmoid=False)
        self.depth_decoder = SkipDecoder(apply_sigmoid=True)
        self.compute_dtype = compute_dtype
        self.train_dtype = 'fp16'      # activations AND inter-layer gradients in .train() mode ('fp16' + loss scale, or 'bf16')
        self._engine = None
        self._train_engine = None

    def engine(self):
        if self._engine is None:
            self._engine = Engine(self, self.compute_dtype, self.precise)
        return self._engine

    def train_engine(self):
        if self._train_engine is None:
            from .train_engine import TrainEngine
            self._train_engine = TrainEngine(self, self.train_dtype, self.precise)
        return self._train_engine

    def forward(self, input_image):
        if not input_image.is_cuda:
            raise RuntimeError('footprints_b200.FootprintNetwork runs on a B200 only (no CPU fallback); '
                               'move the module and the input to cuda')
        if self.training:
            # batch-statistics BatchNorm + hand-written backward behind one autograd node
            from collections import OrderedDict
            from .autograd import NetFunction
            params = [p for p in self.parameters() if p.requires_grad]
            outs = NetFunction.apply(self, input_image, *params)
            return OrderedDict(zip(SCALES, outs))
        return self.engine().forward(input_image)


    def predict_for_saving(self, input_image):
        """Output edge of the reference's batched inference (footprints/evaluation/inference.py:104-108,
        datasets/inference_dataset.py:35-39): the full-resolution prediction with the sigmoid applied to the two mask
        channels, as float16 ``[B,4,H,W]`` -- ``preds[:, 0:2] = sigmoid(preds[:, 0:2]); preds.cpu().numpy().astype(np.float16)``
        of the reference, fused into the kernel that writes the output (eval mode only)."""
        if self.training:
            raise RuntimeError('predict_for_saving is an inference call; use model.eval()')
        if not input_image.is_cuda:
            raise RuntimeError('footprints_b200.FootprintNetwork runs on a B200 only (no CPU fallback)')
        return self.engine().forward(input_image, post='inference_f16')


# BASELINE.json's north_star calls the class "FootprintsModel"; the reference symbol is FootprintNetwork.
FootprintsModel = FootprintNetwork
