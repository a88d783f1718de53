"""``Segmentor`` -- drop-in for the ground-segmentation network of the reference's pre-processing stage,
``footprints/preprocessing/segmentation/network.py:13-104``.

Same torchvision ResNet-34 encoder and the same ``ConvBlock`` / ``ConvUpsampleAndConcatBlock`` decoder blocks as
``FootprintNetwork`` -- so every kernel is reused -- with ONE decoder, 1-channel heads (``OutConvBlock`` :158-170:
reflection pad + 3x3 convolution, no sigmoid, no in-network upsampling) and a LIST of four outputs
``[B,1,H/8,W/8], [B,1,H/4,W/4], [B,1,H/2,W/2], [B,1,H,W]`` (:83-101).  ``state_dict`` keys and shapes are the reference's
for both ``use_PSP=False`` (the constructor default) and ``use_PSP=True`` (the pyramid-pooling block of :174-207 in front
of the decoder: one fused kernel, ``fp_psp_fwd``).  Implemented: the eval forward (what ``segmentation/inference.py``
runs).  Not implemented: the training path of this network.
"""
import torch.nn as nn

from .engine import Engine
from .network import ConvBlock, ConvUpsampleAndConcatBlock, OutConvBlock, ResnetEncoder


class PSPBlock(nn.Module):
    """preprocessing/segmentation/network.py:174-188 -- parameter container (``reduce.weight`` [128,512,1,1]); the
    arithmetic of all four levels is ``fp_psp_fwd``."""

    def __init__(self, pool_size, feats, reduce_factor=4):
        super().__init__()
        self.pool_size = pool_size
        self.pooling = nn.AdaptiveAvgPool2d(output_size=(pool_size, pool_size))
        self.reduce = nn.Conv2d(feats, feats // reduce_factor, kernel_size=1, bias=False)


class PSP(nn.Module):
    """preprocessing/segmentation/network.py:191-207."""

    def __init__(self):
        super().__init__()
        self.block1 = PSPBlock(1, 512)
        self.block2 = PSPBlock(2, 512)
        self.block3 = PSPBlock(4, 512)
        self.block4 = PSPBlock(6, 512)


class SegSkipDecoder(nn.Module):
    """preprocessing/segmentation/network.py:56-101."""

    def __init__(self, use_PSP=False):
        super().__init__()
        self.use_PSP = use_PSP
        if use_PSP:
            self.PSP = PSP()
        self.block1 = ConvUpsampleAndConcatBlock(1024 if use_PSP else 512, 256)
        self.block2 = ConvUpsampleAndConcatBlock(256, 128)
        self.block3 = ConvUpsampleAndConcatBlock(128, 64)
        self.block4 = ConvUpsampleAndConcatBlock(64, 64)
        self.outconv1 = OutConvBlock(128, 1, 1, apply_sigmoid=False)
        self.outconv2 = OutConvBlock(64, 1, 1, apply_sigmoid=False)
        self.outconv3 = OutConvBlock(64, 1, 1, apply_sigmoid=False)
        self.outconv4 = nn.Sequential(ConvBlock(64, 32), OutConvBlock(32, 1, 1, apply_sigmoid=False))


class Segmentor(nn.Module):
    """preprocessing/segmentation/network.py:13-24."""

    def __init__(self, pretrained=True, use_PSP=False, blocks=(3, 4, 6, 3), compute_dtype='fp16'):
        super().__init__()
        self.encoder = ResnetEncoder(pretrained=pretrained, blocks=blocks)
        self.decoder = SegSkipDecoder(use_PSP=use_PSP)
        self.compute_dtype = compute_dtype
        self.precise = False
        self._engine = None

    def engine(self):
        if self._engine is None:
            self._engine = Engine(self, self.compute_dtype)
        return self._engine

    def forward(self, input_image):
        if not input_image.is_cuda:
            raise RuntimeError('footprints_b200.Segmentor runs on a B200 only (no CPU fallback); move the module and the input to cuda')
        if self.training:
            raise NotImplementedError('the training path of the segmentor is not implemented; call .eval()')
        return self.engine().forward(input_image)
