"""``Segmentor`` -- drop-in for the ground-segmentation network of the reference's pre-processing stage,
``footprints/preprocessing/segmentation/network.py:13-104``.

Same torchvision ResNet-34 encoder and the same ``ConvBlock`` / ``ConvUpsampleAndConcatBlock`` decoder blocks as
``FootprintNetwork`` -- so every kernel is reused -- with ONE decoder, 1-channel heads (``OutConvBlock`` :158-170:
reflection pad + 3x3 convolution, no sigmoid, no in-network upsampling) and a LIST of four outputs
``[B,1,H/8,W/8], [B,1,H/4,W/4], [B,1,H/2,W/2], [B,1,H,W]`` (:83-101).  ``state_dict`` keys and shapes are the reference's
(``use_PSP=False``, the constructor default).  Implemented: the eval forward (what ``segmentation/inference.py`` runs).
Not implemented: the PSP block (``use_PSP=True``) and the training path of this network.
"""
import torch.nn as nn

from .engine import Engine
from .network import ConvBlock, ConvUpsampleAndConcatBlock, OutConvBlock, ResnetEncoder


class SegSkipDecoder(nn.Module):
    """preprocessing/segmentation/network.py:56-101 (use_PSP=False)."""

    def __init__(self):
        super().__init__()
        self.use_PSP = False
        self.block1 = ConvUpsampleAndConcatBlock(512, 256)
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
        if use_PSP:
            raise NotImplementedError('the pyramid-pooling block (use_PSP=True) is not implemented')
        self.encoder = ResnetEncoder(pretrained=pretrained, blocks=blocks)
        self.decoder = SegSkipDecoder()
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
