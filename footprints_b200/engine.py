"""Host-side executor of the Footprints hot path on one B200.

The engine turns a ``FootprintNetwork`` (parameter containers) into a flat plan of C-ABI calls into
``libfootprints_b200.so`` and runs it on the current CUDA stream.  PyTorch is used for device
memory (activation/weight buffers), streams and — in the data-parallel path — NCCL; none of the
network arithmetic is a PyTorch operator.

Layout: activations are NHWC 16-bit (fp16 or bf16); weights are packed K-major
``[cout_pad][taps*cin]`` 16-bit (see ``fp_pack_conv_weight``); eval-mode BatchNorm is folded into a
per-channel fp32 scale/shift applied in the convolution epilogue.
"""
import ctypes
import os
from collections import OrderedDict

import torch

from . import _lib
from ._lib import (FP_ACT_ELU, FP_ACT_NONE, FP_ACT_RELU, FP_ACT_SIGMOID, FP_BF16, FP_F16, FP_PAD_REFLECT,
                   FP_PAD_ZERO, FpConvDesc)

SCALES = ('1/8', '1/4', '1/2', '1/1')


def _bn_tile(cout):
    if cout >= 256:
        return 256
    for bn in (16, 32, 64, 128):
        if cout <= bn:
            return bn
    return 128


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class ConvLayer:
    """One convolution of the plan: static geometry + the parameters it reads."""

    def __init__(self, name, weight, cin0, cin1=0, k=3, stride=1, pad=1, pad_mode=FP_PAD_ZERO, up0=0, act=FP_ACT_NONE,
                 bias=None, bn=None, out_f32=False, stem=False):
        self.name = name
        self.weight = weight          # nn.Parameter [cout, cin, k, k]
        self.bias = bias              # nn.Parameter or None
        self.bn = bn                  # nn.BatchNorm2d or None
        self.cin0, self.cin1 = cin0, cin1
        self.k, self.stride, self.pad, self.pad_mode, self.up0 = k, stride, pad, pad_mode, up0
        self.act, self.out_f32, self.stem = act, out_f32, stem
        self.is_head = out_f32            # 2-channel output heads run on dedicated CUDA-core kernels
        self.cout = weight.shape[0]
        self.bn_tile = _bn_tile(self.cout)
        self.npad = (self.cout + self.bn_tile - 1) // self.bn_tile * self.bn_tile
        ktot = 256 if stem else k * k * (cin0 + cin1)
        self.kpad = (ktot + 63) // 64 * 64
        # filled by Engine.refresh_weights()
        self.w_packed = None
        self.scale = None
        self.shift = None


# layers computed in split precision (hi + lo operand pairs, three-term products) in ``precise`` mode: everything upstream
# of the low-resolution MASK heads, whose raw logits dominate the output error (scripts/precision_study.py: with this set
# every output scale lands within 4e-4 of the fp32 reference; splitting only the deep stages does not help because the
# 16-bit rounding error accumulates evenly over the ~40 layers in front of the '1/8' head)
PRECISE_PREFIXES = ('stem', 'enc', 'mask.block1', 'mask.block2', 'mask.head1')


class Engine:
    def __init__(self, net, compute_dtype='fp16', precise=False):
        self.net = net
        self.precise = bool(precise)
        self.lib = _lib.load()
        self.fmt = {'fp16': FP_F16, 'bf16': FP_BF16}[compute_dtype]
        self.tdtype = {FP_F16: torch.float16, FP_BF16: torch.bfloat16}[self.fmt]
        self.layers = OrderedDict()
        self._build_layers()
        self._weight_key = None
        self._bufs = {}
        self.head_tensor_cores = True
        self.nvtx = os.environ.get('FP_NVTX', '0') == '1'      # per-launch NVTX ranges (profiling)
        self.launches = 0           # kernels launched through the C ABI (bench.py reports them)
        self.profile = None         # when a list: (kind, layer, flops, start_event, end_event, algorithmic_bytes) per conv launch
        self._prof_open = None

    # ------------------------------------------------------------------------------ instrumentation
    def _prof(self, kind, L, m_pixels, abytes=0):
        """abytes: algorithmic HBM bytes of the launch (unique input elements once, pre-upsample / pre-pad,
        the concatenated skip included, + outputs once + weights once), 0 when the caller does not track it."""
        if self.nvtx:                   # NVTX range "<kind>:<layer>" around the launch (ncu --nvtx --nvtx-include ...)
            torch.cuda.nvtx.range_push(f'{kind}:{L.name}')
        if self.profile is None:
            return
        k_real = 147 if L.stem else L.k * L.k * (L.cin0 + L.cin1)
        flops = 2.0 * m_pixels * L.cout * k_real
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        self._prof_open = (kind, L.name, flops, e0, e1, abytes)

    def _prof_end(self):
        if self.nvtx:
            torch.cuda.nvtx.range_pop()
        if self._prof_open is not None:
            self._prof_open[4].record()
            self.profile.append(self._prof_open)
            self._prof_open = None

    # ------------------------------------------------------------------------------ plan
    def _add(self, layer):
        self.layers[layer.name] = layer
        return layer

    def _build_layers(self):
        enc = self.net.encoder
        self._add(ConvLayer('stem', enc.layer0[0].weight, 16, k=4, stride=1, pad=2, act=FP_ACT_RELU, bn=enc.layer0[1],
                            stem=True))
        self.enc_blocks = []
        cin = 64
        for li, seq in enumerate((enc.layer1[1], enc.layer2, enc.layer3, enc.layer4)):
            for bi, blk in enumerate(seq):
                p = f'enc{li + 1}.{bi}'
                cout = blk.conv1.weight.shape[0]
                c1 = self._add(ConvLayer(p + '.conv1', blk.conv1.weight, cin, k=3, stride=blk.stride, pad=1,
                                         act=FP_ACT_RELU, bn=blk.bn1))
                ds = None
                if blk.downsample is not None:
                    ds = self._add(ConvLayer(p + '.down', blk.downsample[0].weight, cin, k=1, stride=blk.stride, pad=0,
                                             act=FP_ACT_NONE, bn=blk.downsample[1]))
                c2 = self._add(ConvLayer(p + '.conv2', blk.conv2.weight, cout, k=3, stride=1, pad=1, act=FP_ACT_RELU,
                                         bn=blk.bn2))
                self.enc_blocks.append((li, c1, ds, c2))
                cin = cout
        self.dec = {}
        # FootprintNetwork: two decoders with 2-channel heads; Segmentor (segmentor.py): ONE decoder, 1-channel heads
        self.is_segmentor = hasattr(self.net, 'decoder')
        decoders = (('seg', self.net.decoder),) if self.is_segmentor else \
            (('mask', self.net.mask_decoder), ('depth', self.net.depth_decoder))
        for dname, dec in decoders:
            d = {}
            head_act = FP_ACT_SIGMOID if getattr(dec, 'apply_sigmoid', False) else FP_ACT_NONE
            # Segmentor(use_PSP=True): the pyramid-pooling block doubles the channels entering block1 (segmentation/network.py:62)
            first_in = 1024 if getattr(dec, 'use_PSP', False) else 512
            for bi, (ci, co) in enumerate(((first_in, 256), (256, 128), (128, 64), (64, 64)), start=1):
                blk = getattr(dec, f'block{bi}')
                pre, post = blk.pre_concat_conv, blk.post_concat_conv
                p = f'{dname}.block{bi}'
                d[f'b{bi}.pre1'] = self._add(ConvLayer(p + '.pre1', pre.conv1.weight, ci, pad_mode=FP_PAD_REFLECT,
                                                       act=FP_ACT_ELU, bias=pre.conv1.bias))
                d[f'b{bi}.pre2'] = self._add(ConvLayer(p + '.pre2', pre.conv2.weight, co, pad_mode=FP_PAD_REFLECT,
                                                       act=FP_ACT_ELU, bias=pre.conv2.bias))
                d[f'b{bi}.post1'] = self._add(ConvLayer(p + '.post1', post.conv1.weight, co, co, pad_mode=FP_PAD_REFLECT,
                                                        up0=1, act=FP_ACT_ELU, bias=post.conv1.bias))
                d[f'b{bi}.post2'] = self._add(ConvLayer(p + '.post2', post.conv2.weight, co, pad_mode=FP_PAD_REFLECT,
                                                        act=FP_ACT_ELU, bias=post.conv2.bias))
            for oi, ci in ((1, 128), (2, 64), (3, 64)):
                oc = getattr(dec, f'outconv{oi}')
                d[f'head{oi}'] = self._add(ConvLayer(f'{dname}.head{oi}', oc.conv1.weight, ci, pad_mode=FP_PAD_REFLECT,
                                                     act=head_act, bias=oc.conv1.bias, out_f32=True))
            oc4 = dec.outconv4
            d['o4.c1'] = self._add(ConvLayer(f'{dname}.o4.c1', oc4[0].conv1.weight, 64, pad_mode=FP_PAD_REFLECT, up0=1,
                                             act=FP_ACT_ELU, bias=oc4[0].conv1.bias))
            d['o4.c2'] = self._add(ConvLayer(f'{dname}.o4.c2', oc4[0].conv2.weight, 32, pad_mode=FP_PAD_REFLECT,
                                             act=FP_ACT_ELU, bias=oc4[0].conv2.bias))
            d['head4'] = self._add(ConvLayer(f'{dname}.head4', oc4[1].conv1.weight, 32, pad_mode=FP_PAD_REFLECT,
                                             act=head_act, bias=oc4[1].conv1.bias, out_f32=True))
            self.dec[dname] = d

    # ------------------------------------------------------------------------------ weights
    def _param_key(self):
        # `_weights_epoch` counts the updates the library made through raw pointers (fp_adam_step on the Trainer's flat
        # buffer, fp_bn_stats on the running statistics): those never bump a tensor's `_version`
        key = [getattr(self.net, '_weights_epoch', 0)]
        for p in self.net.parameters():
            key.append((p.data_ptr(), p._version))
        for b in self.net.buffers():
            key.append((b.data_ptr(), b._version))
        return tuple(key)

    def refresh_weights(self, force=False):
        """(Re)pack weights and fold eval-mode BatchNorm when any parameter/buffer changed."""
        key = self._param_key()
        if not force and key == self._weight_key:
            return
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        dev = self.layers['stem'].weight.device
        for L in self.layers.values():
            w = L.weight.detach()
            if w.dtype != torch.float32 or not w.is_contiguous():
                w = w.float().contiguous()
            if L.w_packed is None:
                L.w_packed = torch.empty((L.npad, L.kpad), dtype=self.tdtype, device=dev)
            if L.stem:
                _lib.check(self.lib.fp_pack_stem_weight(_ptr(w), L.cout, L.npad, self.fmt, _ptr(L.w_packed), stream))
            else:
                cin = L.cin0 + L.cin1
                _lib.check(self.lib.fp_pack_conv_weight(_ptr(w), L.cout, cin, L.k, L.k, cin, L.npad, L.kpad, self.fmt, 0,
                                                        _ptr(L.w_packed), stream))
            if self.precise and L.name.startswith(PRECISE_PREFIXES):
                # [Whi | Whi | Wlo] per tap (fp_conv_fwd_split)
                if getattr(L, 'w_packed3', None) is None:
                    L.kpad3 = 768 if L.stem else (3 * L.k * L.k * (L.cin0 + L.cin1) + 63) // 64 * 64
                    L.w_packed3 = torch.empty((L.npad, L.kpad3), dtype=self.tdtype, device=dev)
                if L.stem:
                    _lib.check(self.lib.fp_pack_stem_weight_split(_ptr(w), L.cout, L.npad, self.fmt, _ptr(L.w_packed3), stream))
                else:
                    _lib.check(self.lib.fp_pack_conv_weight_split(_ptr(w), L.cout, L.cin0 + L.cin1, L.k, L.k, L.npad, L.kpad3,
                                                                  self.fmt, _ptr(L.w_packed3), stream))
            if L.bn is not None:
                # eval-mode BN: y = (x - mean) / sqrt(var + eps) * gamma + beta  (torch BatchNorm2d)
                bn = L.bn
                scale = (bn.weight.detach().float() * torch.rsqrt(bn.running_var.float() + bn.eps))
                L.scale = scale.contiguous()
                L.shift = (bn.bias.detach().float() - bn.running_mean.float() * scale).contiguous()
            else:
                L.scale = None
                L.shift = L.bias.detach().float().contiguous() if L.bias is not None else None
        self._weight_key = key

    # ------------------------------------------------------------------------------ buffers
    def _buf(self, name, shape, dtype=None):
        dtype = dtype or self.tdtype
        t = self._bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.layers['stem'].weight.device)
            self._bufs[name] = t
        return t

    # ------------------------------------------------------------------------------ ops
    def _conv(self, L, src0, src1, out, residual, stream, scale=None, shift=None, act=None):
        b, h0, w0, c0 = src0.shape
        hin, win = (2 * h0, 2 * w0) if L.up0 else (h0, w0)
        d = FpConvDesc()
        d.batch, d.c0, d.h0, d.w0, d.up0 = b, c0, h0, w0, L.up0
        d.c1 = src1.shape[3] if src1 is not None else 0
        d.kh = d.kw = L.k
        d.stride, d.pad_t, d.pad_l, d.pad_mode = L.stride, L.pad, L.pad, L.pad_mode
        d.ho, d.wo = out.shape[1], out.shape[2]
        d.cout, d.kpad, d.npad, d.bn = L.cout, L.kpad, L.npad, L.bn_tile
        d.act = L.act if act is None else act
        d.dtype = self.fmt
        d.out_f32 = 1 if L.out_f32 else 0
        d.out_ld = out.shape[3]
        d.res_ld = residual.shape[3] if residual is not None else 0
        d.b_via_tma = 1
        assert c0 == L.cin0 and d.c1 == L.cin1, (L.name, c0, d.c1, L.cin0, L.cin1)
        if src1 is not None:
            assert tuple(src1.shape[1:3]) == (hin, win), (L.name, src1.shape, hin, win)
        self._prof('fwd', L, b * d.ho * d.wo,
                   2 * (src0.numel() + (src1.numel() if src1 is not None else 0) + b * d.ho * d.wo * L.cout) +
                   2 * L.cout * L.k * L.k * (c0 + d.c1))
        _lib.check(self.lib.fp_conv_fwd(ctypes.byref(d), _ptr(src0), _ptr(src1), _ptr(L.w_packed),
                                        _ptr(L.scale if scale is None else scale),
                                        _ptr(L.shift if shift is None else shift), _ptr(residual), _ptr(out), stream))
        self._prof_end()
        self.launches += 1

    def _conv_split(self, L, src0, src1, out, residual, stream):
        """Split-precision convolution (fp_conv_fwd_split): src0 / src1 / residual are (hi, lo) pairs or None; out is
        (hi, lo) -- lo None rounds the output once (fp32 head map or a plain 16-bit tensor)."""
        s0h, s0l = src0
        b, h0, w0, c0 = s0h.shape
        d = FpConvDesc()
        d.batch, d.c0, d.h0, d.w0, d.up0 = b, c0, h0, w0, L.up0
        d.c1 = src1[0].shape[3] if src1 is not None else 0
        d.kh = d.kw = L.k
        d.stride, d.pad_t, d.pad_l, d.pad_mode = L.stride, L.pad, L.pad, L.pad_mode
        oh, ol = out
        d.ho, d.wo = oh.shape[1], oh.shape[2]
        d.cout, d.kpad, d.npad, d.bn = L.cout, L.kpad3, L.npad, L.bn_tile
        d.act, d.dtype = L.act, self.fmt
        d.out_f32 = 1 if L.out_f32 else 0
        d.out_ld = oh.shape[3]
        d.res_ld = residual[0].shape[3] if residual is not None else 0
        d.b_via_tma = 1
        low = _lib.FpSplitLow(s0l.data_ptr(), src1[1].data_ptr() if src1 is not None else None,
                              residual[1].data_ptr() if residual is not None else None,
                              ol.data_ptr() if ol is not None else None)
        self._prof('fwd', L, b * d.ho * d.wo)
        _lib.check(self.lib.fp_conv_fwd_split(ctypes.byref(d), _ptr(s0h), _ptr(src1[0]) if src1 is not None else None,
                                              ctypes.byref(low), _ptr(L.w_packed3), _ptr(L.scale), _ptr(L.shift),
                                              _ptr(residual[0]) if residual is not None else None, _ptr(oh), stream))
        self._prof_end()
        self.launches += 1

    def _pair(self, name, shape):
        return self._buf(name + '.hi', shape), self._buf(name + '.lo', shape)

    def _head(self, L, src, out, stream):
        """OutConvBlock conv (+ sigmoid): footprints/network.py:174-178.  Default: the row-streaming tensor-core kernel
        with the two output channels padded to N = 16 (the tensor pipe is idle on these HBM-bound layers; one
        N = 48 instruction per input row, dx and K step); ``head_tensor_cores = False`` selects the CUDA-core kernels."""
        if self.head_tensor_cores or L.cout != 2:
            return self._conv(L, src, None, out, None, stream)
        b, h, w, c = src.shape
        _lib.check(self.lib.fp_head_conv_fwd(_ptr(src), _ptr(L.weight.detach()), _ptr(L.bias.detach()), b, h, w, c,
                                             1 if L.act == FP_ACT_SIGMOID else 0, self.fmt, _ptr(out), stream))
        self.launches += 1

    # ------------------------------------------------------------------------------ forward (inference)
    @torch.no_grad()
    def forward(self, image, post=None):
        """Eval-mode forward (folded BatchNorm).  Matches FootprintNetwork.forward of the reference in
        ``model.eval()`` (footprints/network.py:21-30)."""
        if self.net.training:
            raise NotImplementedError('training-mode forward (batch-statistics BatchNorm) goes through '
                                      'footprints_b200.training; call model.eval() for inference')
        _lib.require_device()
        if image.dim() != 4 or image.shape[1] != 3:
            raise RuntimeError(f'expected input [B,3,H,W], got {tuple(image.shape)}')
        b, _, h, w = image.shape
        if h % 32 or w % 32:
            # the reference fails in torch.cat for such sizes
            raise RuntimeError(f'input height and width must be multiples of 32, got {h}x{w}')
        image = image.contiguous().float()
        self.refresh_weights()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib = self.lib

        if self.precise:
            if self.is_segmentor:
                raise NotImplementedError('precise mode is implemented for FootprintNetwork')
            return self._forward_precise(image, stream, post)
        s2d = self._buf('s2d', (b, h // 2, w // 2, 16))
        _lib.check(lib.fp_stem_prep(_ptr(image), b, h, w, self.fmt, _ptr(s2d), stream))
        f0 = self._buf('f0', (b, h // 2, w // 2, 64))
        self._conv(self.layers['stem'], s2d, None, f0, None, stream)
        x = self._buf('pool', (b, h // 4, w // 4, 64))
        _lib.check(lib.fp_maxpool3x3s2(_ptr(f0), b, h // 2, w // 2, 64, self.fmt, _ptr(x), stream))
        feats = [f0]
        last_li = 0
        for idx, (li, c1, ds, c2) in enumerate(self.enc_blocks):
            if li != last_li:
                feats.append(x)
                last_li = li
            hh, ww = x.shape[1] // c1.stride, x.shape[2] // c1.stride
            t1 = self._buf(f'{c1.name}.out', (b, hh, ww, c1.cout))
            self._conv(c1, x, None, t1, None, stream)
            idt = x
            if ds is not None:
                idt = self._buf(f'{ds.name}.out', (b, hh, ww, ds.cout))
                self._conv(ds, x, None, idt, None, stream)
            out = self._buf(f'{c2.name}.out', (b, hh, ww, c2.cout))
            self._conv(c2, t1, None, out, idt, stream)
            x = out
        feats.append(x)

        heads = {}
        for dname in self.dec:
            d = self.dec[dname]
            x = feats[4]
            if self.is_segmentor and self.net.decoder.use_PSP:
                # segmentation/network.py:82-83,174-207: pool / reduce / upsample / concat of all four levels in one kernel
                psp = self.net.decoder.PSP
                wts = [getattr(psp, f'block{i}').reduce.weight.detach().float().contiguous() for i in (1, 2, 3, 4)]
                xp = self._buf('seg.psp', (b, x.shape[1], x.shape[2], 1024))
                _lib.check(lib.fp_psp_fwd(_ptr(x), b, x.shape[1], x.shape[2], self.fmt, _ptr(wts[0]), _ptr(wts[1]), _ptr(wts[2]),
                                          _ptr(wts[3]), _ptr(xp), stream))
                x = xp
            for bi in range(1, 5):
                skip = feats[4 - bi]
                hh, ww = x.shape[1], x.shape[2]
                co = d[f'b{bi}.pre1'].cout
                t = self._buf(f'{dname}.b{bi}.pre1', (b, hh, ww, co))
                self._conv(d[f'b{bi}.pre1'], x, None, t, None, stream)
                t2 = self._buf(f'{dname}.b{bi}.pre2', (b, hh, ww, co))
                self._conv(d[f'b{bi}.pre2'], t, None, t2, None, stream)
                t3 = self._buf(f'{dname}.b{bi}.post1', (b, 2 * hh, 2 * ww, co))
                self._conv(d[f'b{bi}.post1'], t2, skip, t3, None, stream)
                x = self._buf(f'{dname}.b{bi}.post2', (b, 2 * hh, 2 * ww, co))
                self._conv(d[f'b{bi}.post2'], t3, None, x, None, stream)
                if bi >= 2:
                    hd = self._buf(f'{dname}.head{bi - 1}', (b, 2 * hh, 2 * ww, d[f'head{bi - 1}'].cout), torch.float32)
                    self._head(d[f'head{bi - 1}'], x, hd, stream)
                    heads[(dname, bi - 1)] = hd
            hh, ww = x.shape[1], x.shape[2]
            t = self._buf(f'{dname}.o4.c1', (b, 2 * hh, 2 * ww, 32))
            self._conv(d['o4.c1'], x, None, t, None, stream)
            t2 = self._buf(f'{dname}.o4.c2', (b, 2 * hh, 2 * ww, 32))
            self._conv(d['o4.c2'], t, None, t2, None, stream)
            hd = self._buf(f'{dname}.head4', (b, 2 * hh, 2 * ww, d['head4'].cout), torch.float32)
            self._head(d['head4'], t2, hd, stream)
            heads[(dname, 4)] = hd

        if post == 'inference_f16':
            # evaluation/inference.py:104-108 + inference_dataset.py:35-39 fused into the output kernel: only the full-resolution
            # prediction, sigmoid on the mask channels, float16 -- half the D2H bytes of the fp32 '1/1' tensor, no three extra scales
            out = torch.empty((b, 4, h, w), dtype=torch.float16, device=image.device)
            _lib.check(lib.fp_heads_to_nchw_f16_post(_ptr(heads[('mask', 4)]), _ptr(heads[('depth', 4)]), b, h, w, _ptr(out), stream))
            return out
        if self.is_segmentor:
            # preprocessing/segmentation/network.py:83-101: a LIST of the four 1-channel maps at their own resolution
            # ([B,h,w,1] NHWC is [B,1,h,w] NCHW); fresh tensors, the head buffers are reused by the next call
            return [heads[('seg', i)].view(b, 1, heads[('seg', i)].shape[1], heads[('seg', i)].shape[2]).clone() for i in (1, 2, 3, 4)]
        return self._assemble(heads, b, h, w, image.device, stream)

    def _assemble(self, heads, b, h, w, dev, stream):
        outputs = OrderedDict()
        for hi, key in enumerate(SCALES, start=1):
            scale = (8, 4, 2, 1)[hi - 1]
            # fresh tensors: callers mutate the dict / write into the outputs in place
            out = torch.empty((b, 4, h, w), dtype=torch.float32, device=dev)
            mh = heads[('mask', hi)]
            _lib.check(self.lib.fp_heads_to_nchw(_ptr(mh), _ptr(heads[('depth', hi)]), b, mh.shape[1], mh.shape[2], scale,
                                                 _ptr(out), stream))
            outputs[key] = out
        return outputs

    def _forward_precise(self, image, stream, post=None):
        """Eval forward with the PRECISE_PREFIXES layers in split precision; everything else as in forward()."""
        lib = self.lib
        b, _, h, w = image.shape
        s2d = self._pair('s2d', (b, h // 2, w // 2, 16))
        _lib.check(lib.fp_stem_prep_split(_ptr(image), b, h, w, self.fmt, _ptr(s2d[0]), _ptr(s2d[1]), stream))
        f0 = self._pair('f0', (b, h // 2, w // 2, 64))
        self._conv_split(self.layers['stem'], s2d, None, f0, None, stream)
        x = self._pair('pool', (b, h // 4, w // 4, 64))
        _lib.check(lib.fp_maxpool3x3s2_split(_ptr(f0[0]), _ptr(f0[1]), b, h // 2, w // 2, 64, self.fmt, _ptr(x[0]), _ptr(x[1]),
                                             stream))
        feats = [f0]
        last_li = 0
        for li, c1, ds, c2 in self.enc_blocks:
            if li != last_li:
                feats.append(x)
                last_li = li
            hh, ww = x[0].shape[1] // c1.stride, x[0].shape[2] // c1.stride
            t1 = self._pair(f'{c1.name}.out', (b, hh, ww, c1.cout))
            self._conv_split(c1, x, None, t1, None, stream)
            idt = x
            if ds is not None:
                idt = self._pair(f'{ds.name}.out', (b, hh, ww, ds.cout))
                self._conv_split(ds, x, None, idt, None, stream)
            out = self._pair(f'{c2.name}.out', (b, hh, ww, c2.cout))
            self._conv_split(c2, t1, None, out, idt, stream)
            x = out
        feats.append(x)

        heads = {}
        for dname in ('mask', 'depth'):
            d = self.dec[dname]
            x = feats[4] if dname == 'mask' else feats[4][0]
            for bi in range(1, 5):
                split = dname == 'mask' and bi <= 2
                skip = feats[4 - bi]
                xs = x[0] if isinstance(x, tuple) else x
                hh, ww = xs.shape[1], xs.shape[2]
                co = d[f'b{bi}.pre1'].cout
                if split:
                    t = self._pair(f'{dname}.b{bi}.pre1', (b, hh, ww, co))
                    self._conv_split(d[f'b{bi}.pre1'], x, None, t, None, stream)
                    t2 = self._pair(f'{dname}.b{bi}.pre2', (b, hh, ww, co))
                    self._conv_split(d[f'b{bi}.pre2'], t, None, t2, None, stream)
                    t3 = self._pair(f'{dname}.b{bi}.post1', (b, 2 * hh, 2 * ww, co))
                    self._conv_split(d[f'b{bi}.post1'], t2, skip, t3, None, stream)
                    x = self._pair(f'{dname}.b{bi}.post2', (b, 2 * hh, 2 * ww, co))
                    self._conv_split(d[f'b{bi}.post2'], t3, None, x, None, stream)
                    if bi == 2:
                        hd = self._buf(f'{dname}.head1', (b, 2 * hh, 2 * ww, 2), torch.float32)
                        self._conv_split(d['head1'], x, None, (hd, None), None, stream)
                        heads[(dname, 1)] = hd
                    continue
                xs = x[0] if isinstance(x, tuple) else x          # leaving the precise set: the high halves are the tensor
                t = self._buf(f'{dname}.b{bi}.pre1', (b, hh, ww, co))
                self._conv(d[f'b{bi}.pre1'], xs, None, t, None, stream)
                t2 = self._buf(f'{dname}.b{bi}.pre2', (b, hh, ww, co))
                self._conv(d[f'b{bi}.pre2'], t, None, t2, None, stream)
                t3 = self._buf(f'{dname}.b{bi}.post1', (b, 2 * hh, 2 * ww, co))
                self._conv(d[f'b{bi}.post1'], t2, skip[0], t3, None, stream)
                x = self._buf(f'{dname}.b{bi}.post2', (b, 2 * hh, 2 * ww, co))
                self._conv(d[f'b{bi}.post2'], t3, None, x, None, stream)
                if bi >= 2:
                    hd = self._buf(f'{dname}.head{bi - 1}', (b, 2 * hh, 2 * ww, 2), torch.float32)
                    self._head(d[f'head{bi - 1}'], x, hd, stream)
                    heads[(dname, bi - 1)] = hd
            hh, ww = x.shape[1], x.shape[2]
            t = self._buf(f'{dname}.o4.c1', (b, 2 * hh, 2 * ww, 32))
            self._conv(d['o4.c1'], x, None, t, None, stream)
            t2 = self._buf(f'{dname}.o4.c2', (b, 2 * hh, 2 * ww, 32))
            self._conv(d['o4.c2'], t, None, t2, None, stream)
            hd = self._buf(f'{dname}.head4', (b, 2 * hh, 2 * ww, 2), torch.float32)
            self._head(d['head4'], t2, hd, stream)
            heads[(dname, 4)] = hd
        if post == 'inference_f16':
            out = torch.empty((b, 4, h, w), dtype=torch.float16, device=image.device)
            _lib.check(lib.fp_heads_to_nchw_f16_post(_ptr(heads[('mask', 4)]), _ptr(heads[('depth', 4)]), b, h, w, _ptr(out), stream))
            return out
        return self._assemble(heads, b, h, w, image.device, stream)
