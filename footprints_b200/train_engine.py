"""Training step of the Footprints hot path on one B200: train-mode forward (batch-statistics
BatchNorm), and the hand-written backward that replaces autograd over footprints/network.py
(reference: footprints/training/train.py:153-156,224-225).

Every arithmetic step is a C-ABI call into ``libfootprints_b200.so``:

  forward   conv (tcgen05 implicit GEMM) -> fp_bn_stats -> fp_bn_apply      (encoder)
            conv + bias + ELU fused                                          (decoders, as in eval)
  backward  fp_heads_bwd -> per conv: fp_conv_wgrad, fp_bias_grad,
            fp_conv_fwd with the flipped weights as a zero-padded full correlation (data gradient of
            the padded virtual input) -> fp_fold_bwd (adjoint of reflect-pad / nearest x2 / concat,
            fused with ELU')                                                 (decoders)
            fp_bn_bwd -> fp_conv_wgrad -> fp_conv_fwd(flipped weights [, zero-inserted x2 for the
            stride-2 layers]) with the residual-branch gradient added in the epilogue  (encoder)

Storage types: activations and inter-layer gradients are ONE 16-bit type (``tcgen05.mma kind::f16`` rejects mixed
f16 x bf16 operands and the weight gradient multiplies the two).  Default ``fp16``: 11-bit significands keep the
train-mode forward as close to the fp32 reference as the eval path, and a power-of-two LOSS SCALE applied where the
backward starts (``fp_heads_bwd``) keeps the gradients inside fp16's range -- over the whole network |dZ| spans
2^-17 .. 2^5 x u, u = 1/(4*B*H*W), on the synthetic batch, i.e. 2^-12 .. 2^10 after the default scale 32/u: normal fp16
numbers with a 2^6 margin to the top (65504); every 16-bit conversion SATURATES (``cvt.rn.satfinite``), so a gradient
spike clips element-wise instead of turning into inf - inf = NaN downstream.  Every backward kernel is linear in that
tensor and the fp32 parameter gradients are un-scaled exactly (power of two) by the optimiser / the autograd glue.  ``bf16`` (scale 1) remains selectable.  Parameter gradients are fp32.
"""
import ctypes
import os
from collections import OrderedDict

import torch

from . import _lib
import math

from ._lib import FP_ACT_ELU, FP_ACT_NONE, FP_ACT_RELU, FP_BF16, FP_F16, FP_PAD_ZERO, FpConvDesc
from .engine import PRECISE_PREFIXES, SCALES, Engine, _bn_tile, _ptr

_RAW = object()   # sentinel: run the conv with no scale/shift (NULL pointers)


class TrainEngine(Engine):

    def __init__(self, net, compute_dtype='fp16', precise=False):
        super().__init__(net, compute_dtype, precise)
        self.gfmt = self.fmt                 # gradients share the activations' 16-bit type (see the module docstring)
        self.gdtype = self.tdtype
        self.loss_scale = None               # None: 2^k ~ 32 * 4*B*H*W for fp16 gradients, 1 for bf16; or a fixed float
        self.grad_scale = 1.0                # scale of the parameter gradients the last backward() produced
        self._red_ws = None
        self._wg_ws = None
        self._loss_ws = None
        self.head_tensor_cores_wgrad = True      # measured: the tiled CUDA-core head wgrad (0.64 ms for head4) is no faster than the tensor-core one (0.55)
        self.head_tensor_cores_dgrad = False     # measured: full correlation on the general kernel (K = 9 x 8) + fold: 0.71 + 0.25 ms for head4 vs 0.9 ms CUDA-core
        self._head_ws = None
        # weight / bias gradients run on a second stream beside the data-gradient chain that is the backward's critical
        # path (both only consume dZ): the many small latency-bound kernels (split-K reduces, bias column sums, head
        # weight gradients) and the tails of the persistent convolution kernels back-fill each other
        self.overlap_wgrad = os.environ.get('FP_OVERLAP_WGRAD', '1') != '0'
        self._side = None
        self._side_keep = []
        self._red_ws_side = None
        self.grad_views = None      # optional {parameter: fp32 view} — gradients are written there (Trainer's flat buffer)

    def _gout(self, param, shape=None):
        """fp32 destination of a parameter gradient: the Trainer's flat-buffer view when present."""
        if self.grad_views is not None and param in self.grad_views:
            return self.grad_views[param]
        return torch.empty(param.shape if shape is None else shape, dtype=torch.float32, device=param.device)

    # ------------------------------------------------------------------------------ helpers
    def _dev(self):
        return self.layers['stem'].weight.device

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _reduce_ws(self):
        if self._red_ws is None:
            n = self.lib.fp_reduce_workspace_floats(512)
            self._red_ws = torch.empty(n, dtype=torch.float32, device=self._dev())
        return self._red_ws

    def _desc(self, L, src0, src1, ho, wo, dtype=None):
        b, h0, w0, c0 = src0.shape
        d = FpConvDesc()
        d.batch, d.c0, d.h0, d.w0, d.up0 = b, c0, h0, w0, L.up0
        d.c1 = src1.shape[3] if src1 is not None else 0
        d.kh = d.kw = L.k
        d.stride, d.pad_t, d.pad_l, d.pad_mode = L.stride, L.pad, L.pad, L.pad_mode
        d.ho, d.wo = ho, wo
        d.cout, d.kpad, d.npad, d.bn = L.cout, L.kpad, L.npad, L.bn_tile
        d.act = L.act
        d.dtype = self.fmt if dtype is None else dtype
        d.b_via_tma = 1
        return d

    def _pack_plan(self):
        """Device table of FpPackJob records: forward and rotated (data-gradient) packing of every convolution weight.
        Rebuilt only when a parameter moved (the Trainer re-points them into its flat buffer once)."""
        convs = list(self.layers.values())        # heads too: their forward runs on the tensor-core kernel (N padded to 16)
        key = tuple(L.weight.data_ptr() for L in convs)
        if getattr(self, '_pack_key', None) == key:
            return
        dev = self._dev()
        jobs, first = [], 0
        for L in convs:
            w = L.weight.detach()
            cin = L.cin0 + L.cin1
            if L.w_packed is None:
                L.w_packed = torch.empty((L.npad, L.kpad), dtype=self.tdtype, device=dev)
            if L.stem:
                jobs.append(_lib.FpPackJob(w.data_ptr(), L.w_packed.data_ptr(), L.cout, 3, 7, 7, 16, L.npad, 256, self.fmt, 0, 1, first))
                first += L.npad * 256
                if self.precise:
                    if getattr(L, 'w_packed3', None) is None:
                        L.kpad3 = 768
                        L.w_packed3 = torch.empty((L.npad, 768), dtype=self.tdtype, device=dev)
                    jobs.append(_lib.FpPackJob(w.data_ptr(), L.w_packed3.data_ptr(), L.cout, 3, 7, 7, 16, L.npad, 768, self.fmt, 2, 1, first))
                    first += L.npad * 768
                continue
            jobs.append(_lib.FpPackJob(w.data_ptr(), L.w_packed.data_ptr(), L.cout, cin, L.k, L.k, cin, L.npad, L.kpad,
                                       self.fmt, 0, 0, first))
            first += L.npad * L.kpad
            if self.precise and L.name.startswith(PRECISE_PREFIXES):       # [Whi | Whi | Wlo] per tap (flip = 2)
                if getattr(L, 'w_packed3', None) is None:
                    L.kpad3 = 768 if L.stem else (3 * L.k * L.k * cin + 63) // 64 * 64
                    L.w_packed3 = torch.empty((L.npad, L.kpad3), dtype=self.tdtype, device=dev)
                jobs.append(_lib.FpPackJob(w.data_ptr(), L.w_packed3.data_ptr(), L.cout, 3 if L.stem else cin,
                                           7 if L.stem else L.k, 7 if L.stem else L.k, 16 if L.stem else cin, L.npad,
                                           L.kpad3, self.fmt, 2, 1 if L.stem else 0, first))
                first += L.npad * L.kpad3
            # data-gradient operand (heads too: their data gradient is a K = 9 x 8 full correlation on the general tensor-core
            # kernel followed by the usual fold -- half the time of the CUDA-core head kernel)
            # data-gradient operand: rows = cin, K = taps * cout_pad, filter rotated by 180 degrees
            if getattr(L, 'w_flip', None) is None:
                L.cpad_out = max(8, L.cout)
                L.d_bn = _bn_tile(cin)
                L.d_npad = (cin + L.d_bn - 1) // L.d_bn * L.d_bn
                L.d_kpad = (L.k * L.k * L.cpad_out + 63) // 64 * 64
                L.w_flip = torch.empty((L.d_npad, L.d_kpad), dtype=self.gdtype, device=dev)
            jobs.append(_lib.FpPackJob(w.data_ptr(), L.w_flip.data_ptr(), L.cout, cin, L.k, L.k, L.cpad_out, L.d_npad,
                                       L.d_kpad, self.gfmt, 1, 0, first))
            first += L.d_npad * L.d_kpad
        arr = (_lib.FpPackJob * len(jobs))(*jobs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._pack_table = host.to(dev)
        self._pack_n, self._pack_total, self._pack_key = len(jobs), first, key

    def refresh_weights(self, force=False):
        """Training: weights change every step, so every operand is re-packed -- ONE launch for all of them."""
        self._pack_plan()
        _lib.check(self.lib.fp_pack_weights_batched(_ptr(self._pack_table), self._pack_n, self._pack_total, self._stream()))
        self.launches += 1
        for L in self.layers.values():
            L.scale = None
            L.shift = L.bias.detach() if L.bias is not None else None
        self._weight_key = None

    # ------------------------------------------------------------------------------ forward
    def _conv_bn(self, L, src, residual, relu, stream, saved):
        """conv -> batch-stat BatchNorm (-> + residual) (-> ReLU); saves z and y for the backward."""
        b, h, w, _ = src.shape
        ho, wo = (h // L.stride, w // L.stride) if not L.stem else (h, w)
        z = torch.empty((b, ho, wo, L.cout), dtype=self.tdtype, device=src.device)
        self._conv(L, src, None, z, None, stream, scale=_RAW, shift=_RAW, act=FP_ACT_NONE)
        # batch statistics belong to THIS forward (a second forward before the backward must not overwrite them)
        stv = torch.empty((4, L.cout), dtype=torch.float32, device=src.device)
        st = {'scale': stv[0], 'shift': stv[1], 'mean': stv[2], 'invstd': stv[3]}
        bn = L.bn
        M = b * ho * wo
        _lib.check(self.lib.fp_bn_stats(_ptr(z), M, L.cout, self.fmt, _ptr(bn.weight.detach()), _ptr(bn.bias.detach()),
                                        _ptr(bn.running_mean), _ptr(bn.running_var), float(bn.momentum), float(bn.eps),
                                        _ptr(st['scale']), _ptr(st['shift']), _ptr(st['mean']), _ptr(st['invstd']),
                                        _ptr(self._reduce_ws()), stream))
        y = torch.empty_like(z)
        _lib.check(self.lib.fp_bn_apply(_ptr(z), _ptr(st['scale']), _ptr(st['shift']), _ptr(residual), 1 if relu else 0, M,
                                        L.cout, self.fmt, _ptr(y), stream))
        saved[L.name] = (src, z, y)
        saved['bn'][L.name] = st
        return y

    def _conv(self, L, src0, src1, out, residual, stream, scale=None, shift=None, act=None):
        if scale is _RAW:
            # temporarily present NULL scale/shift
            s0, s1 = L.scale, L.shift
            L.scale, L.shift = None, None
            try:
                super()._conv(L, src0, src1, out, residual, stream, act=act)
            finally:
                L.scale, L.shift = s0, s1
        else:
            super()._conv(L, src0, src1, out, residual, stream, scale, shift, act)

    def forward_train(self, image, save=True):
        """Train-mode forward (footprints/network.py:21-30 with BatchNorm in training mode).  Returns
        ``(outputs, saved)``: the dict of four fp32 NCHW outputs and everything the backward of THIS forward needs
        (activations, per-layer batch statistics) -- the caller keeps ``saved`` (the autograd node's ctx, or the
        Trainer) and hands it to ``backward``; ``save=False`` (no gradient wanted) keeps nothing alive."""
        _lib.require_device()
        b, _, h, w = image.shape
        if h % 32 or w % 32:
            raise RuntimeError(f'input height and width must be multiples of 32, got {h}x{w}')
        image = image.contiguous().float()
        self.refresh_weights()
        stream = self._stream()
        lib = self.lib
        dev = image.device
        sv = {'bn': {}}

        def new(shape, dtype=None):
            return torch.empty(shape, dtype=dtype or self.tdtype, device=dev)

        if self.precise:
            return self._forward_train_precise(image, save, stream, sv)
        s2d = new((b, h // 2, w // 2, 16))
        _lib.check(lib.fp_stem_prep(_ptr(image), b, h, w, self.fmt, _ptr(s2d), stream))
        f0 = self._conv_bn(self.layers['stem'], s2d, None, True, stream, sv)
        x = new((b, h // 4, w // 4, 64))
        _lib.check(lib.fp_maxpool3x3s2(_ptr(f0), b, h // 2, w // 2, 64, self.fmt, _ptr(x), stream))
        sv['pool'] = x
        feats = [f0]
        last_li = 0
        for li, c1, ds, c2 in self.enc_blocks:
            if li != last_li:
                feats.append(x)
                last_li = li
            t1 = self._conv_bn(c1, x, None, True, stream, sv)
            idt = x
            if ds is not None:
                idt = self._conv_bn(ds, x, None, False, stream, sv)
            x = self._conv_bn(c2, t1, idt, True, stream, sv)
        feats.append(x)
        sv['feats'] = feats

        heads = {}
        for dname in ('mask', 'depth'):
            d = self.dec[dname]
            x = feats[4]
            for bi in range(1, 5):
                skip = feats[4 - bi]
                hh, ww = x.shape[1], x.shape[2]
                co = d[f'b{bi}.pre1'].cout
                t = new((b, hh, ww, co))
                self._conv(d[f'b{bi}.pre1'], x, None, t, None, stream)
                sv[d[f'b{bi}.pre1'].name] = (x, None, t)
                t2 = new((b, hh, ww, co))
                self._conv(d[f'b{bi}.pre2'], t, None, t2, None, stream)
                sv[d[f'b{bi}.pre2'].name] = (t, None, t2)
                t3 = new((b, 2 * hh, 2 * ww, co))
                self._conv(d[f'b{bi}.post1'], t2, skip, t3, None, stream)
                sv[d[f'b{bi}.post1'].name] = (t2, skip, t3)
                x = new((b, 2 * hh, 2 * ww, co))
                self._conv(d[f'b{bi}.post2'], t3, None, x, None, stream)
                sv[d[f'b{bi}.post2'].name] = (t3, None, x)
                if bi >= 2:
                    hd = new((b, 2 * hh, 2 * ww, 2), torch.float32)
                    self._head(d[f'head{bi - 1}'], x, hd, stream)
                    sv[d[f'head{bi - 1}'].name] = (x, None, hd)
                    heads[(dname, bi - 1)] = hd
            hh, ww = x.shape[1], x.shape[2]
            t = new((b, 2 * hh, 2 * ww, 32))
            self._conv(d['o4.c1'], x, None, t, None, stream)
            sv[d['o4.c1'].name] = (x, None, t)
            t2 = new((b, 2 * hh, 2 * ww, 32))
            self._conv(d['o4.c2'], t, None, t2, None, stream)
            sv[d['o4.c2'].name] = (t, None, t2)
            hd = new((b, 2 * hh, 2 * ww, 2), torch.float32)
            self._head(d['head4'], t2, hd, stream)
            sv[d['head4'].name] = (t2, None, hd)
            heads[(dname, 4)] = hd

        outputs = OrderedDict()
        for hi, key in enumerate(SCALES, start=1):
            scale = (8, 4, 2, 1)[hi - 1]
            out = torch.empty((b, 4, h, w), dtype=torch.float32, device=dev)
            mh = heads[('mask', hi)]
            _lib.check(lib.fp_heads_to_nchw(_ptr(mh), _ptr(heads[('depth', hi)]), b, mh.shape[1], mh.shape[2], scale,
                                            _ptr(out), stream))
            outputs[key] = out
        sv['s2d'] = s2d
        sv['shape'] = (b, h, w)
        # torchvision BatchNorm2d bookkeeping (num_batches_tracked += 1 per BN per forward): one multi-tensor launch
        torch._foreach_add_([L.bn.num_batches_tracked for L in self.layers.values() if L.bn is not None], 1)
        self.net._weights_epoch = getattr(self.net, '_weights_epoch', 0) + 1      # running statistics changed (raw pointers)
        return outputs, (sv if save else None)

    # ------------------------------------------------------------------------------ precise (split-precision) forward
    def _conv_bn_split(self, L, src, residual, relu, stream, saved):
        """_conv_bn on hi + lo pairs: split-precision convolution, batch statistics over hi + lo, split BatchNorm apply.
        The backward reads the HIGH halves only (they are what the fast path stores), so they are what is saved."""
        b, h, w, _ = src[0].shape
        ho, wo = (h // L.stride, w // L.stride) if not L.stem else (h, w)
        dev = src[0].device
        z = tuple(torch.empty((b, ho, wo, L.cout), dtype=self.tdtype, device=dev) for _ in range(2))
        s0, s1, a0 = L.scale, L.shift, L.act
        L.scale, L.shift, L.act = None, None, FP_ACT_NONE
        try:
            self._conv_split(L, src, None, z, None, stream)
        finally:
            L.scale, L.shift, L.act = s0, s1, a0
        stv = torch.empty((4, L.cout), dtype=torch.float32, device=dev)
        st = {'scale': stv[0], 'shift': stv[1], 'mean': stv[2], 'invstd': stv[3]}
        bn = L.bn
        M = b * ho * wo
        _lib.check(self.lib.fp_bn_stats_split(_ptr(z[0]), _ptr(z[1]), M, L.cout, self.fmt, _ptr(bn.weight.detach()),
                                              _ptr(bn.bias.detach()), _ptr(bn.running_mean), _ptr(bn.running_var),
                                              float(bn.momentum), float(bn.eps), _ptr(st['scale']), _ptr(st['shift']),
                                              _ptr(st['mean']), _ptr(st['invstd']), _ptr(self._reduce_ws()), stream))
        y = (torch.empty_like(z[0]), torch.empty_like(z[0]))
        _lib.check(self.lib.fp_bn_apply_split(_ptr(z[0]), _ptr(z[1]), _ptr(st['scale']), _ptr(st['shift']),
                                              _ptr(residual[0]) if residual is not None else None,
                                              _ptr(residual[1]) if residual is not None else None, 1 if relu else 0, M,
                                              L.cout, self.fmt, _ptr(y[0]), _ptr(y[1]), stream))
        self.launches += 3
        saved[L.name] = (src[0], z[0], y[0])
        saved['bn'][L.name] = st
        return y

    def _forward_train_precise(self, image, save, stream, sv):
        """forward_train with the PRECISE_PREFIXES layers in split precision (same saved-state layout: the backward is
        unchanged and reads the high halves)."""
        lib = self.lib
        b, _, h, w = image.shape
        dev = image.device

        def new(shape, dtype=None):
            return torch.empty(shape, dtype=dtype or self.tdtype, device=dev)

        def pair(shape):
            return new(shape), new(shape)

        s2d = pair((b, h // 2, w // 2, 16))
        _lib.check(lib.fp_stem_prep_split(_ptr(image), b, h, w, self.fmt, _ptr(s2d[0]), _ptr(s2d[1]), stream))
        f0 = self._conv_bn_split(self.layers['stem'], s2d, None, True, stream, sv)
        x = pair((b, h // 4, w // 4, 64))
        _lib.check(lib.fp_maxpool3x3s2_split(_ptr(f0[0]), _ptr(f0[1]), b, h // 2, w // 2, 64, self.fmt, _ptr(x[0]), _ptr(x[1]),
                                             stream))
        sv['pool'] = x[0]
        feats = [f0]
        last_li = 0
        for li, c1, ds, c2 in self.enc_blocks:
            if li != last_li:
                feats.append(x)
                last_li = li
            t1 = self._conv_bn_split(c1, x, None, True, stream, sv)
            idt = x
            if ds is not None:
                idt = self._conv_bn_split(ds, x, None, False, stream, sv)
            x = self._conv_bn_split(c2, t1, idt, True, stream, sv)
        feats.append(x)
        sv['feats'] = [f[0] for f in feats]

        heads = {}
        for dname in ('mask', 'depth'):
            d = self.dec[dname]
            x = feats[4] if dname == 'mask' else feats[4][0]
            for bi in range(1, 5):
                skip = feats[4 - bi]
                if dname == 'mask' and bi <= 2:
                    hh, ww = x[0].shape[1], x[0].shape[2]
                    co = d[f'b{bi}.pre1'].cout
                    t = pair((b, hh, ww, co))
                    self._conv_split(d[f'b{bi}.pre1'], x, None, t, None, stream)
                    sv[d[f'b{bi}.pre1'].name] = (x[0], None, t[0])
                    t2 = pair((b, hh, ww, co))
                    self._conv_split(d[f'b{bi}.pre2'], t, None, t2, None, stream)
                    sv[d[f'b{bi}.pre2'].name] = (t[0], None, t2[0])
                    t3 = pair((b, 2 * hh, 2 * ww, co))
                    self._conv_split(d[f'b{bi}.post1'], t2, skip, t3, None, stream)
                    sv[d[f'b{bi}.post1'].name] = (t2[0], skip[0], t3[0])
                    x = pair((b, 2 * hh, 2 * ww, co))
                    self._conv_split(d[f'b{bi}.post2'], t3, None, x, None, stream)
                    sv[d[f'b{bi}.post2'].name] = (t3[0], None, x[0])
                    if bi == 2:
                        hd = new((b, 2 * hh, 2 * ww, 2), torch.float32)
                        self._conv_split(d['head1'], x, None, (hd, None), None, stream)
                        sv[d['head1'].name] = (x[0], None, hd)
                        heads[(dname, 1)] = hd
                    continue
                if isinstance(x, tuple):
                    x = x[0]                  # leaving the precise set: the high halves are the tensor
                hh, ww = x.shape[1], x.shape[2]
                co = d[f'b{bi}.pre1'].cout
                t = new((b, hh, ww, co))
                self._conv(d[f'b{bi}.pre1'], x, None, t, None, stream)
                sv[d[f'b{bi}.pre1'].name] = (x, None, t)
                t2 = new((b, hh, ww, co))
                self._conv(d[f'b{bi}.pre2'], t, None, t2, None, stream)
                sv[d[f'b{bi}.pre2'].name] = (t, None, t2)
                t3 = new((b, 2 * hh, 2 * ww, co))
                self._conv(d[f'b{bi}.post1'], t2, skip[0], t3, None, stream)
                sv[d[f'b{bi}.post1'].name] = (t2, skip[0], t3)
                x = new((b, 2 * hh, 2 * ww, co))
                self._conv(d[f'b{bi}.post2'], t3, None, x, None, stream)
                sv[d[f'b{bi}.post2'].name] = (t3, None, x)
                if bi >= 2:
                    hd = new((b, 2 * hh, 2 * ww, 2), torch.float32)
                    self._head(d[f'head{bi - 1}'], x, hd, stream)
                    sv[d[f'head{bi - 1}'].name] = (x, None, hd)
                    heads[(dname, bi - 1)] = hd
            hh, ww = x.shape[1], x.shape[2]
            t = new((b, 2 * hh, 2 * ww, 32))
            self._conv(d['o4.c1'], x, None, t, None, stream)
            sv[d['o4.c1'].name] = (x, None, t)
            t2 = new((b, 2 * hh, 2 * ww, 32))
            self._conv(d['o4.c2'], t, None, t2, None, stream)
            sv[d['o4.c2'].name] = (t, None, t2)
            hd = new((b, 2 * hh, 2 * ww, 2), torch.float32)
            self._head(d['head4'], t2, hd, stream)
            sv[d['head4'].name] = (t2, None, hd)
            heads[(dname, 4)] = hd
        outputs = self._assemble(heads, b, h, w, dev, stream)
        sv['s2d'] = s2d[0]
        sv['shape'] = (b, h, w)
        torch._foreach_add_([L.bn.num_batches_tracked for L in self.layers.values() if L.bn is not None], 1)
        self.net._weights_epoch = getattr(self.net, '_weights_epoch', 0) + 1
        return outputs, (sv if save else None)

    # ------------------------------------------------------------------------------ backward pieces
    def _on_side(self, fn, *keep):
        """Run fn(stream_ptr) on the side stream, ordered after everything enqueued on the current stream so far; `keep`
        are the tensors it reads (held until backward() joins the streams: the main stream's allocator must not reuse them
        while the side stream still reads)."""
        if not self.overlap_wgrad:
            return fn(self._stream())
        if self._side is None:
            self._side = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        self._side_keep.extend(t for t in keep if t is not None)
        with torch.cuda.stream(self._side):
            fn(ctypes.c_void_p(self._side.cuda_stream))

    def _join_side(self):
        if self._side is not None and self.overlap_wgrad:
            torch.cuda.current_stream().wait_stream(self._side)
        self._side_keep.clear()

    def _wgrad(self, L, src0, src1, dz, grads, stream):
        """dW (and db) of one convolution (on the side stream, see _on_side)."""
        self._on_side(lambda st: self._wgrad_now(L, src0, src1, dz, grads, st), src0, src1, dz)

    def _reduce_ws_for(self, stream):
        if self._side is None or stream.value != self._side.cuda_stream:
            return self._reduce_ws()
        if self._red_ws_side is None:
            self._red_ws_side = torch.empty_like(self._reduce_ws())
        return self._red_ws_side

    def _wgrad_now(self, L, src0, src1, dz, grads, stream):
        """dW (and db) of one convolution from its gathered input and dZ [B][ho][wo][ld]."""
        b, ho, wo, ld = dz.shape
        d = self._desc(L, src0, src1, ho, wo)
        need = self.lib.fp_conv_wgrad_workspace_bytes(ctypes.byref(d))
        if self._wg_ws is None or self._wg_ws.numel() * 4 < need:
            self._wg_ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=dz.device)
        dw = self._gout(L.weight)
        self._prof('wgrad', L, b * ho * wo)
        _lib.check(self.lib.fp_conv_wgrad(ctypes.byref(d), _ptr(src0), _ptr(src1), _ptr(dz), ld, self.gfmt,
                                          1 if L.stem else 0, 0, _ptr(dw), _ptr(self._wg_ws), stream))
        self._prof_end()
        self.launches += 2
        grads[L.weight] = dw
        if L.bias is not None:
            if ld == L.cout:
                db = self._gout(L.bias)
                _lib.check(self.lib.fp_bias_grad(_ptr(dz), b * ho * wo, ld, self.gfmt, _ptr(db), _ptr(self._reduce_ws_for(stream)), stream))
            else:       # heads: dZ rows are padded to 8 channels, the bias has 2
                tmp = torch.empty(ld, dtype=torch.float32, device=dz.device)
                _lib.check(self.lib.fp_bias_grad(_ptr(dz), b * ho * wo, ld, self.gfmt, _ptr(tmp), _ptr(self._reduce_ws_for(stream)), stream))
                db = self._gout(L.bias)
                db.copy_(tmp[:L.cout])
            self.launches += 2
            grads[L.bias] = db

    def _dgrad_full(self, L, dz, stream):
        """Gradient of the reflection-padded virtual input: zero-padded full correlation of dZ with the rotated
        filter -> [B][ho+2][wo+2][cin] (decoder convs are 3x3 stride 1)."""
        b, ho, wo, ld = dz.shape
        cin = L.cin0 + L.cin1
        out = torch.empty((b, ho + 2, wo + 2, cin), dtype=self.gdtype, device=dz.device)
        d = FpConvDesc()
        d.batch, d.c0, d.h0, d.w0, d.up0, d.c1 = b, ld, ho, wo, 0, 0
        d.kh = d.kw = 3
        d.stride, d.pad_t, d.pad_l, d.pad_mode = 1, 2, 2, FP_PAD_ZERO
        d.ho, d.wo = ho + 2, wo + 2
        d.cout, d.kpad, d.npad, d.bn = cin, L.d_kpad, L.d_npad, L.d_bn
        d.act, d.dtype, d.out_f32, d.out_ld, d.res_ld, d.b_via_tma = FP_ACT_NONE, self.gfmt, 0, cin, 0, 1
        assert ld == L.cpad_out, (L.name, ld, L.cpad_out)
        self._prof('dgrad', L, b * ho * wo)
        _lib.check(self.lib.fp_conv_fwd(ctypes.byref(d), _ptr(dz), None, _ptr(L.w_flip), None, None, None, _ptr(out), stream))
        self._prof_end()
        self.launches += 1
        return out

    def _dgrad_plain(self, L, dz, out_hw, residual, stream):
        """Encoder data gradient (zero padding): stride-1 convs are a plain correlation with the rotated filter,
        stride-2 convs gather from the zero-inserted dZ (up0 = 2)."""
        b, ho, wo, ld = dz.shape
        cin = L.cin0
        H, W = out_hw
        out = torch.empty((b, H, W, cin), dtype=self.gdtype, device=dz.device)
        d = FpConvDesc()
        d.batch, d.c0, d.h0, d.w0, d.c1 = b, ld, ho, wo, 0
        d.up0 = 2 if L.stride == 2 else 0
        d.kh = d.kw = L.k
        d.stride, d.pad_mode = 1, FP_PAD_ZERO
        d.pad_t = d.pad_l = L.k - 1 - L.pad
        d.ho, d.wo = H, W
        d.cout, d.kpad, d.npad, d.bn = cin, L.d_kpad, L.d_npad, L.d_bn
        d.act, d.dtype, d.out_f32, d.out_ld, d.b_via_tma = FP_ACT_NONE, self.gfmt, 0, cin, 1
        d.res_ld = cin if residual is not None else 0
        self._prof('dgrad', L, b * ho * wo)
        _lib.check(self.lib.fp_conv_fwd(ctypes.byref(d), _ptr(dz), None, _ptr(L.w_flip), None, None, _ptr(residual),
                                        _ptr(out), stream))
        self._prof_end()
        self.launches += 1
        return out

    def _fold(self, dvp, hv, wv, coff, c, up, dst, accumulate, y_act, act, stream):
        b = dvp.shape[0]
        self.launches += 1
        _lib.check(self.lib.fp_fold_bwd(_ptr(dvp), b, hv, wv, dvp.shape[3], coff, c, 1 if up else 0, 1 if accumulate else 0,
                                        _ptr(y_act), act, self.fmt, self.gfmt, _ptr(dst), stream))

    def _bn_bwd(self, sv, L, dy, relu, want_g, grads, stream):
        src, z, y = sv[L.name]
        st = sv['bn'][L.name]
        b, ho, wo, c = z.shape
        dz = torch.empty(z.shape, dtype=self.gdtype, device=z.device)
        g = torch.empty(z.shape, dtype=self.gdtype, device=z.device) if want_g else None
        dgamma = self._gout(L.bn.weight)
        dbeta = self._gout(L.bn.bias)
        _lib.check(self.lib.fp_bn_bwd(_ptr(dy), _ptr(y), _ptr(z), _ptr(L.bn.weight.detach()), _ptr(st['mean']),
                                      _ptr(st['invstd']), 1 if relu else 0, b * ho * wo, c, self.fmt, self.gfmt, _ptr(dz),
                                      _ptr(g), _ptr(dgamma), _ptr(dbeta), _ptr(self._reduce_ws()), stream))
        self.launches += 3
        grads[L.bn.weight] = dgamma
        grads[L.bn.bias] = dbeta
        return dz, g

    # ------------------------------------------------------------------------------ backward
    def backward(self, douts, sv, on_decoders_done=None):       # on_decoders_done(side_stream, tag): tag 4 = both decoders,
        #                                                              3 = encoder layer4, 2 = layer3 finished their gradients
        """douts: {'1/8','1/4','1/2','1/1'} -> d(loss)/d(output) fp32 [B,4,H,W].  Returns {parameter: fp32 grad}; the
        gradients are multiplied by ``self.grad_scale`` (the loss scale; a power of two) -- the caller divides it out."""
        if sv is None:
            raise RuntimeError('backward() needs the saved state forward_train() returned')
        b, h, w = sv['shape']
        if self.loss_scale is not None:
            gs = float(self.loss_scale)
        elif self.gfmt == FP_F16:
            gs = 2.0 ** round(math.log2(32.0 * 4.0 * b * h * w))
        else:
            gs = 1.0
        self.grad_scale = gs
        stream = self._stream()
        lib = self.lib
        dev = self._dev()
        grads = {}
        feats = sv['feats']
        dfeat = [None] * 5          # gradient accumulators of the encoder features f0..f4 (from the decoders)

        def gnew(shape):
            return torch.empty(shape, dtype=self.gdtype, device=dev)

        for dname in ('mask', 'depth'):
            d = self.dec[dname]
            ch0 = 0 if dname == 'mask' else 2
            sigmoid = d['head4'].act != FP_ACT_NONE

            def head_dz(hi):
                L = d[f'head{hi}']
                src, _, hd = sv[L.name]
                hh, ww = hd.shape[1], hd.shape[2]
                dz = gnew((b, hh, ww, 8))
                _lib.check(lib.fp_heads_bwd(_ptr(douts[SCALES[hi - 1]]), ch0, _ptr(hd) if sigmoid else None, b, hh, ww,
                                            h // hh, self.gfmt, gs, _ptr(dz), stream))
                return L, src, dz

            def conv_bwd(L, dz):
                """wgrad + full dgrad of a decoder conv; returns the padded virtual-input gradient."""
                src0, src1, _ = sv[L.name]
                self._wgrad(L, src0, src1, dz, grads, stream)
                return self._dgrad_full(L, dz, stream)

            def head_bwd(L, src, dzh, dst, accumulate):
                """weight/bias gradient of a head (tiled CUDA-core kernel: N = 2 would waste 126 of the tensor-core
                kernel's 128 accumulator rows at the cost of a full 32-channel layer) and its data gradient (CUDA-core
                kernel: reflect adjoint + ELU' of src fused) into dst."""
                bb, hh_, ww_, cc_ = src.shape
                if self.head_tensor_cores_wgrad:
                    self._wgrad(L, src, None, dzh, grads, stream)
                else:
                    need = lib.fp_head_conv_wgrad_workspace_floats(cc_)
                    if self._head_ws is None or self._head_ws.numel() < need:
                        self._head_ws = torch.empty(need, dtype=torch.float32, device=src.device)
                    dw, db = self._gout(L.weight), self._gout(L.bias)

                    def run(st, L=L, src=src, dzh=dzh, dw=dw, db=db):
                        self._prof('wgrad', L, bb * hh_ * ww_)
                        _lib.check(lib.fp_head_conv_wgrad(_ptr(src), _ptr(dzh), bb, hh_, ww_, cc_, self.fmt, self.gfmt, 0,
                                                          _ptr(dw), _ptr(db), _ptr(self._head_ws), st))
                        self._prof_end()
                    self._on_side(run, src, dzh)
                    self.launches += 2
                    grads[L.weight], grads[L.bias] = dw, db
                if self.head_tensor_cores_dgrad:
                    dvp = self._dgrad_full(L, dzh, stream)
                    self._fold(dvp, hh_, ww_, 0, cc_, False, dst, accumulate, src, FP_ACT_ELU, stream)
                    return
                _lib.check(lib.fp_head_conv_dgrad(_ptr(dzh), _ptr(L.weight.detach()), bb, hh_, ww_, cc_, 1 if accumulate else 0,
                                                  _ptr(src), FP_ACT_ELU, self.fmt, self.gfmt, _ptr(dst), stream))
                self.launches += 1

            # ---- outconv4 chain at full resolution
            L, src, dzh = head_dz(4)
            o4c2_out = src
            dz = gnew(o4c2_out.shape)
            head_bwd(L, src, dzh, dz, False)
            L = d['o4.c2']
            dvp = conv_bwd(L, dz)
            o4c1_out = sv[L.name][0]
            dz = gnew(o4c1_out.shape)
            self._fold(dvp, h, w, 0, 32, False, dz, False, o4c1_out, FP_ACT_ELU, stream)
            L = d['o4.c1']
            dvp = conv_bwd(L, dz)
            x = sv[L.name][0]                       # block4 output [b, h/2, w/2, 64]
            dx = gnew(x.shape)
            self._fold(dvp, h, w, 0, 64, True, dx, False, None, FP_ACT_NONE, stream)
            del dvp, dz, dzh
            # ---- blocks 4..1 (each block output also feeds a head, except block1)
            for bi in range(4, 0, -1):
                hh, ww = x.shape[1], x.shape[2]
                co = x.shape[3]
                if bi >= 2:
                    L, src, dzh = head_dz(bi - 1)
                    head_bwd(L, src, dzh, dx, True)                                            # + head path, then ELU'
                    dz = dx
                else:
                    dz = gnew(x.shape)
                    _lib.check(lib.fp_act_bwd(_ptr(dx), _ptr(x), FP_ACT_ELU, x.numel(), self.fmt, self.gfmt, _ptr(dz), stream))
                L = d[f'b{bi}.post2']
                dvp = conv_bwd(L, dz)
                t3 = sv[L.name][0]
                dz = gnew(t3.shape)
                self._fold(dvp, hh, ww, 0, co, False, dz, False, t3, FP_ACT_ELU, stream)
                L = d[f'b{bi}.post1']
                dvp = conv_bwd(L, dz)
                t2, skip, _ = sv[L.name]
                dz = gnew(t2.shape)
                self._fold(dvp, hh, ww, 0, co, True, dz, False, t2, FP_ACT_ELU, stream)          # upsampled half
                fi = 4 - bi
                if dfeat[fi] is None:
                    dfeat[fi] = gnew(skip.shape)
                    self._fold(dvp, hh, ww, co, skip.shape[3], False, dfeat[fi], False, None, FP_ACT_NONE, stream)
                else:
                    self._fold(dvp, hh, ww, co, skip.shape[3], False, dfeat[fi], True, None, FP_ACT_NONE, stream)
                L = d[f'b{bi}.pre2']
                dvp = conv_bwd(L, dz)
                t = sv[L.name][0]
                dz = gnew(t.shape)
                self._fold(dvp, hh // 2, ww // 2, 0, co, False, dz, False, t, FP_ACT_ELU, stream)
                L = d[f'b{bi}.pre1']
                dvp = conv_bwd(L, dz)
                x = sv[L.name][0]                   # previous block output, or f4 for block1
                if bi >= 2:
                    dx = gnew(x.shape)
                    self._fold(dvp, hh // 2, ww // 2, 0, x.shape[3], False, dx, False, None, FP_ACT_NONE, stream)
                else:
                    if dfeat[4] is None:
                        dfeat[4] = gnew(x.shape)
                        self._fold(dvp, hh // 2, ww // 2, 0, x.shape[3], False, dfeat[4], False, None, FP_ACT_NONE, stream)
                    else:
                        self._fold(dvp, hh // 2, ww // 2, 0, x.shape[3], False, dfeat[4], True, None, FP_ACT_NONE, stream)
                del dvp

        if on_decoders_done is not None:
            # every decoder parameter gradient has been enqueued (weight gradients on the side stream): the data-parallel
            # trainer starts all-reducing that bucket while the encoder's backward runs
            on_decoders_done(self._side if self.overlap_wgrad else None, 4)
        # ---- encoder, last block first.  dy = gradient of the current block's output.
        dy = dfeat[4]
        for idx in range(len(self.enc_blocks) - 1, -1, -1):
            li, c1, ds, c2 = self.enc_blocks[idx]
            x_in = sv[c1.name][0]
            dz2, g = self._bn_bwd(sv, c2, dy, True, True, grads, stream)
            t1 = sv[c2.name][0]
            self._wgrad(c2, t1, None, dz2, grads, stream)
            dt1 = self._dgrad_plain(c2, dz2, (t1.shape[1], t1.shape[2]), None, stream)
            dz1, _ = self._bn_bwd(sv, c1, dt1, True, False, grads, stream)
            self._wgrad(c1, x_in, None, dz1, grads, stream)
            first_of_stage = idx == 0 or self.enc_blocks[idx - 1][0] != li
            if ds is not None:
                dzd, _ = self._bn_bwd(sv, ds, g, False, False, grads, stream)
                self._wgrad(ds, x_in, None, dzd, grads, stream)
                # the block input is the previous stage's feature: add the decoders' contribution here
                extra = dfeat[li] if first_of_stage else None
                dres = self._dgrad_plain(ds, dzd, (x_in.shape[1], x_in.shape[2]), extra, stream)
            else:
                dres = g
            dy = self._dgrad_plain(c1, dz1, (x_in.shape[1], x_in.shape[2]), dres, stream)
            del dz2, g, dt1, dz1
            if first_of_stage and on_decoders_done is not None and li >= 2:
                on_decoders_done(self._side if self.overlap_wgrad else None, li)     # layer4 (li 3) / layer3 (li 2) are done
        # dy is now the gradient of the max-pool output; f0 also received gradient from the decoders
        f0 = feats[0]
        if dfeat[0] is None:
            dfeat[0] = torch.zeros(f0.shape, dtype=self.gdtype, device=dev)
        _lib.check(lib.fp_maxpool3x3s2_bwd(_ptr(f0), _ptr(dy), b, f0.shape[1], f0.shape[2], 64, self.fmt, self.gfmt, 1,
                                           _ptr(dfeat[0]), stream))
        stem = self.layers['stem']
        dz0, _ = self._bn_bwd(sv, stem, dfeat[0], True, False, grads, stream)
        self._wgrad(stem, sv['s2d'], None, dz0, grads, stream)
        self._join_side()
        return grads
