"""GPU tests of the kernel variants behind one entry point: the same convolution / weight gradient must come out of every
dispatch variant the FP_* switches select, and out of torch on the same 16-bit operands.

* conv_rowstream_kernel (default for Cout <= 64 on wide rows; FP_ROWSTREAM=0 selects the strip kernel): one N = 3*Cout
  tcgen05.mma per (input row, dx, K step).
* conv_wgrad_roll_kernel (default for Cout <= 64; FP_WGRAD_ROLL=0 selects the strip weight gradient).
* FP_STRIP / FP_STRIP_TMA / FP_STRIP_WRES variants of the strip kernel.
The library caches the switches; fp_reload_options() re-reads the environment.
"""
import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

from footprints_b200 import _lib
from footprints_b200._lib import FpConvDesc

pytestmark = pytest.mark.gpu


class _Env:
    """Set FP_* switches for one block and make the library re-read them (and restore the defaults after)."""

    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        os.environ.update(self.kv)
        _lib.load().fp_reload_options()

    def __exit__(self, *a):
        for k in self.kv:
            os.environ.pop(k, None)
        _lib.load().fp_reload_options()
BF = 1
DT = torch.bfloat16


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize('cfg', [
    dict(c0=32, c1=0, up0=0, pad=1, pad_mode=1, cout=32, h=20, w=256, b=2),     # reflect, cp.async strips, BN 32 / CB 32
    dict(c0=32, c1=0, up0=0, pad=2, pad_mode=0, cout=64, h=19, w=300, b=2),     # full correlation (dgrad), TMA strips
    dict(c0=64, c1=0, up0=1, pad=1, pad_mode=1, cout=32, h=10, w=128, b=2),     # upsampled source, BN 32 / CB 64
    dict(c0=64, c1=64, up0=1, pad=1, pad_mode=1, cout=64, h=9, w=80, b=1),      # upsample + concat, 2 channel blocks
    dict(c0=64, c1=0, up0=0, pad=2, pad_mode=0, cout=64, h=13, w=130, b=3),     # ragged last item (Ho % H != 0)
    dict(c0=64, c1=0, up0=0, pad=1, pad_mode=0, cout=64, h=12, w=160, b=2),     # 48x160-class layer (Wo = 160: 62.5 % of 2 segments)
    dict(c0=128, c1=0, up0=0, pad=1, pad_mode=1, cout=64, h=6, w=80, b=2),      # 24x80-class layer, 2 channel blocks
    dict(c0=64, c1=0, up0=0, pad=2, pad_mode=0, cout=64, h=12, w=160, b=2),     # its data gradient (Wo = 162)
    dict(c0=64, c1=0, up0=0, pad=1, pad_mode=1, cout=64, h=9, w=128, b=1),      # 3 items: an unpaired last item (cp.async strips)
    dict(c0=32, c1=0, up0=0, pad=2, pad_mode=0, cout=32, h=18, w=100, b=1),     # 3 items, TMA strips
    dict(c0=32, c1=0, up0=0, pad=1, pad_mode=1, cout=32, h=6, w=128, b=1),      # a single item (one CTA, partner missing)
])
def test_rowstream_matches_strip_kernel_and_torch(cfg):
    lib = _lib.load()
    _lib.require_device()
    g = torch.Generator().manual_seed(11)
    b, h0, w0 = cfg['b'], cfg['h'], cfg['w']
    c0, c1, cout, pad = cfg['c0'], cfg['c1'], cfg['cout'], cfg['pad']
    hin, win = (2 * h0, 2 * w0) if cfg['up0'] else (h0, w0)
    cin = c0 + c1
    x0 = (torch.randn(b, c0, h0, w0, generator=g)).to(DT).float().cuda()
    x1 = (torch.randn(b, c1, hin, win, generator=g)).to(DT).float().cuda() if c1 else None
    wgt = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).to(DT).float().cuda()
    bias = (torch.randn(cout, generator=g) * 0.1).cuda()
    xin = F.interpolate(x0, scale_factor=2, mode='nearest') if cfg['up0'] else x0
    if x1 is not None:
        xin = torch.cat([xin, x1], 1)
    if cfg['pad_mode'] == 1:
        ref = F.conv2d(F.pad(xin, (1, 1, 1, 1), mode='reflect'), wgt, bias)
    else:
        ref = F.conv2d(xin, wgt, bias, padding=pad)
    ref = F.elu(ref)
    ho, wo = ref.shape[2], ref.shape[3]
    bn_t = 32 if cout <= 32 else 64
    kpad = (9 * cin + 63) // 64 * 64
    wp = torch.empty(bn_t, kpad, dtype=DT, device='cuda')
    _lib.check(lib.fp_pack_conv_weight(P(wgt), cout, cin, 3, 3, cin, bn_t, kpad, BF, 0, P(wp), None))
    d = FpConvDesc(batch=b, c0=c0, h0=h0, w0=w0, up0=cfg['up0'], c1=c1, kh=3, kw=3, stride=1, pad_t=pad, pad_l=pad,
                   pad_mode=cfg['pad_mode'], ho=ho, wo=wo, cout=cout, kpad=kpad, npad=bn_t, bn=bn_t, act=2, dtype=BF,
                   out_f32=0, out_ld=cout, res_ld=0, b_via_tma=1, w_dtype_plus1=0)
    s0 = x0.permute(0, 2, 3, 1).contiguous().to(DT)
    s1 = x1.permute(0, 2, 3, 1).contiguous().to(DT) if x1 is not None else None
    ones = torch.ones(cout, device='cuda')
    outs = {}
    for mode in ('0', '1'):
        with _Env(FP_ROWSTREAM=mode):
            y = torch.full((b, ho, wo, cout), float('nan'), dtype=DT, device='cuda')
            _lib.check(lib.fp_conv_fwd(ctypes.byref(d), P(s0), P(s1), P(wp), P(ones), P(bias), None, P(y), None))
            torch.cuda.synchronize()
            outs[mode] = y.float().permute(0, 3, 1, 2)
    tol = 2e-2 * ref.abs().max().item()
    assert (outs['0'] - ref).abs().max().item() <= tol
    assert (outs['1'] - ref).abs().max().item() <= tol
    assert (outs['1'] - outs['0']).abs().max().item() <= tol


@pytest.mark.parametrize('env', [{'FP_STRIP': '0'}, {'FP_STRIP_TMA': '0'}, {'FP_STRIP_WRES': '0'},
                                 {'FP_STRIP_TMA': '0', 'FP_STRIP_WRES': '0'}])
@pytest.mark.parametrize('cfg', [
    dict(c0=32, c1=0, up0=0, pad=2, pad_mode=0, cout=32, h=12, w=200, b=2),
    dict(c0=64, c1=64, up0=1, pad=1, pad_mode=1, cout=64, h=6, w=80, b=2),
])
def test_dispatch_variants_agree(cfg, env):
    lib = _lib.load()
    _lib.require_device()
    g = torch.Generator().manual_seed(12)
    b, h0, w0 = cfg['b'], cfg['h'], cfg['w']
    c0, c1, cout, pad = cfg['c0'], cfg['c1'], cfg['cout'], cfg['pad']
    hin, win = (2 * h0, 2 * w0) if cfg['up0'] else (h0, w0)
    cin = c0 + c1
    s0 = torch.randn(b, h0, w0, c0, generator=g).to(DT).cuda()
    s1 = torch.randn(b, hin, win, c1, generator=g).to(DT).cuda() if c1 else None
    wgt = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    ho, wo = hin + 2 * pad - 2, win + 2 * pad - 2
    bn_t = 32 if cout <= 32 else 64
    kpad = (9 * cin + 63) // 64 * 64
    wp = torch.empty(bn_t, kpad, dtype=DT, device='cuda')
    _lib.check(lib.fp_pack_conv_weight(P(wgt), cout, cin, 3, 3, cin, bn_t, kpad, BF, 0, P(wp), None))
    d = FpConvDesc(batch=b, c0=c0, h0=h0, w0=w0, up0=cfg['up0'], c1=c1, kh=3, kw=3, stride=1, pad_t=pad, pad_l=pad,
                   pad_mode=cfg['pad_mode'], ho=ho, wo=wo, cout=cout, kpad=kpad, npad=bn_t, bn=bn_t, act=0, dtype=BF,
                   out_f32=0, out_ld=cout, res_ld=0, b_via_tma=1, w_dtype_plus1=0)

    def run():
        y = torch.full((b, ho, wo, cout), float('nan'), dtype=DT, device='cuda')
        _lib.check(lib.fp_conv_fwd(ctypes.byref(d), P(s0), P(s1), P(wp), None, None, None, P(y), None))
        torch.cuda.synchronize()
        return y.float()

    with _Env(FP_ROWSTREAM='0'):
        base = run()
    with _Env(FP_ROWSTREAM='0', **env):
        var = run()
    with _Env():
        dflt = run()
    assert torch.isfinite(dflt).all()
    assert (dflt - base).abs().max().item() <= 2e-2 * base.abs().max().item()
    assert torch.isfinite(var).all()
    assert (var - base).abs().max().item() <= 2e-2 * base.abs().max().item()


@pytest.mark.parametrize('cfg', [
    dict(c0=64, c1=0, up0=0, pad=1, pad_mode=1, cout=64, h=9, w=160, b=2),       # CB 64, one 160-pixel segment per row
    dict(c0=32, c1=0, up0=0, pad=1, pad_mode=1, cout=32, h=7, w=320, b=2),       # CB 32
    dict(c0=64, c1=64, up0=1, pad=1, pad_mode=1, cout=64, h=5, w=80, b=2),       # upsample + concat: 2 channel blocks -> 2 CTAs groups
    dict(c0=32, c1=0, up0=0, pad=1, pad_mode=1, cout=2, h=6, w=200, b=3),        # a head (cout 2, dz_ld 8), ragged segments
    dict(c0=64, c1=0, up0=0, pad=1, pad_mode=0, cout=48, h=11, w=100, b=1),      # zero padding
])
def test_wgrad_roll_matches_default_and_autograd(cfg):
    lib = _lib.load()
    _lib.require_device()
    g = torch.Generator().manual_seed(13)
    b, h0, w0 = cfg['b'], cfg['h'], cfg['w']
    c0, c1, cout, pad = cfg['c0'], cfg['c1'], cfg['cout'], cfg['pad']
    hin, win = (2 * h0, 2 * w0) if cfg['up0'] else (h0, w0)
    cin = c0 + c1
    x0 = torch.randn(b, c0, h0, w0, generator=g).to(DT).float().cuda()
    x1 = torch.randn(b, c1, hin, win, generator=g).to(DT).float().cuda() if c1 else None
    wgt = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).to(DT).float().cuda().requires_grad_(True)
    xin = F.interpolate(x0, scale_factor=2, mode='nearest') if cfg['up0'] else x0
    if x1 is not None:
        xin = torch.cat([xin, x1], 1)
    if cfg['pad_mode'] == 1:
        z = F.conv2d(F.pad(xin, (1, 1, 1, 1), mode='reflect'), wgt)
    else:
        z = F.conv2d(xin, wgt, padding=pad)
    dz = torch.randn(z.shape, generator=g).to(DT).float().cuda()
    z.backward(dz)
    ho, wo = z.shape[2], z.shape[3]
    ld = max(8, cout)
    dz16 = torch.zeros(b, ho, wo, ld, dtype=DT, device='cuda')
    dz16[..., :cout] = dz.permute(0, 2, 3, 1).to(DT)
    d = FpConvDesc(batch=b, c0=c0, h0=h0, w0=w0, up0=cfg['up0'], c1=c1, kh=3, kw=3, stride=1, pad_t=pad, pad_l=pad,
                   pad_mode=cfg['pad_mode'], ho=ho, wo=wo, cout=cout, kpad=(9 * cin + 63) // 64 * 64, npad=0, bn=64, act=0,
                   dtype=BF, out_f32=0, out_ld=cout, res_ld=0, b_via_tma=1, w_dtype_plus1=0)
    s0 = x0.permute(0, 2, 3, 1).contiguous().to(DT)
    s1 = x1.permute(0, 2, 3, 1).contiguous().to(DT) if x1 is not None else None
    outs = {}
    for mode in ('0', '1'):
        with _Env(FP_WGRAD_ROLL=mode):
            nbytes = lib.fp_conv_wgrad_workspace_bytes(ctypes.byref(d))
            wsp = torch.empty(nbytes // 4 + 1, dtype=torch.float32, device='cuda')
            dw = torch.full((cout, cin, 3, 3), float('nan'), dtype=torch.float32, device='cuda')
            _lib.check(lib.fp_conv_wgrad(ctypes.byref(d), P(s0), P(s1), P(dz16), ld, BF, 0, 0, P(dw), P(wsp), None))
            torch.cuda.synchronize()
            outs[mode] = dw
    tol = 2e-3 * wgt.grad.abs().max().item() + 1e-4
    assert (outs['0'] - wgt.grad).abs().max().item() <= tol
    assert (outs['1'] - wgt.grad).abs().max().item() <= tol


@pytest.mark.parametrize('cfg', [
    dict(c=32, h=20, w=256, b=2, sigmoid=1),       # head4-class: CB 32
    dict(c=64, h=12, w=160, b=2, sigmoid=0),       # head2/3-class: CB 64, 62.5 %-used segments
    dict(c=128, h=7, w=80, b=3, sigmoid=1),        # head1-class: two channel blocks, ragged item
])
def test_rowstream_head_n16_matches_cuda_core_head_and_torch(cfg):
    """The 2-channel heads on the tensor-core row-streaming kernel (N padded to 16, fp32 NHWC [.,.,.,2] output,
    bias + optional sigmoid in the epilogue) against the CUDA-core head kernel and torch."""
    lib = _lib.load()
    _lib.require_device()
    g = torch.Generator().manual_seed(21)
    b, h, w, c = cfg['b'], cfg['h'], cfg['w'], cfg['c']
    x = torch.randn(b, c, h, w, generator=g).to(DT).float().cuda()
    wgt = ((torch.rand(2, c, 3, 3, generator=g) * 2 - 1) / (9 * c) ** 0.5).cuda()
    bias = (torch.randn(2, generator=g) * 0.1).cuda()
    ref = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), wgt.to(DT).float(), bias)
    if cfg['sigmoid']:
        ref = torch.sigmoid(ref)
    kpad = (9 * c + 63) // 64 * 64
    wp = torch.empty(16, kpad, dtype=DT, device='cuda')
    _lib.check(lib.fp_pack_conv_weight(P(wgt), 2, c, 3, 3, c, 16, kpad, BF, 0, P(wp), None))
    d = FpConvDesc(batch=b, c0=c, h0=h, w0=w, up0=0, c1=0, kh=3, kw=3, stride=1, pad_t=1, pad_l=1, pad_mode=1, ho=h, wo=w,
                   cout=2, kpad=kpad, npad=16, bn=16, act=3 if cfg['sigmoid'] else 0, dtype=BF, out_f32=1, out_ld=2,
                   res_ld=0, b_via_tma=1, w_dtype_plus1=0)
    xs = x.permute(0, 2, 3, 1).contiguous().to(DT)
    outs = {}
    for mode in ('1', '0'):            # row-streaming N = 16 kernel / general kernel<16>
        with _Env(FP_ROWSTREAM=mode):
            y = torch.full((b, h, w, 2), float('nan'), dtype=torch.float32, device='cuda')
            _lib.check(lib.fp_conv_fwd(ctypes.byref(d), P(xs), None, P(wp), None, P(bias), None, P(y), None))
            torch.cuda.synchronize()
            outs[mode] = y.permute(0, 3, 1, 2)
    yc = torch.empty((b, h, w, 2), dtype=torch.float32, device='cuda')
    _lib.check(lib.fp_head_conv_fwd(P(xs), P(wgt), P(bias), b, h, w, c, cfg['sigmoid'], BF, P(yc), None))
    torch.cuda.synchronize()
    for name, got in (('rowstream', outs['1']), ('general', outs['0'])):
        assert torch.isfinite(got).all(), name
        assert (got - ref).abs().max().item() <= 5e-4, (name, (got - ref).abs().max().item())
    assert (yc.permute(0, 3, 1, 2) - outs['1']).abs().max().item() <= 5e-3      # CUDA-core kernel keeps fp32 weights


@pytest.mark.parametrize('shape', [(2, 9, 14), (2, 37, 70)])                # one ragged tile; several tiles, ragged edges
@pytest.mark.parametrize('C,act,accumulate', [(32, 2, 0), (64, 2, 1), (128, 1, 1), (16, 0, 0), (32, 2, 1)])
def test_head_dgrad_epilogue_variants_match_torch(C, act, accumulate, shape):
    """fp_head_conv_dgrad with fp16 activations and gradients (the training default): the default kernel (epilogue operands
    prefetched before the FMA loop, packed-half epilogue), FP_HEAD_EPI=0 (loads after the loop, fp32 epilogue) and
    FP_HEAD_HMATH=0 (fp32 FMA loop) against autograd in float64 -- with ELU' / ReLU' of the stored activation and with
    accumulation into an existing gradient."""
    lib = _lib.load()
    _lib.require_device()
    g = torch.Generator().manual_seed(C + act)
    b, h, w = shape
    x = torch.randn(b, C, h, w, generator=g).double().requires_grad_(True)
    wgt = (torch.randn(2, C, 3, 3, generator=g) / (9 * C) ** 0.5).float()
    dz = torch.randn(b, 2, h, w, generator=g).half()
    F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), wgt.double()).backward(dz.double())
    ref = x.grad.clone()
    old = torch.randn(b, C, h, w, generator=g).half()
    if accumulate:                      # dst holds the other consumer's gradient w.r.t. the activation ...
        ref = ref + old.double()
    yact = None
    if act:                             # ... and the activation's derivative applies to the sum
        yact = F.elu(torch.randn(b, C, h, w, generator=g) * 1.5).half()      # a stored activation: values in (-1, ..)
        yact[0, :, 0, :4] = -1.0                                             # saturated ELU outputs: ELU' = 0 exactly
        yd = yact.double()
        ref = ref * (torch.where(yd > 0, torch.ones_like(yd), yd + 1) if act == 2 else (yd > 0).double())
    to_nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()          # noqa: E731
    dz8 = torch.zeros(b, h, w, 8, dtype=torch.float16, device='cuda')
    dz8[..., :2] = to_nhwc(dz)
    y16 = to_nhwc(yact) if yact is not None else None
    outs = {}
    for name, env in (('default', {}), ('epi0', {'FP_HEAD_EPI': '0'}), ('fp32_loop', {'FP_HEAD_HMATH': '0'}),
                      ('fp32_loop_epi0', {'FP_HEAD_HMATH': '0', 'FP_HEAD_EPI': '0'})):
        with _Env(**env):
            dx = to_nhwc(old).clone()
            _lib.check(lib.fp_head_conv_dgrad(P(dz8), P(wgt.cuda()), b, h, w, C, accumulate, P(y16), act, 0, 0, P(dx), None))
            torch.cuda.synchronize()
        outs[name] = dx.permute(0, 3, 1, 2).double().cpu()
        assert torch.isfinite(outs[name]).all()
        err = (outs[name] - ref).abs().max().item()
        bar = (1e-2 if 'fp32' not in name else 2e-3) * ref.abs().max().item()
        print(f'head dgrad C={C} act={act} acc={accumulate} {shape} {name}: max abs err {err:.2e} (bar {bar:.2e})')
        assert err <= bar, name
    # the two fp32-loop variants differ only in where the operands are loaded
    assert torch.equal(outs['fp32_loop'], outs['fp32_loop_epi0'])
