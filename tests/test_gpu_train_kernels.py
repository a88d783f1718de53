"""GPU parity of the individual training kernels (called through the C ABI) against plain PyTorch fp32
references of the same ops (autograd on the GPU).  Operands are pre-rounded to the 16-bit storage type so
the only differences are accumulation order and the 16-bit rounding of outputs."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from footprints_b200 import _lib
from footprints_b200._lib import FpConvDesc

pytestmark = pytest.mark.gpu
BF = 1  # FP_BF16
DT = torch.bfloat16


@pytest.fixture(autouse=True, params=['bf16', 'fp16'])
def _storage_type(request):
    """Every kernel test runs for both 16-bit storage types (training defaults to fp16 + loss scale; bf16 selectable)."""
    global BF, DT
    BF, DT = (1, torch.bfloat16) if request.param == 'bf16' else (0, torch.float16)
    yield
    BF, DT = 1, torch.bfloat16
    torch.cuda.synchronize()
    _KEEP.clear()


_KEEP = []


def P(t):
    """Device pointer of `t`.  The tensor is kept alive until the test ends: the launches are asynchronous, and a
    temporary such as ``nhwc(x)`` would otherwise be freed -- and its block handed to the next temporary -- while a
    kernel still reads it."""
    if t is None:
        return None
    _KEEP.append(t)
    return ctypes.c_void_p(t.data_ptr())


def nhwc(t):          # NCHW fp32 -> NHWC bf16
    return t.permute(0, 2, 3, 1).contiguous().to(DT)


def nchw(t):          # NHWC 16-bit -> NCHW fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def rnd(*shape, g, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).to(DT).float().cuda()


@pytest.fixture(scope='module')
def lib():
    l = _lib.load()
    _lib.require_device()
    return l


def ws(lib, c=512):
    return torch.empty(lib.fp_reduce_workspace_floats(c), dtype=torch.float32, device='cuda')


@pytest.mark.parametrize('C,M', [(64, 4999), (256, 700), (512, 129)])
def test_bn_stats_and_apply(lib, C, M):
    g = torch.Generator().manual_seed(0)
    z = (rnd(M, C, g=g, scale=2.0) + 0.7).to(DT).float()
    z16 = z.to(DT)
    gamma, beta = torch.rand(C, generator=g).cuda() + 0.5, torch.randn(C, generator=g).cuda() * 0.1
    rm, rv = torch.randn(C, generator=g).cuda() * 0.1, torch.rand(C, generator=g).cuda() + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    res = rnd(M, C, g=g)
    ref = F.relu(F.batch_norm(z.view(M, C, 1, 1), rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5).view(M, C) + res)
    outs = [torch.empty(C, device='cuda') for _ in range(4)]
    _lib.check(lib.fp_bn_stats(P(z16), M, C, BF, P(gamma), P(beta), P(rm), P(rv), 0.1, 1e-5, *[P(o) for o in outs],
                               P(ws(lib)), None))
    y = torch.empty(M, C, dtype=DT, device='cuda')
    _lib.check(lib.fp_bn_apply(P(z16), P(outs[0]), P(outs[1]), P(res.to(DT)), 1, M, C, BF, P(y), None))
    torch.cuda.synchronize()
    assert torch.allclose(rm, rm_ref, rtol=1e-4, atol=1e-5) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-5)
    assert (y.float() - ref).abs().max() <= 1.0 / 64      # bf16 output rounding of O(1) values
    assert torch.allclose(outs[2], z.mean(0), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('relu', [1, 0])
def test_bn_backward(lib, relu):
    g = torch.Generator().manual_seed(1)
    M, C = 3001, 128
    z = (rnd(M, C, g=g, scale=1.5) + 0.3).to(DT).float().requires_grad_(True)
    gamma = (torch.rand(C, generator=g).cuda() + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g).cuda() * 0.1).requires_grad_(True)
    dy = rnd(M, C, g=g)
    pre = F.batch_norm(z.view(M, C, 1, 1), None, None, gamma, beta, True, 0.1, 1e-5).view(M, C)
    y = F.relu(pre) if relu else pre
    y.backward(dy)
    mean = z.detach().mean(0)
    invstd = torch.rsqrt(z.detach().var(0, unbiased=False) + 1e-5)
    y16, z16 = y.detach().to(DT), z.detach().to(DT)
    dz = torch.empty(M, C, dtype=DT, device='cuda')
    gout = torch.empty(M, C, dtype=DT, device='cuda')
    dgamma, dbeta = torch.empty(C, device='cuda'), torch.empty(C, device='cuda')
    _lib.check(lib.fp_bn_bwd(P(dy.to(DT)), P(y16), P(z16), P(gamma.detach()), P(mean), P(invstd), relu, M, C, BF, BF, P(dz),
                             P(gout), P(dgamma), P(dbeta), P(ws(lib)), None))
    torch.cuda.synchronize()
    assert torch.allclose(dbeta, beta.grad, rtol=2e-3, atol=2e-3)
    assert torch.allclose(dgamma, gamma.grad, rtol=2e-3, atol=5e-2)
    assert (dz.float() - z.grad).abs().max() <= 2e-2 * z.grad.abs().max()
    gref = dy * (y.detach() > 0) if relu else dy
    assert (gout.float() - gref).abs().max() <= 1e-2


@pytest.mark.parametrize('up', [True, False])
def test_fold_is_adjoint_of_up_concat_reflect(lib, up):
    g = torch.Generator().manual_seed(2)
    b, h, w, c0, c1 = 2, 6, 10, 16, 24
    hv, wv = (2 * h, 2 * w) if up else (h, w)
    x0 = rnd(b, c0, h, w, g=g).requires_grad_(True)
    x1 = rnd(b, c1, hv, wv, g=g).requires_grad_(True)
    v = torch.cat([F.interpolate(x0, scale_factor=2, mode='nearest') if up else x0, x1], 1)
    vp = F.pad(v, (1, 1, 1, 1), mode='reflect')
    dvp = rnd(*vp.shape, g=g)
    vp.backward(dvp)
    dvp16 = nhwc(dvp)
    d0 = torch.empty(b, h, w, c0, dtype=DT, device='cuda')
    d1 = torch.empty(b, hv, wv, c1, dtype=DT, device='cuda')
    _lib.check(lib.fp_fold_bwd(P(dvp16), b, hv, wv, c0 + c1, 0, c0, 1 if up else 0, 0, None, 0, BF, BF, P(d0), None))
    _lib.check(lib.fp_fold_bwd(P(dvp16), b, hv, wv, c0 + c1, c0, c1, 0, 0, None, 0, BF, BF, P(d1), None))
    torch.cuda.synchronize()
    assert (nchw(d0) - x0.grad).abs().max() <= 4e-2 * x0.grad.abs().max()
    assert (nchw(d1) - x1.grad).abs().max() <= 4e-2 * x1.grad.abs().max()
    # accumulate + ELU' from the saved output
    yact = rnd(b, c1, hv, wv, g=g)
    prev = d1.clone()
    _lib.check(lib.fp_fold_bwd(P(dvp16), b, hv, wv, c0 + c1, c0, c1, 0, 1, P(nhwc(yact)), 2, BF, BF, P(d1), None))
    torch.cuda.synchronize()
    ya = nhwc(yact).float()
    ref = (prev.float() + nhwc(x1.grad).float()) * torch.where(ya > 0, torch.ones_like(ya), ya + 1)
    assert (d1.float() - ref).abs().max() <= 4e-2 * ref.abs().max()


@pytest.mark.parametrize('scale,sigmoid', [(8, True), (4, False), (2, True), (1, True)])
def test_heads_backward(lib, scale, sigmoid):
    g = torch.Generator().manual_seed(3)
    b, h, w = 2, 6, 10
    H, W = h * scale, w * scale
    pre = rnd(b, 2, h, w, g=g).requires_grad_(True)
    post = torch.sigmoid(pre) if sigmoid else pre
    out = F.interpolate(post, scale_factor=scale, mode='bilinear', align_corners=False) if scale > 1 else post * 1.0
    dout_full = torch.randn(b, 4, H, W, generator=g).cuda()
    ch0 = 2 if sigmoid else 0
    out.backward(dout_full[:, ch0:ch0 + 2])
    head_post = post.detach().permute(0, 2, 3, 1).contiguous()
    dz = torch.empty(b, h, w, 8, dtype=DT, device='cuda')
    gs = 0.25 * scale                 # the loss scale (a power of two) multiplies the result exactly
    _lib.check(lib.fp_heads_bwd(P(dout_full), ch0, P(head_post) if sigmoid else None, b, h, w, scale, BF, gs, P(dz), None))
    torch.cuda.synchronize()
    got = dz.float()[..., :2].permute(0, 3, 1, 2) / gs
    assert (got - pre.grad).abs().max() <= 1e-2 * pre.grad.abs().max() + 1e-6
    assert dz.float()[..., 2:].abs().max() == 0


def test_maxpool_backward(lib):
    g = torch.Generator().manual_seed(4)
    b, h, w, c = 2, 12, 20, 64
    x = rnd(b, c, h, w, g=g).requires_grad_(True)
    y = F.max_pool2d(x, 3, 2, 1)
    dy = rnd(*y.shape, g=g)
    y.backward(dy)
    dx = torch.zeros(b, h, w, c, dtype=DT, device='cuda')
    _lib.check(lib.fp_maxpool3x3s2_bwd(P(nhwc(x.detach())), P(nhwc(dy)), b, h, w, c, BF, BF, 0, P(dx), None))
    torch.cuda.synchronize()
    assert (nchw(dx) - x.grad).abs().max() <= 2e-2 * x.grad.abs().max()


def _fwd_desc(b, c0, h0, w0, up0, c1, k, stride, pad, pad_mode, cout):
    hin, win = (2 * h0, 2 * w0) if up0 == 1 else (h0, w0)
    d = FpConvDesc(batch=b, c0=c0, h0=h0, w0=w0, up0=up0, c1=c1, kh=k, kw=k, stride=stride, pad_t=pad, pad_l=pad,
                   pad_mode=pad_mode, ho=(hin + 2 * pad - k) // stride + 1, wo=(win + 2 * pad - k) // stride + 1, cout=cout,
                   kpad=(k * k * (c0 + c1) + 63) // 64 * 64, npad=0, bn=64, act=0, dtype=BF, out_f32=0, out_ld=cout, res_ld=0,
                   b_via_tma=1, w_dtype_plus1=0)
    return d, hin, win


@pytest.mark.parametrize('cfg', [
    dict(c0=64, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=0, cout=64),
    dict(c0=64, c1=0, up0=0, k=3, stride=2, pad=1, pad_mode=0, cout=128),
    dict(c0=64, c1=0, up0=0, k=1, stride=2, pad=0, pad_mode=0, cout=128),
    dict(c0=32, c1=32, up0=1, k=3, stride=1, pad=1, pad_mode=1, cout=32),
    dict(c0=256, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=0, cout=256),
    dict(c0=64, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=1, cout=2),
    # wide rows -> row-strip kernels (forward/dgrad and wgrad): 1.x segments per row, ragged tail, concat, upsample
    dict(c0=64, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=0, cout=64, w0=160, h0=6, b=2),
    dict(c0=64, c1=64, up0=1, k=3, stride=1, pad=1, pad_mode=1, cout=64, w0=80, h0=4, b=2),
    dict(c0=128, c1=64, up0=1, k=3, stride=1, pad=1, pad_mode=1, cout=256, w0=60, h0=3, b=2),
    dict(c0=192, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=1, cout=32, w0=250, h0=3, b=1),
    dict(c0=32, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=1, cout=32, w0=160, h0=6, b=2),      # 64-byte strip rows
    dict(c0=32, c1=32, up0=1, k=3, stride=1, pad=1, pad_mode=1, cout=2, w0=80, h0=4, b=2),
    dict(c0=16, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=0, cout=16, w0=130, h0=5, b=2),      # 32-byte strip rows
    # weight-gradient strips with 80- and 48-pixel segments (K-flexible), 4 channel blocks in 2 groups
    dict(c0=128, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=0, cout=128, w0=80, h0=4, b=2),
    dict(c0=256, c1=0, up0=0, k=3, stride=1, pad=1, pad_mode=0, cout=256, w0=40, h0=3, b=2),
])
def test_conv_wgrad_and_dgrad(lib, cfg):
    """Weight gradient (fp_conv_wgrad) and data gradient (fp_conv_fwd with the rotated filter) against autograd."""
    g = torch.Generator().manual_seed(5)
    b, h0, w0 = cfg.get('b', 3), cfg.get('h0', 10), cfg.get('w0', 14)
    d, hin, win = _fwd_desc(b, cfg['c0'], h0, w0, cfg['up0'], cfg['c1'], cfg['k'], cfg['stride'], cfg['pad'],
                            cfg['pad_mode'], cfg['cout'])
    cin, cout, k = cfg['c0'] + cfg['c1'], cfg['cout'], cfg['k']
    x0 = rnd(b, cfg['c0'], h0, w0, g=g).requires_grad_(True)
    x1 = rnd(b, cfg['c1'], hin, win, g=g).requires_grad_(True) if cfg['c1'] else None
    wgt = rnd(cout, cin, k, k, g=g, scale=1.0 / (cin * k * k) ** 0.5).requires_grad_(True)
    xin = F.interpolate(x0, scale_factor=2, mode='nearest') if cfg['up0'] else x0
    if x1 is not None:
        xin = torch.cat([xin, x1], 1)
    if cfg['pad_mode'] == 1:
        z = F.conv2d(F.pad(xin, (1, 1, 1, 1), mode='reflect'), wgt, None, stride=cfg['stride'])
    else:
        z = F.conv2d(xin, wgt, None, stride=cfg['stride'], padding=cfg['pad'])
    dz = rnd(*z.shape, g=g)
    z.backward(dz)
    ho, wo = z.shape[2], z.shape[3]
    ld = max(8, cout)
    dz16 = torch.zeros(b, ho, wo, ld, dtype=DT, device='cuda')
    dz16[..., :cout] = nhwc(dz)
    # ---- wgrad
    nbytes = lib.fp_conv_wgrad_workspace_bytes(ctypes.byref(d))
    wsp = torch.empty(nbytes // 4 + 1, dtype=torch.float32, device='cuda')
    dw = torch.empty(cout, cin, k, k, dtype=torch.float32, device='cuda')
    s0 = nhwc(x0.detach())
    s1 = nhwc(x1.detach()) if x1 is not None else None
    _lib.check(lib.fp_conv_wgrad(ctypes.byref(d), P(s0), P(s1), P(dz16), ld, BF, 0, 0, P(dw), P(wsp), None))
    torch.cuda.synchronize()
    assert (dw - wgt.grad).abs().max() <= 2e-3 * wgt.grad.abs().max() + 1e-4, (dw - wgt.grad).abs().max()
    # ---- dgrad
    bn_t = 16 if cin <= 16 else (32 if cin <= 32 else (64 if cin <= 64 else (128 if cin <= 128 else 256)))
    npad = (cin + bn_t - 1) // bn_t * bn_t
    kpad = (k * k * ld + 63) // 64 * 64
    wflip = torch.empty(npad, kpad, dtype=DT, device='cuda')
    _lib.check(lib.fp_pack_conv_weight(P(wgt.detach()), cout, cin, k, k, ld, npad, kpad, BF, 1, P(wflip), None))
    if cfg['pad_mode'] == 1:
        # reflection-padded conv: full correlation -> padded virtual-input gradient -> fold
        dd = FpConvDesc(batch=b, c0=ld, h0=ho, w0=wo, up0=0, c1=0, kh=3, kw=3, stride=1, pad_t=2, pad_l=2, pad_mode=0,
                        ho=ho + 2, wo=wo + 2, cout=cin, kpad=kpad, npad=npad, bn=bn_t, act=0, dtype=BF, out_f32=0,
                        out_ld=cin, res_ld=0, b_via_tma=1, w_dtype_plus1=0)
        dvp = torch.empty(b, ho + 2, wo + 2, cin, dtype=DT, device='cuda')
        _lib.check(lib.fp_conv_fwd(ctypes.byref(dd), P(dz16), None, P(wflip), None, None, None, P(dvp), None))
        dx0 = torch.empty(b, h0, w0, cfg['c0'], dtype=DT, device='cuda')
        _lib.check(lib.fp_fold_bwd(P(dvp), b, hin, win, cin, 0, cfg['c0'], cfg['up0'], 0, None, 0, BF, BF, P(dx0), None))
    else:
        dd = FpConvDesc(batch=b, c0=ld, h0=ho, w0=wo, up0=2 if cfg['stride'] == 2 else 0, c1=0, kh=k, kw=k, stride=1,
                        pad_t=k - 1 - cfg['pad'], pad_l=k - 1 - cfg['pad'], pad_mode=0, ho=h0, wo=w0, cout=cin, kpad=kpad,
                        npad=npad, bn=bn_t, act=0, dtype=BF, out_f32=0, out_ld=cin, res_ld=0, b_via_tma=1, w_dtype_plus1=0)
        dx0 = torch.empty(b, h0, w0, cin, dtype=DT, device='cuda')
        _lib.check(lib.fp_conv_fwd(ctypes.byref(dd), P(dz16), None, P(wflip), None, None, None, P(dx0), None))
    torch.cuda.synchronize()
    assert (nchw(dx0) - x0.grad).abs().max() <= 2e-2 * x0.grad.abs().max()


def test_stem_wgrad(lib):
    g = torch.Generator().manual_seed(6)
    b, h, w = 2, 32, 64
    img = torch.rand(b, 3, h, w, generator=g).cuda()
    s2d = torch.empty(b, h // 2, w // 2, 16, dtype=DT, device='cuda')
    _lib.check(lib.fp_stem_prep(P(img), b, h, w, BF, P(s2d), None))
    xn = ((img - 0.45) / 0.225).to(DT).float()
    wgt = rnd(64, 3, 7, 7, g=g, scale=0.1).requires_grad_(True)
    z = F.conv2d(xn, wgt, None, stride=2, padding=3)
    dz = rnd(*z.shape, g=g)
    z.backward(dz)
    d = FpConvDesc(batch=b, c0=16, h0=h // 2, w0=w // 2, up0=0, c1=0, kh=4, kw=4, stride=1, pad_t=2, pad_l=2, pad_mode=0,
                   ho=h // 2, wo=w // 2, cout=64, kpad=256, npad=64, bn=64, act=0, dtype=BF, out_f32=0, out_ld=64, res_ld=0,
                   b_via_tma=1, w_dtype_plus1=0)
    wsp = torch.empty(lib.fp_conv_wgrad_workspace_bytes(ctypes.byref(d)) // 4 + 1, dtype=torch.float32, device='cuda')
    dw = torch.empty(64, 3, 7, 7, device='cuda')
    _lib.check(lib.fp_conv_wgrad(ctypes.byref(d), P(s2d), None, P(nhwc(dz)), 64, BF, 1, 0, P(dw), P(wsp), None))
    torch.cuda.synchronize()
    assert (dw - wgt.grad).abs().max() <= 2e-3 * wgt.grad.abs().max()


def test_bias_grad_and_adam(lib):
    g = torch.Generator().manual_seed(7)
    M, C = 7777, 32
    dz = rnd(M, C, g=g)
    db = torch.empty(C, device='cuda')
    _lib.check(lib.fp_bias_grad(P(dz.to(DT)), M, C, BF, P(db), P(ws(lib)), None))
    torch.cuda.synchronize()
    assert torch.allclose(db, dz.sum(0), rtol=1e-4, atol=1e-3)
    n = 100003
    p = torch.randn(n, generator=g).cuda()
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-4)
    m, v = torch.zeros(n, device='cuda'), torch.zeros(n, device='cuda')
    for step in range(1, 4):
        grad = torch.randn(n, generator=g).cuda() * 0.01
        ref.grad = grad.clone()
        opt.step()
        _lib.check(lib.fp_adam_step(P(p), P(grad), P(m), P(v), n, 1e-4, 0.9, 0.999, 1e-8, step, 1.0, None))
    torch.cuda.synchronize()
    assert (p - ref.detach()).abs().max() <= 1e-6


@pytest.mark.parametrize('shape', [(2, 9, 14), (2, 37, 70)])        # one ragged tile; several tiles, ragged edges
@pytest.mark.parametrize('C,sigmoid', [(32, 1), (64, 0), (128, 1), (16, 0)])
def test_head_kernels_match_autograd(lib, C, sigmoid, shape):
    """fp_head_conv_fwd / _dgrad / _wgrad (CUDA-core kernels of the 2-channel heads) against autograd."""
    g = torch.Generator().manual_seed(8)
    b, h, w = shape
    x = rnd(b, C, h, w, g=g).requires_grad_(True)
    wgt = rnd(2, C, 3, 3, g=g, scale=1.0 / (9 * C) ** 0.5).requires_grad_(True)
    bias = (torch.randn(2, generator=g).cuda() * 0.1).requires_grad_(True)
    z = F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), wgt, bias)
    y = torch.sigmoid(z) if sigmoid else z
    dz = rnd(*z.shape, g=g)
    z.backward(dz)
    x16 = nhwc(x.detach())
    out = torch.empty(b, h, w, 2, device='cuda')
    _lib.check(lib.fp_head_conv_fwd(P(x16), P(wgt.detach()), P(bias.detach()), b, h, w, C, sigmoid, BF, P(out), None))
    torch.cuda.synchronize()
    assert (out.permute(0, 3, 1, 2) - y.detach()).abs().max() <= 2e-5
    dz8 = torch.zeros(b, h, w, 8, dtype=DT, device='cuda')
    dz8[..., :2] = nhwc(dz)
    dw, db = torch.empty(2, C, 3, 3, device='cuda'), torch.empty(2, device='cuda')
    wsp = torch.empty(lib.fp_head_conv_wgrad_workspace_floats(C), device='cuda')
    _lib.check(lib.fp_head_conv_wgrad(P(x16), P(dz8), b, h, w, C, BF, BF, 0, P(dw), P(db), P(wsp), None))
    dx = torch.empty(b, h, w, C, dtype=DT, device='cuda')
    _lib.check(lib.fp_head_conv_dgrad(P(dz8), P(wgt.detach()), b, h, w, C, 0, None, 0, BF, BF, P(dx), None))
    torch.cuda.synchronize()
    assert (dw - wgt.grad).abs().max() <= 1e-4 * wgt.grad.abs().max() + 1e-5
    assert (db - bias.grad).abs().max() <= 1e-4 * bias.grad.abs().max() + 1e-5
    assert (nchw(dx) - x.grad).abs().max() <= 1e-2 * x.grad.abs().max()
