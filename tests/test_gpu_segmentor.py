"""the ground-segmentation network (footprints/preprocessing/segmentation/network.py:13-207, with and without
the pyramid-pooling block) on the same sm_100a kernels -- eval forward against the oracle restatement and the fixtures generated from the unmodified
reference Segmentor."""
import os

import numpy as np
import pytest
import torch

from footprints_b200 import Segmentor
from footprints_b200.synthetic import segmentor_state_spec, synthetic_batch, synthetic_state_dict
from oracle import footprints_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
# max-abs bars of the four outputs (raw logits at 1/8, 1/4, 1/2 and full resolution), fp16 storage vs the fp32 reference
TOL = (4e-3, 2e-3, 1e-3, 1e-3)
# with the pyramid-pooling block the coarsest head sits behind a K = 9216 contraction of 1024 16-bit channels: measured
# 2.7e-3 (64x96) / 4.5e-3 (192x640) against 2.4e-3 / 3.1e-3 without it (profiles/r02E_pytest_new.log)
TOL_PSP = (6e-3, 2e-3, 1e-3, 1e-3)


def _net(use_PSP=False):
    net = Segmentor(pretrained=False, use_PSP=use_PSP)
    net.load_state_dict(synthetic_state_dict('resnet34', 0, spec=segmentor_state_spec('resnet34', use_PSP=use_PSP)))
    return net.cuda().eval()


@pytest.mark.parametrize('tag,shape', [('ref_segmentor_b2_64x96', (2, 64, 96)), ('ref_segmentor_b1_192x640', (1, 192, 640)),
                                       ('ref_segmentor_psp_b2_64x96', (2, 64, 96)), ('ref_segmentor_psp_b1_192x640', (1, 192, 640))])
def test_segmentor_eval_matches_reference_fixture_and_oracle(tag, shape):
    g = np.load(os.path.join(GOLD, f'{tag}.npz'))
    psp = '_psp_' in tag
    net = _net(psp)
    data = synthetic_batch(*shape, 1)
    outs = net(data['image'].cuda())
    with torch.no_grad():
        ref = O.segmentor_forward(synthetic_state_dict('resnet34', 0, spec=segmentor_state_spec('resnet34', use_PSP=psp)),
                                  data['image'])
    assert isinstance(outs, list) and len(outs) == 4
    b, h, w = shape
    tol = TOL_PSP if psp else TOL
    for si, (v, r) in enumerate(zip(outs, ref)):
        assert v.shape == (b, 1, h >> (3 - si), w >> (3 - si)) and v.dtype == torch.float32
        err = (v.cpu() - r).abs().max().item()
        idx = torch.from_numpy(g[f'eval/{si}/idx']).long()
        err_fix = np.abs(v.cpu().flatten()[idx].numpy() - g[f'eval/{si}/val']).max()
        print(f'segmentor {shape} output {si}: max abs err vs oracle {err:.3e}, vs reference fixture {err_fix:.3e} (bar {tol[si]:.0e})')
        assert err <= tol[si] and err_fix <= tol[si], (si, err, err_fix)


def test_segmentor_surface():
    net = _net()
    assert list(net.state_dict().keys()) == list(segmentor_state_spec('resnet34').keys())
    a = net(synthetic_batch(1, 64, 96, 2)['image'].cuda())
    b = net(synthetic_batch(1, 64, 96, 3)['image'].cuda())
    assert not torch.equal(a[3], b[3])                 # fresh tensors per call, not views of a reused buffer
    assert list(Segmentor(pretrained=False, use_PSP=True).state_dict().keys()) == \
        list(segmentor_state_spec('resnet34', use_PSP=True).keys())
    with pytest.raises(NotImplementedError):
        net.train()(torch.rand(1, 3, 64, 96, device='cuda'))
    with pytest.raises(RuntimeError, match='multiples of 32'):
        net.eval()(torch.rand(1, 3, 60, 96, device='cuda'))


@pytest.mark.parametrize('b,h,w,dtype', [(2, 6, 20, torch.float16), (3, 2, 3, torch.float16), (1, 16, 20, torch.bfloat16), (2, 1, 1, torch.float16)])
def test_pyramid_pooling_kernel_matches_torch(b, h, w, dtype):
    """fp_psp_fwd alone (segmentation/network.py:174-207) against the torch ops of the reference in float64, on the same
    16-bit input: adaptive pooling with cells larger AND smaller than one pixel, align_corners=True upsampling, concat order."""
    import ctypes
    from footprints_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(b * 100 + h)
    x = (torch.randn(b, h, w, 512, generator=g) * 2).to(dtype).cuda()
    ws = [(torch.randn(128, 512, generator=g) * 0.05).cuda() for _ in range(4)]
    out = torch.full((b, h, w, 1024), float('nan'), dtype=dtype, device='cuda')
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.fp_psp_fwd(x.data_ptr(), b, h, w, 0 if dtype == torch.float16 else 1, ws[0].data_ptr(), ws[1].data_ptr(),
                              ws[2].data_ptr(), ws[3].data_ptr(), out.data_ptr(), stream))
    sd = {f'decoder.PSP.block{i + 1}.reduce.weight': ws[i].double().cpu().view(128, 512, 1, 1) for i in range(4)}
    ref = O.psp_forward(sd, x.double().cpu().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    got = out.double().cpu()
    assert torch.equal(got[..., :512], x.double().cpu())                       # x passes through unchanged
    err = (got[..., 512:] - ref[..., 512:]).abs().max().item()
    # the kernel accumulates in fp32 and rounds once to the 16-bit output: half an ulp of the largest value (+ fp32 noise)
    bar = ref[..., 512:].abs().max().item() * (2.0 ** -11 if dtype == torch.float16 else 2.0 ** -8) * 1.02 + 1e-4
    print(f'psp {b}x{h}x{w} {dtype}: max abs err {err:.2e} (bar {bar:.2e} = half an ulp of the largest output)')
    assert err <= bar
