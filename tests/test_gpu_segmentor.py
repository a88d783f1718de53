"""the ground-segmentation network (footprints/preprocessing/segmentation/network.py:13-104, use_PSP=False)
on the same sm_100a kernels -- eval forward against the oracle restatement and the fixtures generated from the unmodified
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


def _net():
    net = Segmentor(pretrained=False)
    net.load_state_dict(synthetic_state_dict('resnet34', 0, spec=segmentor_state_spec('resnet34')))
    return net.cuda().eval()


@pytest.mark.parametrize('tag,shape', [('ref_segmentor_b2_64x96', (2, 64, 96)), ('ref_segmentor_b1_192x640', (1, 192, 640))])
def test_segmentor_eval_matches_reference_fixture_and_oracle(tag, shape):
    g = np.load(os.path.join(GOLD, f'{tag}.npz'))
    net = _net()
    data = synthetic_batch(*shape, 1)
    outs = net(data['image'].cuda())
    with torch.no_grad():
        ref = O.segmentor_forward(synthetic_state_dict('resnet34', 0, spec=segmentor_state_spec('resnet34')), data['image'])
    assert isinstance(outs, list) and len(outs) == 4
    b, h, w = shape
    for si, (v, r) in enumerate(zip(outs, ref)):
        assert v.shape == (b, 1, h >> (3 - si), w >> (3 - si)) and v.dtype == torch.float32
        err = (v.cpu() - r).abs().max().item()
        idx = torch.from_numpy(g[f'eval/{si}/idx']).long()
        err_fix = np.abs(v.cpu().flatten()[idx].numpy() - g[f'eval/{si}/val']).max()
        print(f'segmentor {shape} output {si}: max abs err vs oracle {err:.3e}, vs reference fixture {err_fix:.3e} (bar {TOL[si]:.0e})')
        assert err <= TOL[si] and err_fix <= TOL[si], (si, err, err_fix)


def test_segmentor_surface():
    net = _net()
    assert list(net.state_dict().keys()) == list(segmentor_state_spec('resnet34').keys())
    a = net(synthetic_batch(1, 64, 96, 2)['image'].cuda())
    b = net(synthetic_batch(1, 64, 96, 3)['image'].cuda())
    assert not torch.equal(a[3], b[3])                 # fresh tensors per call, not views of a reused buffer
    with pytest.raises(NotImplementedError):
        Segmentor(pretrained=False, use_PSP=True)
    with pytest.raises(NotImplementedError):
        net.train()(torch.rand(1, 3, 64, 96, device='cuda'))
    with pytest.raises(RuntimeError, match='multiples of 32'):
        net.eval()(torch.rand(1, 3, 60, 96, device='cuda'))
