"""The CPU oracle (oracle/footprints_oracle.py) must reproduce the committed outputs of the real
reference (tests/golden/*.npz, written by oracle/gen_golden.py from /root/reference)."""
import os

import numpy as np
import pytest
import torch

from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict, state_spec
from oracle import footprints_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
CASES = [('resnet34', 'ref_resnet34_b2_64x96'), ('resnet18', 'ref_resnet18_b2_64x96'),
         ('resnet34', 'ref_resnet34_b1_192x640')]


def _load(tag):
    return np.load(os.path.join(GOLD, tag + '.npz'))


def test_state_spec_is_the_reference_484_keys():
    spec = state_spec('resnet34')
    assert len(spec) == 484
    assert spec['encoder.layer0.0.weight'] == (64, 3, 7, 7)
    assert spec['mask_decoder.outconv4.0.conv1.weight'] == (32, 64, 3, 3)
    assert spec['mask_decoder.outconv4.1.conv1.weight'] == (2, 32, 3, 3)
    n_params = sum(int(np.prod(s)) for k, s in spec.items()
                   if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked')))
    assert n_params == 31021392


@pytest.mark.parametrize('arch,tag', CASES[:2])
def test_oracle_eval_forward_matches_reference(arch, tag):
    g = _load(tag)
    b, h, w = (int(v) for v in g['meta'])
    sd = synthetic_state_dict(arch, 0)
    data = synthetic_batch(b, h, w, 1)
    with torch.no_grad():
        out = O.network_forward(sd, data['image'], arch, training=False)
    assert list(out.keys()) == ['1/8', '1/4', '1/2', '1/1']
    for k, v in out.items():
        assert v.shape == (b, 4, h, w)
        idx = torch.from_numpy(g[f'eval/{k}/idx']).long()
        np.testing.assert_allclose(v.flatten()[idx].numpy(), g[f'eval/{k}/val'], rtol=1e-4, atol=2e-5)
        assert abs(v.double().sum().item() - float(g[f'eval/{k}/sum'])) <= 1e-5 * float(g[f'eval/{k}/abssum'])


@pytest.mark.parametrize('arch,tag', CASES[:2])
def test_oracle_train_step_matches_reference(arch, tag):
    g = _load(tag)
    b, h, w = (int(v) for v in g['meta'])
    sd = synthetic_state_dict(arch, 0)
    data = synthetic_batch(b, h, w, 1)
    outs, losses, grads, new_stats = O.train_step_reference(sd, data, arch)
    for k in O.SCALES:
        idx = torch.from_numpy(g[f'train/{k}/idx']).long()
        np.testing.assert_allclose(outs[k].detach().flatten()[idx].numpy(), g[f'train/{k}/val'],
                                   rtol=1e-4, atol=2e-5)
    assert len(losses) == 21
    for k, v in losses.items():
        name = 'loss/' + (k if isinstance(k, str) else f'{k[0]}@{k[1]}')
        assert abs(v.item() - float(g[name])) <= 2e-6 * max(1.0, abs(float(g[name]))), name
    names = [str(n) for n in g['grad_names']]
    assert list(grads.keys()) == names
    for n, ref_norm in zip(names, g['grad_norms']):
        gn = grads[n].double().norm().item()
        assert abs(gn - ref_norm) <= 1e-3 * ref_norm + 1e-12, n
        idx = torch.from_numpy(g[f'grad/{n}/idx']).long()
        ref = g[f'grad/{n}/val']
        np.testing.assert_allclose(grads[n].flatten()[idx].numpy(), ref, rtol=2e-3,
                                   atol=1e-3 * float(np.abs(ref).max()) + 1e-12)
    for k in [k for k in g.files if k.startswith('bnstat/')]:
        np.testing.assert_allclose(new_stats[k[7:]].numpy(), g[k], rtol=1e-5, atol=1e-6)


def test_oracle_full_resolution_eval_matches_reference():
    arch, tag = CASES[2]
    g = _load(tag)
    b, h, w = (int(v) for v in g['meta'])
    sd = synthetic_state_dict(arch, 0)
    data = synthetic_batch(b, h, w, 1)
    with torch.no_grad():
        out = O.network_forward(sd, data['image'], arch, training=False)
    for k, v in out.items():
        idx = torch.from_numpy(g[f'eval/{k}/idx']).long()
        np.testing.assert_allclose(v.flatten()[idx].numpy(), g[f'eval/{k}/val'], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize('tag,shape', [('ref_segmentor_b2_64x96', (2, 64, 96)), ('ref_segmentor_b1_192x640', (1, 192, 640)),
                                       ('ref_segmentor_psp_b2_64x96', (2, 64, 96)), ('ref_segmentor_psp_b1_192x640', (1, 192, 640))])
def test_segmentor_oracle_matches_reference_fixture(tag, shape):
    """oracle.segmentor_forward against outputs of the unmodified reference Segmentor (oracle/gen_golden.py segmentor)."""
    from footprints_b200.synthetic import segmentor_state_spec
    g = np.load(os.path.join(GOLD, f'{tag}.npz'))
    sd = synthetic_state_dict('resnet34', 0, spec=segmentor_state_spec('resnet34', use_PSP='_psp_' in tag))
    with torch.no_grad():
        outs = O.segmentor_forward(sd, synthetic_batch(*shape, 1)['image'])
    assert len(outs) == 4
    for si, v in enumerate(outs):
        assert tuple(v.shape) == tuple(g[f'eval/{si}/shape'])
        idx = torch.from_numpy(g[f'eval/{si}/idx']).long()
        ref = torch.from_numpy(g[f'eval/{si}/val'])
        assert torch.allclose(v.flatten()[idx], ref, rtol=1e-4, atol=1e-5), si
        assert abs(v.double().sum().item() - float(g[f'eval/{si}/sum'])) <= 1e-4 * max(1.0, abs(float(g[f'eval/{si}/sum'])))
