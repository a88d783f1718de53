"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
in this container.  *** TEST INFRASTRUCTURE ONLY *** — run by hand:  python oracle/gen_golden.py

The reference cannot travel to the GPU box, so its outputs on deterministic inputs/weights
(``footprints_b200.synthetic``) are committed as small fixtures: sampled output values, sums, the
21-entry loss dict, per-parameter gradient norms and sampled gradient values, and updated BN
running statistics.  In-process shims only: ``footprints.network.resnet34`` is
replaced by a no-download constructor; nothing under /root/reference is edited.
"""
import os
import sys

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict, state_spec  # noqa: E402

N_SAMPLES = 512


def sample_idx(numel, seed, n=N_SAMPLES):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(n, numel),), generator=g, dtype=torch.int32)


def build_reference(arch):
    import footprints.network as fn
    ctor = getattr(torchvision.models, arch)
    fn.resnet34 = lambda pretrained=True: ctor(weights=None)
    from footprints.training.losses import LossManager
    return fn.FootprintNetwork(pretrained=False), LossManager((0.1, 100.0), 0.25)


def run(arch, batch, height, width, tag):
    torch.manual_seed(0)
    torch.set_num_threads(8)
    net, loss_manager = build_reference(arch)
    sd = synthetic_state_dict(arch, seed=0)
    ref_sd = net.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), 'state_dict key order differs from reference'
    for k in sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    assert list(state_spec(arch).keys()) == list(ref_sd.keys())
    net.load_state_dict(sd)
    data = synthetic_batch(batch, height, width, seed=1)
    out = {}

    # ---- eval forward (configs 1/2) ----
    net.eval()
    with torch.no_grad():
        pred = net(data['image'])
    assert list(pred.keys()) == ['1/8', '1/4', '1/2', '1/1']
    for si, (k, v) in enumerate(pred.items()):
        idx = sample_idx(v.numel(), 100 + si).long()
        out[f'eval/{k}/idx'] = idx.numpy().astype(np.int32)
        out[f'eval/{k}/val'] = v.flatten()[idx].numpy()
        out[f'eval/{k}/sum'] = np.float64(v.double().sum().item())
        out[f'eval/{k}/abssum'] = np.float64(v.double().abs().sum().item())

    # ---- train step: fwd (batch-stat BN) + LossManager + backward (configs 3-5) ----
    net.train()
    net.zero_grad()
    pred = net(data['image'])
    for si, k in enumerate(['1/8', '1/4', '1/2', '1/1']):
        v = pred[k].detach()
        idx = sample_idx(v.numel(), 200 + si).long()
        out[f'train/{k}/idx'] = idx.numpy().astype(np.int32)
        out[f'train/{k}/val'] = v.flatten()[idx].numpy()
        out[f'train/{k}/sum'] = np.float64(v.double().sum().item())
    targets = {k: v for k, v in data.items() if k != 'image'}
    losses = loss_manager(pred, targets)
    for k, v in losses.items():
        name = 'loss/' + (k if isinstance(k, str) else f'{k[0]}@{k[1]}')
        out[name] = np.float64(v.item())
    # side outputs LossManager adds to predictions (losses.py:56-78,90)
    for key in (('visible_ground', '1/1'), ('all_ground', '1/1'), ('depth', '1/1'),
                ('ground_depth', '1/1'), ('ground_depth_masked', '1/1')):
        v = pred[key].detach()
        out[f'side/{key[0]}/sum'] = np.float64(v.double().sum().item())
    losses['loss'].backward()
    names, norms = [], []
    for pi, (k, p) in enumerate(net.named_parameters()):
        if p.grad is None:
            continue
        names.append(k)
        norms.append(p.grad.double().norm().item())
        idx = sample_idx(p.grad.numel(), 1000 + pi, 48).long()
        out[f'grad/{k}/idx'] = idx.numpy().astype(np.int32)
        out[f'grad/{k}/val'] = p.grad.flatten()[idx].numpy()
    out['grad_names'] = np.array(names)
    out['grad_norms'] = np.array(norms, dtype=np.float64)
    new_sd = net.state_dict()
    for k in ('encoder.layer0.1.running_mean', 'encoder.layer0.1.running_var',
              'encoder.layer4.1.bn2.running_mean', 'encoder.layer4.1.bn2.running_var',
              'encoder.layer2.0.downsample.1.running_var'):
        out[f'bnstat/{k}'] = new_sd[k].numpy()
    out['meta'] = np.array([batch, height, width])
    path = os.path.join(ROOT, 'tests', 'golden', f'{tag}.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes;', len(names), 'params with grad; loss',
          out['loss/loss'])


def run_segmentor(batch, height, width, tag, use_PSP=False):
    """Eval outputs of the reference's ground-segmentation network (preprocessing/segmentation/network.py)."""
    from footprints_b200.synthetic import segmentor_state_spec
    import footprints.preprocessing.segmentation.network as sn
    sn.resnet34 = lambda pretrained=True: torchvision.models.resnet34(weights=None)
    torch.manual_seed(0)
    net = sn.Segmentor(pretrained=False, use_PSP=use_PSP)
    spec = segmentor_state_spec('resnet34', use_PSP=use_PSP)
    ref_sd = net.state_dict()
    assert list(ref_sd.keys()) == list(spec.keys()), 'segmentor state_dict key order differs from reference'
    for k in spec:
        assert tuple(ref_sd[k].shape) == tuple(spec[k]), k
    net.load_state_dict(synthetic_state_dict('resnet34', seed=0, spec=spec))
    net.eval()
    data = synthetic_batch(batch, height, width, seed=1)
    with torch.no_grad():
        outs = net(data['image'])
    out = {}
    for si, v in enumerate(outs):
        idx = sample_idx(v.numel(), 300 + si).long()
        out[f'eval/{si}/shape'] = np.array(v.shape)
        out[f'eval/{si}/idx'] = idx.numpy().astype(np.int32)
        out[f'eval/{si}/val'] = v.flatten()[idx].numpy()
        out[f'eval/{si}/sum'] = np.float64(v.double().sum().item())
    path = os.path.join(ROOT, 'tests', 'golden', f'{tag}.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'segmentor':
        run_segmentor(2, 64, 96, 'ref_segmentor_b2_64x96')
        run_segmentor(1, 192, 640, 'ref_segmentor_b1_192x640')
        run_segmentor(2, 64, 96, 'ref_segmentor_psp_b2_64x96', use_PSP=True)
        run_segmentor(1, 192, 640, 'ref_segmentor_psp_b1_192x640', use_PSP=True)
        sys.exit(0)
    run('resnet34', 2, 64, 96, 'ref_resnet34_b2_64x96')
    run('resnet18', 2, 64, 96, 'ref_resnet18_b2_64x96')
    run('resnet34', 1, 192, 640, 'ref_resnet34_b1_192x640')
