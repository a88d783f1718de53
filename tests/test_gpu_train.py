"""GPU parity of the training step (train-mode forward, fused loss, hand-written backward) against the CPU
oracle (oracle/footprints_oracle.py) and the reference fixtures.  Activations and inter-layer gradients are
16-bit on our side (fp16 + loss scale by default, bf16 selectable), fp32 in the reference, so parameter gradients
are compared by direction (cosine) and norm."""
import os

import numpy as np
import pytest
import torch

from footprints_b200 import FootprintNetwork, LossManager
from footprints_b200.losses import fused_loss
from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict
from oracle import footprints_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
BLOCKS = {'resnet34': (3, 4, 6, 3), 'resnet18': (2, 2, 2, 2)}


# max-abs tolerance of the train-mode forward outputs vs the fp32 reference, per training storage type
TRAIN_TOL = {'fp16': {'1/8': 4e-3, '1/4': 2e-3, '1/2': 1e-3, '1/1': 1e-3},
             'bf16': {'1/8': 3e-2, '1/4': 2e-2, '1/2': 1e-2, '1/1': 1e-2}}
QDT = {'fp16': torch.float16, 'bf16': torch.bfloat16}


def _net(arch, train_dtype=None):
    net = FootprintNetwork(pretrained=False, blocks=BLOCKS[arch])
    if train_dtype is not None:
        net.train_dtype = train_dtype
    net.load_state_dict(synthetic_state_dict(arch, 0))
    return net.cuda().train()


def test_fused_loss_matches_oracle_and_autograd():
    g = torch.Generator().manual_seed(11)
    b, h, w = 2, 64, 96
    data = synthetic_batch(b, h, w, 3)
    preds = [torch.randn(b, 4, h, w, generator=g) for _ in range(4)]
    for p in preds:
        p[:, 2:4] = torch.sigmoid(p[:, 2:4])
    ref_in = {k: p.clone().requires_grad_(True) for k, p in zip(O.SCALES, preds)}
    ref = O.loss_forward(ref_in, data)
    ref['loss'].backward()
    tg = {k: v.cuda() for k, v in data.items() if k != 'image'}
    out, grads = fused_loss([p.cuda() for p in preds], tg, (0.1, 100.0), 0.25, True)
    out = out.cpu()
    terms = ('visible_ground', 'all_ground', 'depth', 'ground_depth', 'loss')
    for s, k in enumerate(O.SCALES):
        for c, t in enumerate(terms):
            assert abs(out[5 * s + c].item() - ref[(t, k)].item()) <= 2e-6 * max(1.0, abs(ref[(t, k)].item())), (t, k)
    assert abs(out[20].item() - ref['loss'].item()) <= 2e-6 * abs(ref['loss'].item())
    for s, k in enumerate(O.SCALES):
        gr = ref_in[k].grad
        assert (grads[s].cpu() - gr).abs().max() <= 1e-5 * gr.abs().max() + 1e-12


def test_loss_manager_contract_matches_reference_fixture():
    """Keys, order and values of LossManager on the reference's own train-mode outputs are pinned by the fixture;
    here: same keys/order, and values on our outputs close to the fixture's."""
    gold = np.load(os.path.join(GOLD, 'ref_resnet34_b2_64x96.npz'))
    net = _net('resnet34')
    data = synthetic_batch(2, 64, 96, 1)
    out = net(data['image'].cuda())
    targets = {k: v.cuda() for k, v in data.items() if k != 'image'}
    losses = LossManager((0.1, 100.0), 0.25)(out, targets)
    keys = list(losses.keys())
    expect = []
    for k in O.SCALES:
        expect += [('visible_ground', k), ('all_ground', k), ('depth', k), ('ground_depth', k), ('loss', k)]
    expect.append('loss')
    assert keys == expect
    for k, v in losses.items():
        name = 'loss/' + (k if isinstance(k, str) else f'{k[0]}@{k[1]}')
        assert abs(v.item() - float(gold[name])) <= 2e-2 * max(1.0, abs(float(gold[name]))), (name, v.item(), float(gold[name]))
    for key in (('visible_ground', '1/1'), ('all_ground', '1/1'), ('depth', '1/1'), ('ground_depth', '1/1'),
                ('ground_depth_masked', '1/1')):
        assert out[key].shape == (2, 64, 96)


@pytest.mark.parametrize('arch,train_dtype', [('resnet34', 'fp16'), ('resnet18', 'fp16'), ('resnet34', 'bf16')])
def test_train_step_matches_oracle(arch, train_dtype):
    """Forward, loss and BatchNorm running statistics against the fp32 oracle; parameter gradients by direction
    and norm.

    A ReLU/BatchNorm network's gradient is piecewise constant in its activations, so 16-bit storage rounding
    (which flips ~1 % of the ReLU masks per layer) bounds the attainable agreement: inside the CPU oracle itself,
    bf16-storage emulation vs fp32 gives cosines of only 0.84-0.95 for the deep encoder weights on this input, and
    two emulated runs whose inputs differ by 1e-6 agree to 0.89-0.95.  The per-parameter
    bar is therefore: at least as close to the storage-emulating oracle as that oracle is to fp32 (minus 0.03), and
    norms within 10 %.  The backward kernels themselves are pinned tightly, op by op, in test_gpu_train_kernels.py."""
    sd = synthetic_state_dict(arch, 0)
    data = synthetic_batch(2, 64, 96, 1)
    ref_out, ref_losses, g32, ref_stats = O.train_step_reference(sd, data, arch)
    _, _, gq, _ = O.train_step_reference(sd, data, arch, qdtype=QDT[train_dtype])
    net = _net(arch, train_dtype)
    out = net(data['image'].cuda())
    for k in O.SCALES:
        err = (out[k].detach().cpu() - ref_out[k].detach()).abs().max().item()
        print(f'{arch} {train_dtype} train fwd scale {k}: max abs err {err:.3e}')
        assert err <= TRAIN_TOL[train_dtype][k], (k, err)       # 16-bit activations, batch-stat BN
    targets = {k: v.cuda() for k, v in data.items() if k != 'image'}
    losses = LossManager((0.1, 100.0), 0.25, side_outputs=False)(out, targets)
    assert abs(losses['loss'].item() - ref_losses['loss'].item()) <= 1e-2 * ref_losses['loss'].item()
    net.zero_grad()
    losses['loss'].backward()
    named = dict(net.named_parameters())

    def cosine(a, r):
        return (torch.dot(a, r) / (a.norm() * r.norm() + 1e-30)).item()

    worst = (1.0, None)
    for name, gref in gq.items():
        gours = named[name].grad
        assert gours is not None, name
        a, r = gours.detach().float().cpu().flatten(), gref.flatten()
        assert torch.isfinite(a).all(), name
        cos = cosine(a, r)
        floor = min(0.995, cosine(r, g32[name].flatten())) - 0.03
        ratio = (a.norm() / (r.norm() + 1e-30)).item()
        worst = min(worst, (cos, name))
        assert cos >= floor, (name, cos, floor, ratio)
        assert 0.9 <= ratio <= 1.1, (name, cos, ratio)
        if name.startswith(('mask_decoder', 'depth_decoder')):
            assert cos >= 0.99, (name, cos)
    print(f'{arch} {train_dtype}: worst gradient cosine vs storage-emulating oracle {worst}')
    # parameters that get no gradient in the reference get none here either (unused decoder BNs)
    for name, p in named.items():
        if name not in gq:
            assert p.grad is None, name
    # running statistics were updated like torch BatchNorm2d(momentum=0.1)
    new_sd = net.state_dict()
    for k, v in ref_stats.items():
        # first layer: only bf16 input/weight rounding separates us from fp32; deep layers see 12-48 samples per
        # channel at this test size, so their batch variance inherits the accumulated bf16 forward error
        tight = k.startswith('encoder.layer0.')
        assert torch.allclose(new_sd[k].cpu(), v, rtol=1e-2 if tight else 0.15, atol=1e-2 if tight else 0.05), k
    assert int(new_sd['encoder.layer0.1.num_batches_tracked']) == 1


def test_train_step_reference_fixture_gradient_norms():
    gold = np.load(os.path.join(GOLD, 'ref_resnet34_b2_64x96.npz'))
    net = _net('resnet34')
    data = synthetic_batch(2, 64, 96, 1)
    out = net(data['image'].cuda())
    targets = {k: v.cuda() for k, v in data.items() if k != 'image'}
    LossManager((0.1, 100.0), 0.25, side_outputs=False)(out, targets)['loss'].backward()
    named = dict(net.named_parameters())
    # against the reference's own fp32 gradients: norms within 10 %; element-wise the bf16 forward flips ~1 % of
    # the ReLU masks per layer, so deep encoder gradients agree in direction only to cos ~0.85
    for name, ref_norm in zip(gold['grad_names'], gold['grad_norms']):
        n = named[str(name)].grad.double().norm().item()
        assert abs(n - ref_norm) <= 0.1 * ref_norm + 1e-12, (str(name), n, ref_norm)
        if str(name).startswith(('mask_decoder', 'depth_decoder')):
            idx = torch.from_numpy(gold[f'grad/{name}/idx']).long()
            ref = torch.from_numpy(gold[f'grad/{name}/val'])
            got = named[str(name)].grad.flatten().cpu()[idx]
            assert (got - ref).abs().max() <= 0.1 * ref.abs().max() + 1e-9, str(name)


def test_optimizer_step_through_reference_training_loop_surface():
    """model.zero_grad(); losses['loss'].backward(); optimiser.step() as in footprints/training/train.py:153-156."""
    from footprints_b200 import ModelManager
    mm = ModelManager(use_cuda=True, is_inference=False, pretrained=False, learning_rate=1e-4)
    mm.model.train()
    data = synthetic_batch(2, 64, 96, 4)
    targets = {k: v.cuda() for k, v in data.items() if k != 'image'}
    lm = LossManager((0.1, 100.0), 0.25)
    first = None
    for _ in range(3):
        out = mm.model(data['image'].cuda())
        losses = lm(out, targets)
        mm.model.zero_grad()
        losses['loss'].backward()
        mm.optimiser.step()
        first = first if first is not None else losses['loss'].item()
    assert losses['loss'].item() < first          # three Adam steps on one batch reduce its loss


def test_trainer_fast_path_matches_autograd_path_and_learns():
    """Trainer.step (no autograd, flat buffers, fused Adam) produces the same first-step loss as the autograd drop-in
    path and the same parameter update as torch.optim.Adam on the autograd gradients; loss decreases over steps."""
    from footprints_b200.trainer import Trainer
    data = synthetic_batch(2, 64, 96, 9)
    img = data['image'].cuda()
    targets = {k: v.cuda() for k, v in data.items() if k != 'image'}
    net_a, net_b = _net('resnet34'), _net('resnet34')
    # autograd path + torch Adam
    opt = torch.optim.Adam(net_a.parameters(), lr=1e-4)
    losses_a = LossManager((0.1, 100.0), 0.25, side_outputs=False)(net_a(img), targets)
    opt.zero_grad()
    losses_a['loss'].backward()
    opt.step()
    # fast path
    tr = Trainer(net_b, lr=1e-4)
    l0 = tr.step(img, targets)
    assert abs(l0[20].item() - losses_a['loss'].item()) <= 1e-5 * abs(losses_a['loss'].item())
    sd_a, sd_b = net_a.state_dict(), net_b.state_dict()
    for k in sd_a:
        if sd_a[k].is_floating_point():
            assert torch.allclose(sd_a[k], sd_b[k], rtol=1e-5, atol=2e-7), k
    last = l0[20].item()
    for _ in range(4):
        last = tr.step(img, targets)[20].item()
    assert last < l0[20].item()
