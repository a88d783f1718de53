"""Autograd / checkpoint / cache semantics the reference's callers rely on: saved state lives on the
autograd node, the fast-path Trainer checkpoints like torch.optim.Adam, eval sees raw-pointer weight updates."""
import os

import pytest
import torch

from footprints_b200 import FootprintNetwork, LossManager, ModelManager
from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict
from footprints_b200.trainer import Trainer

pytestmark = pytest.mark.gpu


def _net():
    net = FootprintNetwork(pretrained=False)
    net.load_state_dict(synthetic_state_dict('resnet34', 0))
    return net.cuda().train()


def _targets(d):
    return {k: v.cuda() for k, v in d.items() if k != 'image'}


def _grads(net):
    return {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}


def test_two_forwards_then_two_backwards_and_accumulation():
    lm = LossManager((0.1, 100.0), 0.25, side_outputs=False)
    d1, d2 = synthetic_batch(2, 64, 96, 1), synthetic_batch(2, 64, 96, 2)
    # reference order: forward(x1), backward; forward(x2), backward -- each alone
    net = _net()
    lm(net(d1['image'].cuda()), _targets(d1))['loss'].backward()
    g1 = _grads(net)
    net = _net()
    lm(net(d2['image'].cuda()), _targets(d2))['loss'].backward()
    g2 = _grads(net)
    # both forwards first, then the first one's backward: must use x1's activations and batch statistics
    net = _net()
    o1 = net(d1['image'].cuda())
    o2 = net(d2['image'].cuda())
    l1, l2 = lm(o1, _targets(d1))['loss'], lm(o2, _targets(d2))['loss']
    l1.backward()
    for n, g in _grads(net).items():
        assert torch.equal(g, g1[n]), n
    # gradient accumulation over the two nodes
    l2.backward()
    for n, g in _grads(net).items():
        ref = g1[n] + g2[n]
        assert (g - ref).abs().max() <= 1e-6 * ref.abs().max() + 1e-12, n
    # a second backward through a consumed node is an error, not silent garbage
    with pytest.raises(RuntimeError):
        lm(o1, _targets(d1))['loss'].backward()


def test_train_mode_forward_under_no_grad_saves_nothing():
    net = _net()
    d = synthetic_batch(4, 192, 640, 3)       # the saved activation set of this forward would be ~0.8 GB
    img = d['image'].cuda()
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        out = net(img)
    assert not out['1/1'].requires_grad
    torch.cuda.synchronize()
    held = torch.cuda.memory_allocated() - base - sum(v.numel() * 4 for v in out.values())
    assert held < 300 << 20, held          # packed weights (2 x 62 MB) and workspaces only; no activation set pinned


def test_eval_after_trainer_step_sees_the_updated_weights():
    net = _net()
    d = synthetic_batch(2, 64, 96, 4)
    img = d['image'].cuda()
    net.eval()
    before = net(img)['1/1'].clone()
    net.train()
    tr = Trainer(net, lr=1e-3)
    tr.step(img, _targets(d))
    net.eval()
    after = net(img)['1/1']
    assert (after - before).abs().max() > 1e-4
    fresh = FootprintNetwork(pretrained=False).cuda().eval()
    fresh.load_state_dict(net.state_dict())
    assert torch.equal(fresh(img)['1/1'], after)


def test_trainer_checkpoint_resumes_exactly_and_is_torch_adam_format(tmp_path):
    d = [synthetic_batch(2, 64, 96, 10 + i) for i in range(4)]

    def run(n_first, resume):
        net = _net()
        mm = ModelManager(save_folder=str(tmp_path), use_cuda=True, pretrained=False, learning_rate=1e-4, lr_step_size=1)
        mm.model.load_state_dict(net.state_dict())
        mm.model.train()
        tr = mm.attach_trainer(Trainer(mm.model, optimiser=mm.optimiser))
        for i in range(n_first):
            tr.step(d[i]['image'].cuda(), _targets(d[i]))
        if resume:
            mm.scheduler.step()                            # StepLR drives the fast path through the attached optimiser
            mm.save_model('ckpt')
            mm = ModelManager(save_folder=str(tmp_path), use_cuda=True, pretrained=False, learning_rate=1e-4, lr_step_size=1)
            mm.model.train()
            tr = mm.attach_trainer(Trainer(mm.model, optimiser=mm.optimiser))
            mm.load_model(os.path.join(str(tmp_path), 'ckpt'), load_optimiser=True)
            assert tr.step_count == n_first and abs(tr.lr - 1e-5) < 1e-12
        else:
            mm.scheduler.step()
        for i in range(n_first, 4):
            tr.step(d[i]['image'].cuda(), _targets(d[i]))
        return mm

    a, b = run(2, False), run(2, True)
    for (k, va), vb in zip(a.model.state_dict().items(), b.model.state_dict().values()):
        assert torch.equal(va, vb), k
    # optimiser.pth written through the Trainer loads into a plain torch.optim.Adam over the same parameters
    sd = torch.load(os.path.join(str(tmp_path), 'ckpt', 'optimiser.pth'))
    plain = torch.optim.Adam(FootprintNetwork(pretrained=False).parameters(), lr=1.0)
    plain.load_state_dict(sd)
    assert abs(plain.param_groups[0]['lr'] - 1e-5) < 1e-12
    some = next(iter(plain.state.values()))
    assert float(some['step']) == 2.0 and some['exp_avg'].abs().max() > 0


def test_trainer_cuda_graph_replays_match_eager_steps():
    """Trainer.enable_cuda_graph(): steps 1-2 run eagerly, step 3 is captured, later steps are replays with fresh inputs
    and refreshed Adam scalars -- parameters must follow the eager trainer's trajectory."""
    batches = [synthetic_batch(2, 64, 96, 30 + i) for i in range(6)]
    nets = [_net(), _net()]
    eager = Trainer(nets[0], lr=1e-4)
    graph = Trainer(nets[1], lr=1e-4).enable_cuda_graph()
    for i, d in enumerate(batches):
        img, tg = d['image'].cuda(), _targets(d)
        la = eager.step(img, tg).clone()
        lb = graph.step(img, tg).clone()
        assert torch.equal(la, lb), (i, (la - lb).abs().max().item(), la[20].item(), lb[20].item())
    assert graph._graph is not None and graph.step_count == eager.step_count == 6
    for (k, a), b in zip(nets[0].state_dict().items(), nets[1].state_dict().values()):
        assert torch.equal(a, b), k
    # a new input shape falls back to eager steps and re-captures
    d = synthetic_batch(1, 64, 128, 40)
    for _ in range(4):
        graph.step(d['image'].cuda(), _targets(d))
    assert graph._graph is not None and graph._graph_key[0] == (1, 3, 64, 128)


def test_fp16_gradient_overflow_saturates_instead_of_nan():
    """A loss scale 2^20 too large: the 16-bit conversions saturate (cvt.rn.satfinite), so the parameter gradients come
    out finite (clipped), never inf / NaN."""
    net = _net()
    d = synthetic_batch(2, 64, 96, 6)
    net.train_engine().loss_scale = 2.0 ** 44
    LossManager((0.1, 100.0), 0.25, side_outputs=False)(net(d['image'].cuda()), _targets(d))['loss'].backward()
    for n, p in net.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), n
