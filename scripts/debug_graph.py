import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from footprints_b200 import FootprintNetwork
from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict
from footprints_b200.trainer import Trainer

def mk():
    net = FootprintNetwork(pretrained=False)
    net.load_state_dict(synthetic_state_dict('resnet34', 0))
    return net.cuda().train()

batches = [synthetic_batch(2, 64, 96, 30 + i) for i in range(6)]
nets = [mk(), mk()]
eager = Trainer(nets[0], lr=1e-3)
graph = Trainer(nets[1], lr=1e-3).enable_cuda_graph()
for i, d in enumerate(batches):
    img = d['image'].cuda(); tg = {k: v.cuda() for k, v in d.items() if k != 'image'}
    la = eager.step(img, tg).clone(); lb = graph.step(img, tg).clone()
    torch.cuda.synchronize()
    gd = (eager.grad - graph.grad).abs()
    pd = (eager.flat - graph.flat).abs()
    md = (eager.exp_avg - graph.exp_avg).abs()
    print(f'step {i}: graph={graph._graph is not None} loss diff {(la-lb).abs().max().item():.3e} grad diff max {gd.max().item():.3e} (n={int((gd>0).sum())}) '
          f'param diff {pd.max().item():.3e} (n={int((pd>0).sum())}) m diff {md.max().item():.3e} hyper {eager._hyper_dev.tolist()[:7]} | {graph._hyper_dev.tolist()[:7]}')
    if gd.max() > 0:
        offs = eager._offs
        bad = []
        for p, o, (name, _) in zip(eager.params, offs, [(n, q) for n, q in nets[0].named_parameters() if any(q is pp for pp in eager.params)]):
            n = p.numel()
            dd = gd[o:o+n].max().item()
            if dd > 0: bad.append((name, dd, eager.grad[o:o+n].abs().max().item()))
        print('   differing grads:', bad[:12], len(bad))
