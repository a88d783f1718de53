"""Per-parameter gradient agreement (cosine, norm ratio) between the CUDA training step and the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from footprints_b200 import FootprintNetwork, LossManager
from footprints_b200.synthetic import synthetic_batch, synthetic_state_dict
from oracle import footprints_oracle as O

arch = sys.argv[1] if len(sys.argv) > 1 else 'resnet34'
dtype = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
blocks = {'resnet34': (3, 4, 6, 3), 'resnet18': (2, 2, 2, 2)}[arch]
sd = synthetic_state_dict(arch, 0)
B, H, W = (int(v) for v in (sys.argv[3:6] if len(sys.argv) > 5 else (2, 64, 96)))
data = synthetic_batch(B, H, W, 1)
qd = torch.bfloat16 if os.environ.get('ORACLE_Q') else None
ref_out, ref_losses, ref_grads, ref_stats = O.train_step_reference(sd, data, arch, qdtype=qd)
net = FootprintNetwork(pretrained=False, blocks=blocks)
net.train_dtype = dtype
net.load_state_dict(sd)
net = net.cuda().train()
out = net(data['image'].cuda())
targets = {k: v.cuda() for k, v in data.items() if k != 'image'}
losses = LossManager((0.1, 100.0), 0.25, side_outputs=False)(out, targets)
losses['loss'].backward()
named = dict(net.named_parameters())
for name, gref in ref_grads.items():
    a, r = named[name].grad.float().cpu().flatten(), gref.flatten()
    cos = (torch.dot(a, r) / (a.norm() * r.norm() + 1e-30)).item()
    print(f'{name:60s} cos {cos:.4f} ratio {(a.norm() / r.norm()).item():.4f} |g| {r.norm().item():.3e}')
