"""Per-launch CUDA-event times of the convolution kernels in one training step (bs 64, 192x640)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from footprints_b200 import FootprintNetwork
from footprints_b200.synthetic import synthetic_batch
from footprints_b200.trainer import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
net = FootprintNetwork(pretrained=False).cuda().train()
tr = Trainer(net)
data = {k: v.cuda() for k, v in synthetic_batch(B, 192, 640, 1).items()}
tg = {k: v for k, v in data.items() if k != 'image'}
for _ in range(3):
    tr.step(data['image'], tg)
tr.eng.profile = []
tr.eng.overlap_wgrad = False      # serialise the launches: per-kernel CUDA-event times
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); tr.step(data['image'], tg); e1.record(); torch.cuda.synchronize()
prof = tr.eng.profile; tr.eng.profile = None
print('step ms (with event overhead)', e0.elapsed_time(e1))
tot = {}
rows = []
for kind, name, flops, a, b, _ in prof:
    ms = a.elapsed_time(b)
    rows.append((ms, kind, name, flops / ms / 1e9))
    tot[kind] = tot.get(kind, 0) + ms
print(tot)
top = int(os.environ.get('TOP', '45'))
for ms, kind, name, tf in sorted(rows, reverse=True)[:top]:
    print(f'{ms:8.3f} ms  {kind:6s} {name:28s} {tf:8.1f} TFLOP/s')
