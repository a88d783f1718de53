"""Two training steps at bs 64 (for ncu launch lists: `ncu -s <first step's launches> ...`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from footprints_b200 import FootprintNetwork
from footprints_b200.synthetic import synthetic_batch
from footprints_b200.trainer import Trainer
torch.manual_seed(0)
net = FootprintNetwork(pretrained=False).cuda().train()
tr = Trainer(net)
data = {k: v.cuda() for k, v in synthetic_batch(64, 192, 640, 1).items()}
tg = {k: v for k, v in data.items() if k != 'image'}
for _ in range(2):
    tr.step(data['image'], tg)
torch.cuda.synchronize()
print('launches per step', tr.eng.launches // 2)
