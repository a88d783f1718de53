"""Host enqueue time vs device time of one training step (is the step launch-bound?)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from footprints_b200 import FootprintNetwork
from footprints_b200.synthetic import synthetic_batch
from footprints_b200.trainer import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = FootprintNetwork(pretrained=False).cuda().train()
tr = Trainer(net)
data = {k: v.cuda() for k, v in synthetic_batch(B, 192, 640, 1).items()}
tg = {k: v for k, v in data.items() if k != 'image'}
for _ in range(3):
    tr.step(data['image'], tg)
torch.cuda.synchronize()
for it in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    l0 = tr.eng.launches
    t0 = time.perf_counter(); e0.record()
    tr.step(data['image'], tg)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f'batch {B}: host enqueue {1e3 * (t1 - t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms, launches {tr.eng.launches - l0}')
