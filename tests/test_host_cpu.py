"""CPU-only checks of the host side: state_dict/contract parity with the reference layout, the C ABI
exports every symbol the header declares, and the no-fallback behaviour."""
import ctypes
import os
import re
import sys

import pytest
import torch

from footprints_b200 import FootprintNetwork, FootprintsModel, ModelManager, _lib
from footprints_b200.synthetic import state_spec, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('arch,blocks', [('resnet34', (3, 4, 6, 3)), ('resnet18', (2, 2, 2, 2))])
def test_state_dict_matches_reference_layout(arch, blocks):
    net = FootprintNetwork(pretrained=False, blocks=blocks)
    sd = net.state_dict()
    spec = state_spec(arch)
    assert list(sd.keys()) == list(spec.keys())
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    if arch == 'resnet34':
        assert len(sd) == 484
        n_grad = sum(p.numel() for p in net.parameters())
        assert n_grad == 31021392
    # a reference-format checkpoint loads strictly
    net.load_state_dict(synthetic_state_dict(arch, 0), strict=True)
    assert FootprintsModel is FootprintNetwork


def test_random_init_matches_reference_constructor():
    """Same seed -> same random-init weights as the reference constructor (needs /root/reference)."""
    if not os.path.isdir('/root/reference/footprints'):
        pytest.skip('reference checkout not present on this machine')
    import torchvision
    sys.path.insert(0, '/root/reference')
    import footprints.network as fn
    fn.resnet34 = lambda pretrained=True: torchvision.models.resnet34(weights=None)
    torch.manual_seed(0)
    ref = fn.FootprintNetwork(pretrained=False).state_dict()
    torch.manual_seed(0)
    ours = FootprintNetwork(pretrained=False).state_dict()
    assert list(ref.keys()) == list(ours.keys())
    for k in ref:
        assert torch.equal(ref[k], ours[k]), k


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'footprints_b200.h')).read()
    declared = set(re.findall(r'\b(fp_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert declared == set(_lib.PROTOTYPES.keys()), (declared ^ set(_lib.PROTOTYPES.keys()))
    _lib.load()
    assert lib.fp_version() == 100


def test_no_cpu_fallback():
    net = FootprintNetwork(pretrained=False).eval()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        net(torch.rand(1, 3, 64, 96))


def test_model_manager_surface(tmp_path):
    mm = ModelManager(save_folder=str(tmp_path), use_cuda=False, is_inference=False, pretrained=False)
    assert isinstance(mm.optimiser, torch.optim.Adam)
    assert isinstance(mm.scheduler, torch.optim.lr_scheduler.StepLR)
    mm.save_model('weights_0')
    assert os.path.exists(tmp_path / 'weights_0' / 'model.pth')
    mm2 = ModelManager(save_folder=str(tmp_path), use_cuda=False, is_inference=True, pretrained=False)
    mm2.load_model(str(tmp_path / 'weights_0'))
    for a, b in zip(mm.model.state_dict().values(), mm2.model.state_dict().values()):
        assert torch.equal(a, b)


def test_every_environment_switch_is_documented():
    """Each FP_* variable the library reads (csrc getenv calls) is listed in INTEGRATION.md."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in glob.glob(os.path.join(root, 'footprints_b200', 'csrc', '*.cu')) + glob.glob(os.path.join(root, 'footprints_b200', 'csrc', '*.cuh')):
        if os.path.basename(f) in ('test_harness.cu', 'probe_umma_shift.cu'):
            continue
        names |= set(re.findall(r'(?:getenv|env_int)\("(FP_[A-Z_]+)"', open(f).read()))
    doc = open(os.path.join(root, 'INTEGRATION.md')).read()
    missing = sorted(n for n in names if n not in doc)
    assert names and not missing, missing


def test_segmentor_state_dict_and_random_init_match_reference():
    """footprints_b200.Segmentor: same keys / shapes as the reference's segmentation network, and -- when the reference
    checkout is present -- the same random-init weights for a given torch seed."""
    from footprints_b200 import Segmentor
    from footprints_b200.synthetic import segmentor_state_spec
    net = Segmentor(pretrained=False)
    sd, spec = net.state_dict(), segmentor_state_spec('resnet34')
    assert list(sd.keys()) == list(spec.keys())
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    if not os.path.isdir('/root/reference/footprints'):
        return
    import torchvision
    sys.path.insert(0, '/root/reference')
    import footprints.preprocessing.segmentation.network as sn
    sn.resnet34 = lambda pretrained=True: torchvision.models.resnet34(weights=None)
    torch.manual_seed(0)
    ref = sn.Segmentor(pretrained=False).state_dict()
    torch.manual_seed(0)
    ours = Segmentor(pretrained=False).state_dict()
    assert list(ref.keys()) == list(ours.keys())
    for k in ref:
        assert torch.equal(ref[k], ours[k]), k


def test_gradient_buckets_partition_the_flat_buffer_in_backward_order():
    """Trainer's all-reduce buckets: decoders | encoder layer4 | layer3 | the rest, contiguous slices of the
    flat gradient that cover it exactly once, in the order the backward finishes them."""
    from footprints_b200.trainer import Trainer, trainable_parameters
    net = FootprintNetwork(pretrained=False).train()
    tr = Trainer(net)
    total = tr.grad.numel()
    (d0, d1), (a0, a1), (b0, b1) = tr._bucket[4], tr._bucket[3], tr._bucket[2]
    assert d1 == total and a1 == d0 and b1 == a0 and 0 < tr._rest_end == b0
    names = {id(p): n for n, p in net.named_parameters()}
    where = {}
    for p, o in zip(tr.params, tr._offs):
        where[names[id(p)]] = o
    for n, o in where.items():
        if n.startswith(('mask_decoder', 'depth_decoder')):
            assert d0 <= o < d1, n
        elif n.startswith('encoder.layer4'):
            assert a0 <= o < a1, n
        elif n.startswith('encoder.layer3'):
            assert b0 <= o < b1, n
        else:
            assert o < b0, n
    assert len(tr.params) == len(trainable_parameters(net)) == len([p for p in net.parameters()]) - 72   # unused decoder BNs


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm): one JSON line, CPU only, unmodified reference modules
    when baseline/_ref (or /root/reference) is importable."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--cpu-batch', '1',
                          '--hw', '64x96'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['kind'] in ('reference', 'port')
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
