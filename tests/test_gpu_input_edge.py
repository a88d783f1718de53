"""Input edge on the B200: ResizeToTensor (fp_resize_u8_horizontal + fp_resize_u8_vertical_to_f32) must return, BIT FOR BIT,
the tensor the reference builds on the host with transforms.Resize((H, W), Image.ANTIALIAS) + transforms.ToTensor()
(predict_simple.py:41-43,51-60; datasets/inference_dataset.py:25-27,45-48).  Checked against Pillow itself (the reference's
dependency, present on the box) and against the oracle's restatement."""
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pil_resample as oracle  # noqa: E402

pytestmark = pytest.mark.gpu


def _pillow(frame, ho, wo):
    """resizer + totensor of the reference on one uint8 HWC frame."""
    img = Image.fromarray(frame).resize((wo, ho), Image.LANCZOS)
    return torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).contiguous().to(torch.float32).div(255)


@pytest.mark.parametrize('hin,win,ho,wo', [(375, 1242, 192, 640), (370, 1226, 192, 640), (480, 640, 512, 640),
                                           (100, 200, 192, 640), (192, 640, 192, 640), (1080, 1920, 192, 640),
                                           (37, 53, 32, 64), (33, 1001, 64, 96)])
def test_resize_to_tensor_is_bit_exact_with_pillow(hin, win, ho, wo):
    from footprints_b200.preprocess import ResizeToTensor
    rng = np.random.default_rng(hin + win)
    frames = rng.integers(0, 256, (3, hin, win, 3), dtype=np.uint8)
    out = ResizeToTensor((ho, wo))(frames)
    assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (3, 3, ho, wo)
    out = out.cpu()
    for i in range(3):
        assert torch.equal(out[i], _pillow(frames[i], ho, wo)), f'frame {i}'


def test_resize_matches_the_oracle_and_saturates_like_pillow():
    from footprints_b200.preprocess import ResizeToTensor
    y, x = np.mgrid[0:375, 0:1242]
    checker = (((x // 7 + y // 5) % 2) * 255).astype(np.uint8)
    frame = np.stack([checker, 255 - checker, checker], -1)
    out = ResizeToTensor((192, 640))(frame)
    assert tuple(out.shape) == (3, 192, 640)
    assert np.array_equal(out.cpu().numpy(), oracle.resize_to_tensor(frame, (192, 640)))
    assert torch.equal(out.cpu(), _pillow(frame, 192, 640))


def test_accepts_pil_numpy_host_and_device_tensors_and_unaligned_views():
    from footprints_b200.preprocess import ResizeToTensor
    rng = np.random.default_rng(5)
    frame = rng.integers(0, 256, (120, 333, 3), dtype=np.uint8)         # 999-byte rows: every row start is misaligned
    rt = ResizeToTensor((64, 96))
    ref = _pillow(frame, 64, 96)
    assert torch.equal(rt(Image.fromarray(frame)).cpu(), ref)
    assert torch.equal(rt(frame).cpu(), ref)
    assert torch.equal(rt(torch.from_numpy(frame)).cpu(), ref)
    assert torch.equal(rt(torch.from_numpy(frame).cuda()).cpu(), ref)
    # a device buffer whose first byte is not 16-byte aligned (the guarded head / tail vectors of the row loader)
    raw = torch.zeros(frame.size + 7, dtype=torch.uint8, device='cuda')
    raw[5:5 + frame.size] = torch.from_numpy(frame).cuda().flatten()
    view = raw[5:5 + frame.size].view(120, 333, 3)
    assert view.data_ptr() % 16 != 0
    assert torch.equal(rt(view).cpu(), ref)
    grey = Image.fromarray(frame[:, :, 0], mode='L')                     # pil_loader converts to RGB (utils.py:24-29)
    assert torch.equal(rt(grey).cpu(), _pillow(np.asarray(grey.convert('RGB')), 64, 96))


def test_very_wide_rows_use_the_large_shared_memory_path():
    from footprints_b200.preprocess import ResizeToTensor
    rng = np.random.default_rng(6)
    frame = rng.integers(0, 256, (8, 17000, 3), dtype=np.uint8)         # 51 KB per row > the 48 KB default
    out = ResizeToTensor((32, 640))(frame)
    assert torch.equal(out.cpu(), _pillow(frame, 32, 640))


def test_wide_rows_with_few_taps_use_the_row_group_kernel_with_opt_in_shared_memory():
    from footprints_b200.preprocess import ResizeToTensor
    rng = np.random.default_rng(8)
    frame = rng.integers(0, 256, (8, 13000, 3), dtype=np.uint8)         # 31 taps, 52 KB per staged RGBX row
    out = ResizeToTensor((16, 2600))(frame)
    assert torch.equal(out.cpu(), _pillow(frame, 16, 2600))


@pytest.mark.parametrize('hin,win,ho,wo', [(375, 1242, 192, 640), (100, 200, 192, 640), (33, 1001, 63, 97)])
def test_first_generation_kernels_give_the_same_bits(hin, win, ho, wo):
    """FP_RESIZE_ROWS=0: the one-row-per-CTA / one-output-row-per-thread kernels (also the fallback for more than 32 taps)."""
    from footprints_b200 import _lib
    from footprints_b200.preprocess import ResizeToTensor
    rng = np.random.default_rng(hin)
    frames = rng.integers(0, 256, (2, hin, win, 3), dtype=np.uint8)
    default = ResizeToTensor((ho, wo))(frames)
    os.environ['FP_RESIZE_ROWS'] = '0'
    _lib.load().fp_reload_options()
    try:
        first = ResizeToTensor((ho, wo))(frames)
    finally:
        os.environ.pop('FP_RESIZE_ROWS')
        _lib.load().fp_reload_options()
    assert torch.equal(default, first)
    assert torch.equal(default[1].cpu(), _pillow(frames[1], ho, wo))


def test_bad_inputs_raise():
    from footprints_b200.preprocess import ResizeToTensor
    rt = ResizeToTensor((64, 96))
    with pytest.raises(RuntimeError):
        rt(np.zeros((10, 10, 3), np.float32))
    with pytest.raises(RuntimeError):
        rt(np.zeros((10, 10, 4), np.uint8))


def test_network_input_from_the_gpu_edge_gives_the_same_prediction():
    """frame -> ResizeToTensor -> FootprintNetwork equals frame -> Pillow -> ToTensor -> .cuda() -> FootprintNetwork."""
    import footprints_b200
    from footprints_b200.preprocess import ResizeToTensor
    torch.manual_seed(0)
    net = footprints_b200.FootprintNetwork(pretrained=False).cuda().eval()
    rng = np.random.default_rng(7)
    frame = rng.integers(0, 256, (375, 1242, 3), dtype=np.uint8)
    with torch.no_grad():
        a = net(ResizeToTensor((192, 640))(frame)[None])['1/1']
        b = net(_pillow(frame, 192, 640)[None].cuda())['1/1']
    assert torch.equal(a, b)
