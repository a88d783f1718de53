"""Input edge, CPU side: the oracle's restatement of Pillow's fixed-point Lanczos resampling is pinned against
Pillow itself -- the implementation the reference calls through transforms.Resize(..., Image.ANTIALIAS)
(predict_simple.py:41-42, datasets/inference_dataset.py:25-27) -- and the host's tap tables equal the oracle's."""
import os
import sys

import numpy as np
import pytest
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pil_resample as oracle  # noqa: E402

GEOMETRIES = [(375, 1242, 192, 640),      # KITTI frame -> kitti model input
              (370, 1226, 192, 640),
              (480, 640, 512, 640),       # upscaling rows only (matterport / handheld geometry)
              (100, 200, 192, 640),       # upscaling both axes
              (192, 640, 192, 640),       # identity
              (1080, 1920, 192, 640),
              (37, 53, 32, 64)]


@pytest.mark.parametrize('hin,win,ho,wo', GEOMETRIES)
def test_oracle_resize_is_bit_exact_with_pillow(hin, win, ho, wo):
    rng = np.random.default_rng(hin * 7 + win)
    frame = rng.integers(0, 256, (hin, win, 3), dtype=np.uint8)
    ours = oracle.resize_lanczos_u8(frame, (ho, wo))
    ref = np.asarray(Image.fromarray(frame).resize((wo, ho), Image.LANCZOS))
    assert np.array_equal(ours, ref)


def test_oracle_matches_pillow_on_extreme_frames():
    """Saturation (clip8 both ways: Lanczos overshoots on black/white edges) and constant frames."""
    y, x = np.mgrid[0:375, 0:1242]
    checker = (((x // 7 + y // 5) % 2) * 255).astype(np.uint8)
    for frame in (np.stack([checker, 255 - checker, checker], -1), np.full((375, 1242, 3), 255, np.uint8),
                  np.zeros((375, 1242, 3), np.uint8)):
        ours = oracle.resize_lanczos_u8(frame, (192, 640))
        ref = np.asarray(Image.fromarray(frame).resize((640, 192), Image.LANCZOS))
        assert np.array_equal(ours, ref)


def test_oracle_to_tensor_is_torchvision_to_tensor():
    import torch
    frame = np.arange(256 * 3, dtype=np.uint32).reshape(16, 16, 3).astype(np.uint8)
    ref = torch.from_numpy(frame).permute(2, 0, 1).contiguous().to(torch.float32).div(255).numpy()   # transforms.ToTensor
    assert np.array_equal(oracle.to_tensor(frame), ref)


def test_reference_test_image_resizes_bit_exactly():
    path = os.path.join(ROOT, 'baseline', '_ref', 'test_data', 'cyclist.jpg')
    if not os.path.exists(path):
        pytest.skip('no reference checkout (scripts/install_reference.sh)')
    img = Image.open(path).convert('RGB')
    ours = oracle.resize_lanczos_u8(np.asarray(img), (192, 640))
    assert np.array_equal(ours, np.asarray(img.resize((640, 192), Image.LANCZOS)))


@pytest.mark.parametrize('n_in,n_out', [(1242, 640), (375, 192), (100, 192), (640, 640), (1920, 640), (53, 64)])
def test_host_tap_tables_equal_the_oracle(n_in, n_out):
    from footprints_b200.preprocess import tap_table
    k1, b1, t1 = tap_table(n_in, n_out)
    k2, b2, t2 = oracle.precompute_coeffs(n_in, n_out)
    assert k1 == k2 and np.array_equal(b1, b2) and np.array_equal(t1.T, t2)
    assert t1.dtype == np.int32 and b1.dtype == np.int32 and t1.shape == (k1, n_out)


def test_resize_has_no_cpu_fallback():
    from footprints_b200.preprocess import ResizeToTensor
    with pytest.raises(RuntimeError):
        ResizeToTensor((192, 640), device='cpu')
