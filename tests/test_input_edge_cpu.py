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


def _model_of_the_row_group_kernels(frame, ho, wo):
    """Python model of the ADDRESSING of the default CUDA kernels (csrc/input_edge.cu), not of Pillow: rows of a CTA staged back to
    back with KS zero words behind the last one and KS unconditional taps per output (weights past an output's tap count are
    zero, so a tap may read the next row's pixels); two output rows per thread fed from the union of their source rows through
    zero-padded weight pairs.  If this model reproduces the oracle, those shortcuts are exact."""
    from footprints_b200.preprocess import tap_table
    hin, win = frame.shape[:2]
    ks, bh, th = tap_table(win, wo)
    kv, bv, tv = tap_table(hin, ho)
    y0, y1 = int(bv[0, 0]), int(bv[-1, 0] + bv[-1, 1])
    ht = y1 - y0
    bv = bv.copy()
    bv[:, 0] -= y0
    half = 1 << (oracle.PRECISION_BITS - 1)
    clip = lambda a: np.clip(a.astype(np.int32) >> oracle.PRECISION_BITS, 0, 255).astype(np.uint8)      # noqa: E731
    tmp = np.zeros((ht, wo, 3), np.uint8)
    rows_per_cta = 6
    for r0 in range(0, ht, rows_per_cta):
        nr = min(rows_per_cta, ht - r0)
        flat = np.zeros((nr * win + ks, 3), np.int64)
        flat[:nr * win] = frame[y0 + r0:y0 + r0 + nr].reshape(-1, 3)
        for xo in range(wo):
            for r in range(nr):
                acc = np.full(3, half, np.int64)
                for t in range(ks):
                    acc += flat[r * win + bh[xo, 0] + t] * int(th[t, xo])
                tmp[r0 + r, xo] = clip(acc)
    out = np.zeros((ho, wo, 3), np.uint8)
    for yo0 in range(0, ho, 2):
        two = yo0 + 1 < ho
        first0, n0 = int(bv[yo0, 0]), int(bv[yo0, 1])
        first1, n1 = (int(bv[yo0 + 1, 0]), int(bv[yo0 + 1, 1])) if two else (first0, 0)
        shift = first1 - first0
        assert shift >= 0
        span = max(n0, shift + n1)
        assert span <= 64 and first0 + span <= ht
        a0 = np.full((wo, 3), half, np.int64)
        a1 = a0.copy()
        for j in range(span):
            k0 = int(tv[j, yo0]) if j < n0 else 0
            k1 = int(tv[j - shift, yo0 + 1]) if 0 <= j - shift < n1 else 0
            v = tmp[first0 + j].astype(np.int64)
            a0 += v * k0
            a1 += v * k1
        out[yo0] = clip(a0)
        if two:
            out[yo0 + 1] = clip(a1)
    return out


@pytest.mark.parametrize('hin,win,ho,wo', [(37, 53, 32, 64), (20, 30, 13, 17), (50, 41, 33, 64), (12, 20, 25, 7), (9, 9, 9, 9)])
def test_row_group_kernel_addressing_is_exact(hin, win, ho, wo):
    rng = np.random.default_rng(hin + wo)
    frame = rng.integers(0, 256, (hin, win, 3), dtype=np.uint8)
    assert np.array_equal(_model_of_the_row_group_kernels(frame, ho, wo), oracle.resize_lanczos_u8(frame, (ho, wo)))
