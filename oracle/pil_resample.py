"""TEST INFRASTRUCTURE — CPU restatement of the input edge of the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

The reference turns a uint8 RGB frame into the network input with
    transforms.Resize((H, W), interpolation=Image.ANTIALIAS)   predict_simple.py:41-42,54; datasets/inference_dataset.py:25-27,46
    transforms.ToTensor()                                       predict_simple.py:43,55;    datasets/inference_dataset.py:47
The arithmetic lives in a third-party dependency that is not under /root/reference: Pillow (`pillow=6.2.1` in the
reference's environment.yml; 12.2.0 in this image), `src/libImaging/Resample.c`.  Its published algorithm for 8-bit images,
restated here:

  * per output coordinate xx: scale = in/out, filterscale = max(scale, 1), support = 3 * filterscale (Lanczos a=3),
    centre = (xx + 0.5) * scale, window [int(centre - support + 0.5), int(centre + support + 0.5)) clipped to the image,
    weights w = lanczos((x - centre + 0.5) / filterscale) normalised to sum 1 in double precision;
  * weights become 22-bit fixed point: int(+-0.5 + w * 2^22) (truncation toward zero after the half offset);
  * a horizontal pass over the source rows the vertical pass will need, then a vertical pass; each output byte is
    clip8((2^21 + sum(pixel * k)) >> 22), so the intermediate image is rounded to uint8;
  * ToTensor: HWC uint8 -> CHW float32 / 255.

Parity is pinned against Pillow itself (tests/test_input_edge_cpu.py: bit-exact on random frames, KITTI and test-image
geometries, up- and down-scaling), which is the implementation the reference calls.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box.  Returns (ksize, bounds [out,2] int32
    (first tap, tap count), kk [out,ksize] int32 fixed-point weights)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _pass(img, axis, bounds, kk):
    """One resampling pass along `axis` of a [..] uint8 array."""
    img = np.moveaxis(img, axis, 0)
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.uint8)
    for xx in range(bounds.shape[0]):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(n):
            acc += img[x0 + t].astype(np.int64) * int(kk[xx, t])
        out[xx] = _clip8(acc.astype(np.int32))          # the C code accumulates in int32
    return np.moveaxis(out, 0, axis)


def resize_lanczos_u8(frame, out_hw):
    """frame [H,W,C] uint8 -> [Ho,Wo,C] uint8, Image.resize((Wo,Ho), Image.LANCZOS)."""
    hin, win = frame.shape[:2]
    ho, wo = out_hw
    _, bh, kh = precompute_coeffs(win, wo)
    _, bv, kv = precompute_coeffs(hin, ho)
    y0 = int(bv[0, 0])
    y1 = int(bv[-1, 0] + bv[-1, 1])
    tmp = _pass(frame[y0:y1], 1, bh, kh)
    bv = bv.copy()
    bv[:, 0] -= y0
    return _pass(tmp, 0, bv, kv)


def to_tensor(frame_u8):
    """transforms.ToTensor on a uint8 HWC frame: CHW float32 in [0,1]."""
    return (np.ascontiguousarray(frame_u8.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)).astype(np.float32)


def resize_to_tensor(frame, out_hw):
    return to_tensor(resize_lanczos_u8(frame, out_hw))
